"""Sequence evaluation on the device, mirroring the reference's `eval_depthcrafter` (benchmark/eval/eval.py:67-122) and
its metric functions (benchmark/eval/metric.py): least-squares alignment of the predicted disparity to 1/gt over the
sequence, then AbsRel / RMSE / delta1 per frame, averaged.  File reading, cropping and resizing of the prediction stay
with the caller (as in eval.py:20-49); the arithmetic runs in libvda (`vda_eval_sequence`, float64 like the reference)."""
from __future__ import annotations

from typing import List, Union

import numpy as np
import torch

from . import _lib
from .ops import _p, _stream, _count

eval_metrics = ["abs_relative_difference", "rmse_linear", "delta1_acc"]     # eval.py:17-21, the order of the result

EVAL_LSQ_PARTIALS, EVAL_SLABS = 592, 64         # include/vda.h


def eval_sequence(infs: Union[np.ndarray, torch.Tensor], gts: Union[np.ndarray, torch.Tensor], max_depth: float,
                  device="cuda", return_alignment: bool = False) -> List[float]:
    """infs: predicted disparity [T,H,W] (cast to float32 like get_infer, eval.py:26-33); gts: depth [T,H,W], float32 or
    float64, zeros / negatives = no ground truth (get_gt maps 0 to -1, eval.py:47).  Returns [AbsRel, RMSE, delta1]."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("eval_sequence: the B200 engine has no CPU path (device must be 'cuda')")
    pred = torch.as_tensor(infs).to(dev, torch.float32).contiguous()
    gt = torch.as_tensor(gts).to(dev)
    if gt.dtype not in (torch.float32, torch.float64):
        gt = gt.to(torch.float64)
    gt = gt.contiguous()
    if pred.dim() != 3 or tuple(pred.shape) != tuple(gt.shape):
        raise ValueError(f"expected [T,H,W] prediction and ground truth of one shape, got {tuple(pred.shape)} / {tuple(gt.shape)}")
    T, hw = pred.shape[0], pred.shape[1] * pred.shape[2]
    with torch.cuda.device(dev):
        out = torch.empty(3, dtype=torch.float64, device=dev)
        ss = torch.empty(2, dtype=torch.float64, device=dev)
        scratch = torch.empty(EVAL_LSQ_PARTIALS * 5 + T * EVAL_SLABS * 4, dtype=torch.float64, device=dev)
        _lib.check(lib.vda_eval_sequence(_p(pred), _p(gt), int(gt.dtype == torch.float64), T, hw, float(max_depth), _p(out),
                                         _p(ss), _p(scratch), _stream()))
        _count(4)
        res = out.cpu().tolist()
        return (res, ss.cpu().tolist()) if return_alignment else res
