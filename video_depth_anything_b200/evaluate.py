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


def eval_tae(infs, gts, Ks, poses, max_depth: float, masks=None, device="cuda", jobs_per_call: int = 64) -> float:
    """Temporal alignment error of a sequence, mirroring `eval_TAE` (benchmark/eval/eval_tae.py:109-213) after its file
    loading: infs predicted disparity [T,H,W], gts depth [T,H,W] (already cropped like the prediction), Ks [T,3,3] pinhole
    intrinsics, poses [T,4,4] camera-to-world, masks optional bool [T,H,W].  Least-squares alignment of the disparity over
    the sequence (as eval_sequence), disparity -> clipped depth, then for every consecutive pair the re-projection error
    in both directions (`tae_torch`, :60-107); returns 100 * mean.  The 4x4 pose algebra (2 (T-1) inversions / products,
    :167-196) is host NumPy exactly like the reference; every per-pixel operation runs in libvda (float64)."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("eval_tae: the B200 engine has no CPU path (device must be 'cuda')")
    pred = torch.as_tensor(infs).to(dev, torch.float32).contiguous()
    gt = torch.as_tensor(gts).to(dev)
    if gt.dtype not in (torch.float32, torch.float64):
        gt = gt.to(torch.float64)
    gt = gt.contiguous()
    if pred.dim() != 3 or tuple(pred.shape) != tuple(gt.shape):
        raise ValueError(f"expected [T,H,W] prediction and ground truth of one shape, got {tuple(pred.shape)} / {tuple(gt.shape)}")
    T, H, W = pred.shape
    Ks = np.asarray(Ks, dtype=np.float64)
    poses = np.asarray(poses, dtype=np.float64)
    if Ks.shape != (T, 3, 3) or poses.shape != (T, 4, 4):
        raise ValueError("Ks must be [T,3,3] and poses [T,4,4]")
    if T < 2:
        raise ValueError("eval_tae needs at least two frames")
    prm, src, dst = [], [], []
    for i in range(T - 1):                                     # eval_tae.py:162-196
        T_2_1 = np.linalg.inv(poses[i + 1]) @ poses[i]
        T_1_2 = np.linalg.inv(T_2_1)
        K = Ks[i]
        for M, a, b in ((T_2_1, i, i + 1), (T_1_2, i + 1, i)):
            prm.append(np.concatenate([M[:3, :3].reshape(-1), M[:3, 3], [K[0, 0], K[1, 1], K[0, 2], K[1, 2]]]))
            src.append(a)
            dst.append(b)
    with torch.cuda.device(dev):
        hw = H * W
        out3 = torch.empty(3, dtype=torch.float64, device=dev)
        ss = torch.empty(2, dtype=torch.float64, device=dev)
        scratch = torch.empty(EVAL_LSQ_PARTIALS * 5 + T * EVAL_SLABS * 4, dtype=torch.float64, device=dev)
        _lib.check(lib.vda_eval_sequence(_p(pred), _p(gt), int(gt.dtype == torch.float64), T, hw, float(max_depth), _p(out3),
                                         _p(ss), _p(scratch), _stream()))
        depth = torch.empty(T, hw, dtype=torch.float64, device=dev)
        _lib.check(lib.vda_eval_aligned_depth(_p(pred), _p(ss), float(max_depth), T * hw, _p(depth), _stream()))
        _count(5)
        mk = None
        if masks is not None:
            mk = torch.as_tensor(np.asarray(masks) > 0).to(dev, torch.uint8).reshape(T, hw).contiguous()
        n = len(prm)
        prm_d = torch.from_numpy(np.stack(prm)).to(dev)
        src_d = torch.tensor(src, dtype=torch.int32, device=dev)
        dst_d = torch.tensor(dst, dtype=torch.int32, device=dev)
        errs = torch.empty(n, dtype=torch.float64, device=dev)
        per = max(1, min(jobs_per_call, n))
        winners = torch.empty(per * hw, dtype=torch.int32, device=dev)
        partials = torch.empty(per * EVAL_SLABS * 2, dtype=torch.float64, device=dev)
        for j0 in range(0, n, per):
            j1 = min(n, j0 + per)
            _lib.check(lib.vda_eval_tae(_p(depth), _p(mk), _p(prm_d[j0:]), _p(src_d[j0:]), _p(dst_d[j0:]), j1 - j0, H, W,
                                        _p(winners), _p(partials), _p(errs[j0:]), _stream()))
            _count(4)
        return float(errs.sum().item() / (2 * (T - 1)) * 100.0)
