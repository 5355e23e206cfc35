"""Drop-in `VideoDepthAnything` (reference: video_depth_anything/video_depth.py:36-254 and its metric twin
metric_depth/video_depth_anything/video_depth.py:35-154): same constructor kwargs, same state-dict keys, same
`forward` / `infer_video_depth` signatures and error behaviour, with every operator executed by libvda's
sm_100a kernels.  There is no CPU path: calling `forward` without the CUDA extension or off-GPU raises.

The long-video driver keeps the reference's host preprocessing (util/transform.py, cv2 INTER_CUBIC) but does
everything after the H2D copy on the device: per-window forward, output resize, key-frame least-squares
alignment, clamp and 8-frame cross-fade, one D2H copy per window of finished frames.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .engine import Engine
from .synth import ENCODER_DIMS, synth_state_dict
from .windows import (INFER_LEN, INTERP_LEN, KEYFRAMES, OVERLAP, get_resize_hw, preprocess_frames,
                      window_source_indices)


class VideoDepthAnything(nn.Module):
    def __init__(self, encoder="vitl", features=256, out_channels=(256, 512, 1024, 1024), use_bn=False,
                 use_clstoken=False, num_frames=32, pe="ape", metric=False, dtype=torch.bfloat16, **ignored):
        """`metric=True` selects the metric_depth driver (identity alignment,
        metric_depth/video_depth_anything/video_depth.py:132).  The fork's dead kwargs
        (num_block/out_channel/conv, video_depth.py:47-49) are accepted and ignored."""
        super().__init__()
        if encoder not in ENCODER_DIMS:
            raise KeyError(encoder)                      # reference: KeyError from the model_zoo dict (dinov2.py:399-404)
        if use_bn or use_clstoken or pe != "ape":
            raise NotImplementedError("only the configurations the reference instantiates are supported: "
                                      "use_bn=False, use_clstoken=False, pe='ape' (run.py:40-43)")
        self.encoder = encoder
        self.intermediate_layer_idx = {"vits": [2, 5, 8, 11], "vitl": [4, 11, 17, 23]}
        self.features, self.out_channels, self.num_frames = features, list(out_channels), num_frames
        self.metric = metric
        self.dtype = dtype
        # parameters live in a plain fp32 state dict with the reference's key set; the engine owns packed copies
        self._sd = synth_state_dict(encoder, features, self.out_channels, num_frames, seed=0)
        self._engine: Optional[Engine] = None
        self._device = torch.device("cpu")

    # ---- nn.Module surface -------------------------------------------------------------------
    def state_dict(self, *a, **k):
        return OrderedDict((k_, v.clone()) for k_, v in self._sd.items())

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._sd if k not in sd]
        unexpected = [k for k in sd if k not in self._sd]
        bad = [k for k in self._sd if k in sd and tuple(sd[k].shape) != tuple(self._sd[k].shape)]
        if bad or (strict and (missing or unexpected)):
            raise RuntimeError(f"Error(s) in loading state_dict for VideoDepthAnything: missing {missing[:5]} "
                               f"unexpected {unexpected[:5]} size mismatch {bad[:5]}")
        for k in self._sd:
            if k in sd:
                self._sd[k] = sd[k].detach().to("cpu", torch.float32).clone()
        if self._engine is not None:
            self._engine.load(self._sd)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def to(self, device=None, *a, **k):
        if device is not None and not isinstance(device, torch.dtype):
            self._device = torch.device(device)
            if self._device.type == "cuda":
                self._ensure_engine()
        return self

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def _ensure_engine(self) -> Engine:
        if self._device.type != "cuda":
            raise RuntimeError("VideoDepthAnything (B200 engine) has no CPU path: call .to('cuda') first")
        if self._engine is None:
            from . import _lib
            _lib.load()                                  # raises if libvda.so is missing
            with torch.cuda.device(self._device):
                self._engine = Engine(self.encoder, self.features, self.out_channels, self.dtype, self._device,
                                      self.num_frames)
                self._engine.load(self._sd)
        return self._engine

    # ---- forward -----------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, stages=None) -> torch.Tensor:
        """x [B,T,3,H,W] -> depth [B,T,H,W] fp32, >= 0 (video_depth.py:89-164)."""
        eng = self._ensure_engine()
        with torch.cuda.device(self._device):
            return eng.forward(x.to(self._device), stages)

    # ---- long-video driver -------------------------------------------------------------------
    @torch.no_grad()
    def infer_video_depth(self, frames: np.ndarray, target_fps, input_size=518, device="cuda", fp32=False,
                          window_ids: Optional[Sequence[int]] = None, raw_only=False):
        """frames uint8 [N,H0,W0,3] -> (float32 [N,H0,W0], target_fps)   (video_depth.py:166-254).

        `fp32` is accepted for signature compatibility; operand precision is the engine's dtype (bf16/fp16
        tensor-core operands, fp32 accumulation / residual / statistics).
        `window_ids` / `raw_only` are the multi-GPU hooks (parallel.py): compute only those windows and return
        the raw, resized per-window depths [len(window_ids),32,H0,W0] on the device, skipping alignment."""
        if torch.device(device).type != "cuda":
            raise RuntimeError("infer_video_depth: the B200 engine has no CPU path (device must be 'cuda')")
        self.to(device)
        eng = self._ensure_engine()
        n = frames.shape[0]
        h0, w0 = frames.shape[1:3]
        nh, nw = get_resize_hw(h0, w0, input_size)
        wins = window_source_indices(n)
        ids = list(range(len(wins))) if window_ids is None else list(window_ids)
        needed = sorted({i for k in ids for i in wins[k]})
        with torch.cuda.device(self._device):
            pre = preprocess_frames(frames, needed, input_size)          # host, each source frame once
            slot = {i: j for j, i in enumerate(needed)}
            pre_dev = torch.from_numpy(pre).pin_memory().to(self._device, non_blocking=True)
            raws = []
            aligner = None if raw_only else WindowAligner(n, h0, w0, self._device,
                                                          "identity" if self.metric else "affine")
            for k in ids:
                idx = torch.tensor([slot[i] for i in wins[k]], device=self._device)
                x = pre_dev.index_select(0, idx).unsqueeze(0)            # [1,32,3,nh,nw]
                d = eng.forward(x)[0]                                    # [32,nh,nw] fp32
                if (nh, nw) != (h0, w0):
                    d = ops.bilinear_f32(d, h0, w0)                      # video_depth.py:208
                if raw_only:
                    raws.append(d)
                else:
                    aligner.push(d)
            if raw_only:
                return torch.stack(raws) if raws else torch.empty(0, INFER_LEN, h0, w0, device=self._device)
            return aligner.result(), target_fps


class WindowAligner:
    """Sequential scale/shift alignment + cross-fade of consecutive windows on the device
    (video_depth.py:216-252, utils/util.py:40-74).  Finished frames are copied to pinned host memory
    asynchronously; (scale, shift) never leave the GPU."""

    def __init__(self, n_frames: int, h0: int, w0: int, device, mode: str = "affine"):
        self.n, self.h0, self.w0, self.device, self.mode = n_frames, h0, w0, device, mode
        k = -(-n_frames // (INFER_LEN - OVERLAP))
        total = k * (INFER_LEN - OVERLAP) + OVERLAP
        self.out = torch.empty(total, h0, w0, dtype=torch.float32, device=device)
        self.filled = 0
        self.ref = None                      # [2,h0,w0]: (ref_align[0], ref_align[1])
        self.ss = torch.tensor([1.0, 0.0], dtype=torch.float32, device=device)
        self.scratch = torch.zeros(4 * ops.LSQ_MAX_PARTIALS, dtype=torch.float64, device=device)
        step = 1.0 / (INTERP_LEN - 1)
        self.blend_w = torch.tensor([0.0] + [i * step for i in range(1, INTERP_LEN - 1)] + [1.0],
                                    dtype=torch.float32, device=device)
        self.scales = []

    def push(self, d: torch.Tensor) -> None:
        """d: raw depths of the next window, fp32 [32,h0,w0]."""
        d = d.contiguous()
        align_len = OVERLAP - INTERP_LEN                                  # 2; kf_align_list = [0, 12]
        if self.filled == 0:
            self.out[:INFER_LEN].copy_(d)                                 # window 0 copied unclamped (:222-225)
            self.ref = torch.stack([d[KEYFRAMES[0]], d[KEYFRAMES[1]]])
            self.filled = INFER_LEN
            return
        if self.mode == "affine":
            ops.lsq_scale_shift(d[:align_len], self.ref, self.ss, self.scratch)      # :227-232
        tail = self.out[self.filled - INTERP_LEN:self.filled]
        ops.affine_clamp_blend(d[align_len:OVERLAP], self.ss, tail, prev=tail, blend_w=self.blend_w)   # :234-239
        new = self.out[self.filled:self.filled + INFER_LEN - OVERLAP]
        ops.affine_clamp_blend(d[OVERLAP:], self.ss, new)                                              # :241-244
        ref1 = self.ref[1:2]
        ops.affine_clamp_blend(d[KEYFRAMES[1]:KEYFRAMES[1] + 1], self.ss, ref1)                       # :246-250
        self.filled += INFER_LEN - OVERLAP

    def result(self) -> np.ndarray:
        host = torch.empty(self.n, self.h0, self.w0, dtype=torch.float32).pin_memory()
        host.copy_(self.out[:self.n], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()
