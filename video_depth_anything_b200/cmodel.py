"""ctypes wrapper of libvda's handle-level API (vda_create / vda_set_weight / vda_finalize_weights / vda_workspace_bytes /
vda_forward, include/vda.h): what a non-Python host binds.  The Python package itself uses engine.py (same schedule, plus
CUDA-graph replay and the encoder-feature cache); this wrapper exists so that tests can hold the C schedule against it
bit for bit, and as the worked example of INTEGRATION.md §2b."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence

import torch

from . import _lib
from ._lib import check
from .ops import dt_code


class CModel:
    def __init__(self, encoder: str, features: int, out_channels: Sequence[int], num_frames: int = 32,
                 dtype=torch.bfloat16, device="cuda:0"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        oc = (C.c_int32 * 4)(*out_channels)
        self.handle = C.c_void_p()
        check(self.lib.vda_create(encoder.encode(), features, oc, num_frames, dt_code(dtype), self.device.index or 0,
                                  C.byref(self.handle)))
        self._ws = None

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        for k, v in sd.items():
            t = v.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            check(self.lib.vda_set_weight(self.handle, k.encode(), t.data_ptr(), shape, t.dim()))
        with torch.cuda.device(self.device):
            check(self.lib.vda_finalize_weights(self.handle))

    def workspace_bytes(self, B: int, T: int, H: int, W: int) -> int:
        n = self.lib.vda_workspace_bytes(self.handle, B, T, H, W)
        if n < 0:
            raise _lib.VdaError("vda_workspace_bytes failed: " + self.lib.vda_last_error().decode(errors="replace"))
        return n

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x fp32 [B,T,3,H,W] on the model's device -> depth fp32 [B,T,H,W]."""
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 5
        x = x.contiguous()
        B, T, _, H, W = x.shape
        with torch.cuda.device(self.device):
            need = self.workspace_bytes(B, T, H, W)
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            out = torch.empty(B, T, H, W, dtype=torch.float32, device=self.device)
            check(self.lib.vda_forward(self.handle, x.data_ptr(), B, T, H, W, out.data_ptr(), self._ws.data_ptr(),
                                       self._ws.numel(), torch.cuda.current_stream().cuda_stream))
            return out

    def close(self) -> None:
        if self.handle:
            self.lib.vda_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001
            pass
