"""Drop-in `VideoDepthAnything` (reference: video_depth_anything/video_depth.py:36-254 and its metric twin
metric_depth/video_depth_anything/video_depth.py:35-154): same constructor kwargs, same state-dict keys, same
`forward` / `infer_video_depth` signatures and error behaviour, with every operator executed by libvda's
sm_100a kernels.  There is no CPU path: calling `forward` without the CUDA extension or off-GPU raises.

The long-video driver does everything after the upload of the uint8 frames on the device: window gather + cv2-style
INTER_CUBIC resize + normalisation (util/transform.py), per-window forward, output resize, key-frame least-squares
alignment, clamp and 8-frame cross-fade; frames stream in and finished depths stream out through pinned staging
buffers while the windows compute.
"""
from __future__ import annotations

import os
import queue
import threading
import time
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .engine import Engine
from .synth import ENCODER_DIMS, synth_state_dict
from .windows import (INFER_LEN, INTERP_LEN, KEYFRAMES, OVERLAP, aligner_ring_len, aligner_ring_pos, get_resize_hw,
                      plan_feature_cache, upload_ring_plan, window_source_indices)


VALIDATION_DTYPE = torch.float16     # operand type of the `fp32=True` path (see _ensure_engine)


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
        self._engine_val: Optional[Engine] = None        # fp32=True (validation precision), built on first use
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
        for eng in (self._engine, self._engine_val):
            if eng is not None:
                eng.load(self._sd)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def to(self, device=None, *a, **k):
        if device is not None and not isinstance(device, torch.dtype):
            dev = torch.device(device)
            if dev.type == "cuda" and dev.index is None:     # 'cuda' and 'cuda:<current>' are the same engine
                dev = torch.device("cuda", torch.cuda.current_device())
            self._device = dev
            if self._device.type == "cuda":
                self._ensure_engine()
        return self

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def _ensure_engine(self, validation: bool = False) -> Engine:
        """The engine of the model's dtype; `validation=True`: the high-precision engine behind `fp32=True`: fp16
        activations (11 mantissa bits instead of bf16's 8), every weight matrix as a hi | lo pair of fp16 matrices (22
        bits; the GEMMs walk A twice), fp32 accumulation, residual streams, statistics and softmax.  Built lazily, holds
        its own packed copy of the weights, about twice the tensor work of the fast path."""
        if self._device.type != "cuda":
            raise RuntimeError("VideoDepthAnything (B200 engine) has no CPU path: call .to('cuda') first")
        if validation:
            if self._engine_val is None or self._engine_val.device != self._device:
                from . import _lib
                _lib.load()
                with torch.cuda.device(self._device):
                    self._engine_val = Engine(self.encoder, self.features, self.out_channels, VALIDATION_DTYPE,
                                              self._device, self.num_frames, weight_split=True)
                    self._engine_val.load(self._sd)
            return self._engine_val
        if self._engine is None or self._engine.device != self._device:
            from . import _lib
            _lib.load()                                  # raises if libvda.so is missing
            with torch.cuda.device(self._device):
                self._engine = Engine(self.encoder, self.features, self.out_channels, self.dtype, self._device,
                                      self.num_frames)
                self._engine.load(self._sd)
        return self._engine

    # ---- forward -----------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, stages=None, fp32: bool = False) -> torch.Tensor:
        """x [B,T,3,H,W] -> depth [B,T,H,W] fp32, >= 0 (video_depth.py:89-164).  `fp32=True`: validation precision."""
        eng = self._ensure_engine(validation=bool(fp32))
        with torch.cuda.device(self._device):
            return eng.forward(x.to(self._device), stages)

    # ---- long-video driver -------------------------------------------------------------------
    @torch.no_grad()
    def infer_video_depth(self, frames: np.ndarray, target_fps, input_size=518, device="cuda", fp32=False,
                          window_ids: Optional[Sequence[int]] = None, raw_only=False, reuse_features: bool = True,
                          aligner: Optional["WindowAligner"] = None):
        """frames uint8 [N,H0,W0,3] -> (float32 [N,H0,W0], target_fps)   (video_depth.py:166-254).

        Everything after the upload of the uint8 frames runs on the device: per-window gather + cv2-style
        INTER_CUBIC resize + normalisation (one kernel), forward (CUDA-graph replay), output resize, key-frame
        least squares, clamp and cross-fade.  Frames are uploaded in chunks through pinned staging buffers while
        earlier windows compute, finished depth frames stream back the same way; the host never touches a pixel.
        `fp32=True` (the reference disables autocast, video_depth.py:203-205 / benchmark/infer/infer.py:58) selects the
        validation-precision engine: fp16 activations, weights as hi | lo fp16 pairs, fp32 accumulation, residual streams,
        statistics and softmax (<= 1e-3 of the fp32 reference on full vitl AND vits windows, tests/test_forward_gpu.py),
        whatever the model's own dtype; the flag is not silently ignored.
        `window_ids` / `raw_only` are the multi-GPU hooks (parallel.py): compute only those windows and return
        the raw, resized per-window depths [len(window_ids),32,H0,W0] on the device, skipping alignment; with
        `aligner` the windows are pushed into that WindowAligner instead (the caller finishes it) and None is returned.
        `reuse_features`: the DINOv2 encoder is per-frame and 10 of a window's 32 slots repeat frames of earlier
        windows (:200-201), so each source frame is encoded once and its four tap features are kept on the device
        until no later window needs them (FeatureCache); every kernel is batch-invariant, so the result is
        bit-identical to `reuse_features=False` (tests/test_forward_gpu.py)."""
        if torch.device(device).type != "cuda":
            raise RuntimeError("infer_video_depth: the B200 engine has no CPU path (device must be 'cuda')")
        if frames.ndim != 4 or frames.shape[3] != 3 or frames.dtype != np.uint8:
            raise ValueError("frames must be uint8 [N,H0,W0,3]")
        self.to(device)
        eng = self._ensure_engine(validation=bool(fp32))
        n = frames.shape[0]
        h0, w0 = frames.shape[1:3]
        nh, nw = get_resize_hw(h0, w0, input_size)
        wins = window_source_indices(n)
        ids = list(range(len(wins))) if window_ids is None else list(window_ids)
        with torch.cuda.device(self._device):
            if not ids:
                if aligner is not None:
                    return None
                return torch.empty(0, INFER_LEN, h0, w0, device=self._device) if raw_only else \
                    (np.empty((0, h0, w0), np.float32), target_fps)
            needed = sorted({i for k in ids for i in wins[k]})
            # upload ring: windows in increasing order only; not for the two-phase sharded driver (raw_only), which keeps the
            # raw stacks of all its windows on the device anyway and whose per-rank share shrinks with the rank count
            up = FrameUploader(frames, needed, self._device, ring=not raw_only and all(a < b for a, b in zip(ids, ids[1:])))
            # raw stack allocated once: a fresh 34 MB block per window costs a cudaMalloc each (~10 ms at 518x518)
            raws = torch.empty(len(ids), INFER_LEN, h0, w0, dtype=torch.float32, device=self._device) if raw_only else None
            own_aligner = aligner is None and not raw_only
            if own_aligner:
                aligner = WindowAligner(n, h0, w0, self._device, "identity" if self.metric else "affine")
            cache = FeatureCache(eng, up, nh, nw, [wins[k] for k in ids]) if reuse_features else None
            if cache is None:     # all index lists in one upload (a per-window pageable H2D would sync the stream)
                idx_all = torch.tensor([[up.slot[i] for i in wins[k]] for k in ids], dtype=torch.int32).pin_memory() \
                    .to(self._device, non_blocking=True)
            for j, k in enumerate(ids):
                up.ensure(max(wins[k]))                                      # H2D of the chunks this window needs
                if cache is not None:
                    d = cache.window(j)                                      # [32,nh,nw] fp32 (graph-owned buffer)
                else:
                    x = ops.preprocess_frames(up.dev, idx_all[j], nh, nw).unsqueeze(0)  # [1,32,3,nh,nw]   (:197-201)
                    up.reads_done(wins[k])
                    d = eng.forward(x)[0]                                    # [32,nh,nw] fp32   (:203-205)
                if (nh, nw) != (h0, w0):
                    d = ops.bilinear_f32(d, h0, w0, out=raws[j] if raw_only else None)   # video_depth.py:208
                elif raw_only:
                    raws[j].copy_(d)
                if not raw_only:
                    aligner.push(d)
            up.close()
            if raw_only:
                return raws
            if not own_aligner:
                return None
            return aligner.result(), target_fps


class FeatureCache:
    """Per-frame encoder features shared by overlapping windows.  Window k's slots are source frames
    [0, 22k-10, 22k+2 .. 22k+31] (windows.window_source_indices): frame 0, one key frame and 8 overlap frames were
    already encoded for window k-1, and padded tails repeat the last frame, so only the frames not seen yet go
    through preprocessing + the encoder (22 of 32 in steady state = 31 % fewer encoder FLOPs).  Features live in
    32 device slots per tap ([slot, P, D] 16-bit); frames the current window does not use are evicted (later
    windows only ever reuse frames of their predecessor)."""

    def __init__(self, eng: Engine, up: "FrameUploader", nh: int, nw: int, windows: Sequence[Sequence[int]]):
        self.eng, self.up, self.nh, self.nw = eng, up, nh, nw
        self.hp, self.wp = nh // 14, nw // 14
        P = self.hp * self.wp
        self.store = [torch.empty(INFER_LEN, P, eng.D, dtype=eng.hdtype, device=eng.device) for _ in range(4)]   # taps: head operand type
        self.encoded = 0                                  # frames that went through the encoder (for reporting)
        # The slot bookkeeping depends on the window list only, so the index lists of ALL windows are planned on the
        # host now and uploaded once: per window [uploaded-frame index of each new frame | its cache slot | slot of
        # each of the 32 window positions].  (A per-window pageable H2D copy synchronises the stream and drains the
        # GPU between windows.)
        table = np.zeros((max(len(windows), 1), 3 * INFER_LEN), dtype=np.int32)
        self.n_new, self.missing = [], []
        for j, (missing, slots, positions) in enumerate(plan_feature_cache(windows)):
            n = len(missing)
            self.n_new.append(n)
            self.missing.append(list(missing))
            table[j, :n] = [up.slot[f] for f in missing]
            table[j, INFER_LEN:INFER_LEN + n] = slots
            table[j, 2 * INFER_LEN:] = positions
        self.table = torch.from_numpy(table).pin_memory().to(eng.device, non_blocking=True)

    def window(self, j: int) -> torch.Tensor:
        """Depths [32, nh, nw] of the j-th planned window (a graph-owned buffer: consume before the next call)."""
        eng, P, n = self.eng, self.hp * self.wp, self.n_new[j]
        row = self.table[j]
        if n:
            x = ops.preprocess_frames(self.up.dev, row[:n], self.nh, self.nw)          # [n,3,nh,nw]   (:197-198)
            self.up.reads_done(self.missing[j])
            taps = eng.encode_frames(x)
            for i in range(4):
                ops.copy_frames(taps[i].view(n, P, eng.D), None, self.store[i], row[INFER_LEN:INFER_LEN + n], n)
            self.encoded += n
        head_in = eng.head_static_inputs(INFER_LEN, self.hp, self.wp)
        for i in range(4):
            ops.copy_frames(self.store[i], row[2 * INFER_LEN:], head_in[i].view(INFER_LEN, P, eng.D), None, INFER_LEN)
        return eng.head_frames(head_in, INFER_LEN, self.hp, self.wp)


_PINNED: dict = {}
_PINNED_LOCK = threading.Lock()


def _pinned_acquire(key, count: int, shape, dtype):
    """Pinned host staging buffers, cached across calls (cudaHostAlloc of a few hundred MB costs tens of ms and
    synchronises the device; a serving loop calls infer_video_depth once per video).  One cached set per (purpose,
    frame size); a caller that finds the set taken by another live object (concurrent calls from several threads)
    gets a private allocation.  Returns (buffers, token); hand the token back with _pinned_release."""
    with _PINNED_LOCK:
        ent = _PINNED.get(key)
        if ent is not None and not ent["busy"] and len(ent["bufs"]) >= count and tuple(ent["bufs"][0].shape) == tuple(shape):
            ent["busy"] = True
            return ent["bufs"][:count], key
        if ent is not None and ent["busy"]:
            return [torch.empty(*shape, dtype=dtype).pin_memory() for _ in range(count)], None
        bufs = [torch.empty(*shape, dtype=dtype).pin_memory() for _ in range(count)]
        _PINNED[key] = {"bufs": bufs, "busy": True}
        return bufs, key


def _pinned_release(token) -> None:
    if token is not None:
        with _PINNED_LOCK:
            if token in _PINNED:
                _PINNED[token]["busy"] = False


class FrameUploader:
    """Chunked, asynchronous H2D of the uint8 frames a rank needs (pinned double buffer, copy stream); windows wait
    only for the chunks that hold their frames, so the upload overlaps the compute of earlier windows.

    Device memory is bounded: the first needed frame (source frame 0: slot 0 of every window, video_depth.py:200) keeps a
    slot of its own, all others go through a ring of RING_CHUNKS x CHUNK frames.  A window reads frame 0 and frames inside
    a span of 42 consecutive source frames, windows are visited in increasing order and a chunk is only uploaded when the
    current window needs it, so the chunk a new upload replaces (RING_CHUNKS chunks older) is never read again; the copy
    waits for the kernels that read it (`reads_done`).  `ring=False` (windows not in increasing order): everything stays."""

    CHUNK = 64
    RING_CHUNKS = 4 if os.environ.get("VDA_VIDEO_RINGS", "1") != "0" else 1 << 20    # VDA_VIDEO_RINGS=0: keep everything

    def __init__(self, frames: np.ndarray, needed: List[int], device, ring: bool = True):
        self.frames, self.needed, self.device = frames, needed, device
        self.pos = {i: j for j, i in enumerate(needed)}            # position in upload order
        self.ring, n_slots, slots, self.chunks = upload_ring_plan(len(needed), self.CHUNK, self.RING_CHUNKS if ring else 1 << 30)
        self.slot = {i: slots[j] for i, j in self.pos.items()}     # device slot of every needed source frame
        h0, w0 = frames.shape[1:3]
        self.dev = torch.empty(n_slots, h0, w0, 3, dtype=torch.uint8, device=device)
        self.stage, self._stage_token = _pinned_acquire(("upload", h0, w0), 2, (self.CHUNK, h0, w0, 3), torch.uint8)
        self.stage_free = [None, None]           # event: staging buffer consumed by its H2D copy
        self.read_ev = [None] * (self.RING_CHUNKS if self.ring else 0)   # event: the kernels that read the chunk in this ring region are queued
        self.stream = torch.cuda.Stream(device=device)
        # `dev` comes from the caching allocator of the compute stream and may be a block that kernels of an earlier call
        # (still queued on that stream: the raw_only / aligner paths return without synchronising) are yet to read: the
        # copy stream starts after everything issued so far, and the block is not recycled before the copies are done
        self.stream.wait_stream(torch.cuda.current_stream(device))
        self.dev.record_stream(self.stream)
        self.uploaded = 0                        # number of `needed` entries issued so far
        self.chunk_no = 0

    def _region(self, p: int) -> int:
        """Ring region of upload position p >= 1."""
        return ((p - 1) // self.CHUNK) % self.RING_CHUNKS

    def ensure(self, max_frame: int) -> None:
        """Issue uploads until source frame `max_frame` is covered, then make the current stream wait for them."""
        target = self.pos[max_frame] + 1
        last_ev = None
        while self.uploaded < target:
            lo, hi = self.chunks[self.chunk_no]
            b = self.chunk_no & 1
            if self.stage_free[b] is not None:
                self.stage_free[b].synchronize()
            src_idx = self.needed[lo:hi]
            st = self.stage[b][:hi - lo]
            if src_idx[-1] - src_idx[0] == hi - lo - 1:                      # contiguous run: plain slice copy
                np.copyto(st.numpy(), self.frames[src_idx[0]:src_idx[-1] + 1])
            else:
                np.take(self.frames, src_idx, axis=0, out=st.numpy())
            d0 = self.slot[src_idx[0]]
            with torch.cuda.stream(self.stream):
                if self.ring and lo > 0 and self.read_ev[self._region(lo)] is not None:
                    self.stream.wait_event(self.read_ev[self._region(lo)])   # readers of the chunk this one replaces
                self.dev[d0:d0 + hi - lo].copy_(st, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self.stage_free[b] = ev
            last_ev = ev
            self.uploaded = hi
            self.chunk_no += 1
        if last_ev is not None:
            torch.cuda.current_stream().wait_event(last_ev)

    def reads_done(self, frame_ids: Sequence[int]) -> None:
        """The kernels that read these source frames from `dev` have been queued on the current stream."""
        if not self.ring:
            return
        regions = {self._region(self.pos[f]) for f in frame_ids if self.pos[f] > 0}
        if regions:
            ev = torch.cuda.Event()
            ev.record()
            for r in regions:
                self.read_ev[r] = ev

    def close(self) -> None:
        """All uploads have been issued: give the staging buffers back once their H2D copies are done."""
        for ev in self.stage_free:
            if ev is not None:
                ev.synchronize()
        _pinned_release(self._stage_token)
        self._stage_token = None


class HostDrain:
    """Device -> host streaming of finished fp32 frames: D2H into pinned staging buffers on a copy stream, and a drain
    thread (numpy copies release the GIL; every batch is split over a small pool) moves them into the result array,
    so the launching thread never waits for a host memcpy.  The result array is first-touched in the background (one
    write per page): the page faults of a fresh multi-GB allocation otherwise sit inside the drain copies and divide
    their bandwidth by ~5 (tools/microbench/host_touch.py)."""

    STAGES = 4
    COPY_THREADS = 6

    def __init__(self, host: np.ndarray, device, touch: Optional[slice] = None, copy_threads: Optional[int] = None,
                 touch_threads: Optional[int] = None, direct: bool = False):
        """`copy_threads` / `touch_threads`: host threads of the drain copies / of the background first-touch (default
        6 / 6 for a single process; the sharded driver runs one HostDrain per rank on the same host, see parallel.py).
        `direct`: page-lock the `touch` slice of the result array in the background (cudaHostRegister, started 50 ms
        after construction so that the first windows are already queued on the GPU) and let the GPU write finished
        frames straight into it -- no pinned staging, no host copy.  For the two-phase sharded driver, whose frames all
        leave in the tail: the staging -> result copies of all ranks share the host's memory bandwidth (measured: 2.2 GB in
        ~55 ms whatever the rank count).  Falls back to the staged path if the registration fails."""
        self.host, self.device = host, device
        self.direct, self._reg_ok, self._reg_done, self._reg_note = direct, False, threading.Event(), ""
        self._reg_slice = touch if touch is not None else slice(0, host.shape[0])
        if direct:
            threading.Thread(target=self._register, daemon=True).start()
            touch_threads = 1
            touch = slice(0, 0)                   # the registration faults the pages in itself
        self.COPY_THREADS = copy_threads or self.COPY_THREADS
        touch_threads = touch_threads or self.COPY_THREADS
        h0, w0 = host.shape[1:]
        self.stage, self._stage_token, self._stage_shape = None, None, (INFER_LEN, h0, w0)   # staging: on first staged send
        self.stage_free = [threading.Event() for _ in range(self.STAGES)]
        for e in self.stage_free:
            e.set()
        self.copy_stream = torch.cuda.Stream(device=device)
        self.batch_no = 0
        self.jobs: "queue.Queue" = queue.Queue()
        self.error = None
        self.copied_bytes, self.copy_seconds, self.touch_wait_seconds, self.event_wait_seconds = 0, 0.0, 0.0, 0.0   # trace
        self.pool = ThreadPoolExecutor(self.COPY_THREADS)
        flat = (host if touch is None else host[touch]).reshape(-1)
        cuts = np.linspace(0, flat.size, touch_threads + 1).astype(np.int64)
        self.touch_pool = ThreadPoolExecutor(touch_threads)

        def first_touch(a):
            try:     # background work: stay out of the way of the launching thread (Linux: per-thread nice value)
                os.setpriority(os.PRIO_PROCESS, threading.get_native_id(), 10)
            except (OSError, AttributeError):
                pass
            a[::1024].fill(0)

        self.touch = [self.touch_pool.submit(first_touch, flat[cuts[i]:cuts[i + 1]]) for i in range(touch_threads)]
        self.drainer = threading.Thread(target=self._drain_loop, daemon=True)
        self.drainer.start()

    def _register(self) -> None:
        try:
            torch.cuda.set_device(self.device)   # a new thread starts on device 0: register in THIS rank's context
            time.sleep(0.05)
            rt = torch.cuda.cudart()
            region = self.host[self._reg_slice]
            if region.size:
                ptr = region.ctypes.data
                rc = int(rt.cudaHostRegister(ptr, region.nbytes, 0))
                if rc == 712:                    # cudaErrorHostMemoryAlreadyRegistered: a stale entry for this address range
                    rt.cudaHostUnregister(ptr)
                    rc = int(rt.cudaHostRegister(ptr, region.nbytes, 0))
                self._reg_note = f"cudaHostRegister rc={rc}"
                if rc == 0:
                    self._host_t = torch.from_numpy(region)
                    self._reg_ok = True
        except Exception as exc:                 # noqa: BLE001  (no registration: staged copies)
            self._reg_ok = False
            self._reg_note = f"{type(exc).__name__}: {exc}"
        finally:
            if not self._reg_ok:
                # staged path after all: get its pinned buffers and fault the result pages in now, in the background,
                # not in the tail (first-touch page faults divide the copy bandwidth by 5)
                try:
                    self.stage, self._stage_token = _pinned_acquire(("drain",) + self._stage_shape[1:], self.STAGES,
                                                                    self._stage_shape, torch.float32)
                    self.host[self._reg_slice].reshape(-1)[::1024].fill(0)
                except Exception:                # noqa: BLE001
                    pass
            self._reg_done.set()

    def unregister_later(self, close_after=None) -> None:
        """Release the page lock in the background, half a second from now: the data is in place, but
        cudaHostUnregister holds the context lock for tens of ms per 100 MB -- right behind finish() it delayed the
        caller's next CUDA call (a dist.barrier waited 36 ms), and run back to back with the next call's registration it
        stalled that call's launches by 0.5 s.  The array is kept alive until then, and `close_after` (the rank's
        SharedMemory attachment) is closed only afterwards: NumPy does not hold a buffer export, so closing the mapping
        while it is page-locked leaves a stale registration behind that the next mapping at the same address trips
        over (cudaErrorHostMemoryAlreadyRegistered)."""
        if not (self.direct and self._reg_ok):
            if close_after is not None:
                try:
                    close_after.close()
                except BufferError:
                    pass
            return
        region = self.host[self._reg_slice]
        ptr, keep_ref, dev = region.ctypes.data, [self.host, close_after], self.device
        self.host = None

        def work():
            try:
                torch.cuda.set_device(dev)
                time.sleep(0.5)
                torch.cuda.cudart().cudaHostUnregister(ptr)
            finally:
                shm = keep_ref[1]
                del keep_ref[:]
                if shm is not None:
                    try:
                        shm.close()
                    except BufferError:
                        pass
        threading.Thread(target=work, daemon=True).start()

    def _drain_loop(self) -> None:
        while True:
            job = self.jobs.get()
            if job is None:
                self.jobs.task_done()
                return
            ev, b, lo, hi = job
            try:
                t0 = time.perf_counter()
                ev.synchronize()
                t1 = time.perf_counter()
                if self.touch:                   # a late page-touch write would zero a float of a drained frame
                    for f in self.touch:
                        f.result()
                    self.touch = []
                t2 = time.perf_counter()
                src = self.stage[b].numpy()
                cuts = np.linspace(0, hi - lo, self.COPY_THREADS + 1).astype(int)
                list(self.pool.map(lambda i: np.copyto(self.host[lo + cuts[i]:lo + cuts[i + 1]], src[cuts[i]:cuts[i + 1]]),
                                   range(self.COPY_THREADS)))
                self.event_wait_seconds += t1 - t0
                self.touch_wait_seconds += t2 - t1
                self.copy_seconds += time.perf_counter() - t2
                self.copied_bytes += (hi - lo) * src[0].nbytes
            except BaseException as e:       # surfaced by finish()
                self.error = e
            finally:
                self.stage_free[b].set()
                self.jobs.task_done()

    def send(self, frames: torch.Tensor, host_lo: int) -> "torch.cuda.Event":
        """Queue device frames [m,h0,w0] (final once the current stream gets here) for host rows [host_lo, host_lo+m).
        Returns an event on the copy stream behind the last D2H copy that reads `frames` (a caller that recycles the
        device memory makes its stream wait for it)."""
        if self.direct:
            self._reg_done.wait()
            if self._reg_ok:                     # DMA straight into the page-locked result rows
                lo0 = self._reg_slice.start or 0
                ready = torch.cuda.Event()
                ready.record()
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(ready)
                    self._host_t[host_lo - lo0:host_lo - lo0 + frames.shape[0]].copy_(frames, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(self.copy_stream)
                frames.record_stream(self.copy_stream)
                self.copied_bytes += frames.numel() * 4
                return done
        if self.stage is None:
            self.stage, self._stage_token = _pinned_acquire(("drain",) + self._stage_shape[1:], self.STAGES, self._stage_shape,
                                                            torch.float32)
        for off in range(0, frames.shape[0], INFER_LEN):
            part = frames[off:off + INFER_LEN]
            b = self.batch_no % self.STAGES
            self.stage_free[b].wait()
            self.stage_free[b].clear()
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(ready)
                self.stage[b][:part.shape[0]].copy_(part, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            self.jobs.put((ev, b, host_lo + off, host_lo + off + part.shape[0]))
            self.batch_no += 1
        return ev

    def finish(self) -> np.ndarray:
        if self.direct:
            self._reg_done.wait()
            if self._reg_ok:
                self.copy_stream.synchronize()   # (the caller releases the page lock: unregister_later)
        self.jobs.put(None)
        self.jobs.join()
        self.drainer.join()
        self.pool.shutdown()
        self.touch_pool.shutdown()
        _pinned_release(self._stage_token)       # every drain copy out of the staging buffers has completed
        self._stage_token = None
        if self.error is not None:
            raise self.error
        return self.host


def make_blend_weights(device) -> torch.Tensor:
    step = 1.0 / (INTERP_LEN - 1)                         # get_interpolate_frames (utils/util.py:65-74)
    return torch.tensor([0.0] + [i * step for i in range(1, INTERP_LEN - 1)] + [1.0], dtype=torch.float32, device=device)


class WindowAligner:
    """Sequential scale/shift alignment + cross-fade of consecutive windows on the device
    (video_depth.py:216-252, utils/util.py:40-74).  (scale, shift) never leave the GPU; frames that can no longer
    change (everything but the last 8) are streamed to the host (HostDrain) while the next windows compute.

    Device memory is bounded: the aligned frames live in a ring of RING_SEGS segments of 22 frames (a window adds 22
    frames and cross-fades into the last 8 of its predecessor; everything older has been handed to the D2H pipeline).
    Frame a sits at ring position (a + 12) % ring length, which puts the end of window 0 (32 frames) and every later
    window on segment boundaries, so each window's 22 new frames and its 8-frame cross-fade tail are contiguous; a
    segment is only overwritten after the D2H copies that read it have completed (events of HostDrain.send)."""

    SEG = INFER_LEN - OVERLAP          # 22 new frames per window
    RING_SEGS = 4 if os.environ.get("VDA_VIDEO_RINGS", "1") != "0" else 1 << 20

    def __init__(self, n_frames: int, h0: int, w0: int, device, mode: str = "affine"):
        self.n, self.h0, self.w0, self.device, self.mode = n_frames, h0, w0, device, mode
        self.ring_len = aligner_ring_len(n_frames, self.RING_SEGS)
        self.out = torch.empty(self.ring_len, h0, w0, dtype=torch.float32, device=device)
        self.seg_ev = [None] * (self.ring_len // self.SEG)               # last D2H copy that reads the segment
        self.filled = 0
        self.ref = None                      # [2,h0,w0]: (ref_align[0], ref_align[1])
        self.ss = torch.tensor([1.0, 0.0], dtype=torch.float32, device=device)
        self.scratch = torch.zeros(4 * ops.LSQ_MAX_PARTIALS, dtype=torch.float64, device=device)
        self.blend_w = make_blend_weights(device)
        self.drain = HostDrain(np.empty((n_frames, h0, w0), dtype=np.float32), device)
        self.sent = 0                        # frames already handed to the D2H pipeline

    def _pos(self, a: int) -> int:
        return aligner_ring_pos(a, self.ring_len)

    def _claim(self, r: int, m: int) -> torch.Tensor:
        """Ring frames [r, r+m) for writing: the D2H copies of what they held must have completed."""
        for g in range(r // self.SEG, (r + m - 1) // self.SEG + 1):
            if self.seg_ev[g] is not None:
                torch.cuda.current_stream().wait_event(self.seg_ev[g])
                self.seg_ev[g] = None
        return self.out[r:r + m]

    def _send(self, upto: int) -> None:
        """Stream frames [sent, upto) (final values) to the host."""
        upto = min(upto, self.n)
        while upto > self.sent:
            r = self._pos(self.sent)
            m = min(upto - self.sent, self.ring_len - r)                  # (split where the ring wraps)
            ev = self.drain.send(self.out[r:r + m], self.sent)
            for g in range(r // self.SEG, (r + m - 1) // self.SEG + 1):
                self.seg_ev[g] = ev
            self.sent += m

    def push(self, d: torch.Tensor) -> None:
        """d: raw depths of the next window, fp32 [32,h0,w0]."""
        d = d.contiguous()
        align_len = OVERLAP - INTERP_LEN                                  # 2; kf_align_list = [0, 12]
        if self.filled == 0:
            self._claim(self._pos(0), INFER_LEN).copy_(d)                 # window 0 copied unclamped (:222-225)
            self.ref = torch.stack([d[KEYFRAMES[0]], d[KEYFRAMES[1]]])
            self.filled = INFER_LEN
        else:
            if self.mode == "affine":
                ops.lsq_scale_shift(d[:align_len], self.ref, self.ss, self.scratch)      # :227-232
            rt = self._pos(self.filled - INTERP_LEN)                      # the last 8 frames of the previous segment
            tail = self.out[rt:rt + INTERP_LEN]
            ops.affine_clamp_blend(d[align_len:OVERLAP], self.ss, tail, prev=tail, blend_w=self.blend_w)   # :234-239
            new = self._claim(self._pos(self.filled), self.SEG)
            ops.affine_clamp_blend(d[OVERLAP:], self.ss, new)                                              # :241-244
            ref1 = self.ref[1:2]
            ops.affine_clamp_blend(d[KEYFRAMES[1]:KEYFRAMES[1] + 1], self.ss, ref1)                       # :246-250
            self.filled += self.SEG
        self._send(self.filled - INTERP_LEN)          # the last 8 frames are still cross-faded with the next window

    def result(self) -> np.ndarray:
        self._send(self.n)
        return self.drain.finish()
