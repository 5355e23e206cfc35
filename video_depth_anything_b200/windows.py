"""Host-side logic of the long-video driver: window bookkeeping, resize geometry and frame preprocessing.

Reference: video_depth_anything/video_depth.py:166-201 and util/transform.py:62-158."""
from __future__ import annotations

from typing import Iterable, List

import numpy as np

# video_depth.py:29-33 — "infer settings, do not change"
INFER_LEN = 32
OVERLAP = 10
KEYFRAMES = [0, 12, 24, 25, 26, 27, 28, 29, 30, 31]
INTERP_LEN = 8
STEP = INFER_LEN - OVERLAP

_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float64)
_STD = np.array([0.229, 0.224, 0.225], dtype=np.float64)


def num_windows(n_frames: int) -> int:
    return -(-n_frames // STEP)


def window_source_indices(n_frames: int) -> List[List[int]]:
    """Source frame of every slot of every window.  The reference pads the frame list with copies of the last
    frame (video_depth.py:187-191), slices 32 frames every 22, and overwrites the first 10 slots with the
    previous window's KEYFRAMES slots (:200-201).  Unrolled, that recurrence has the closed form
        window 0 : 0..31
        window k : [0, 22k-10, 22k+2, ..., 22k+31]      (indices clipped to n-1)
    so windows depend on input frames only and can be computed in any order / on any GPU."""
    if n_frames <= 0:
        raise ValueError("empty video")
    out = []
    for k in range(num_windows(n_frames)):
        if k == 0:
            idx = list(range(INFER_LEN))
        else:
            idx = [0, STEP * k - OVERLAP] + [STEP * k + 2 + j for j in range(INFER_LEN - 2)]
        out.append([min(i, n_frames - 1) for i in idx])
    return out


def _constrain(x: float, min_val: int) -> int:
    y = int(np.round(x / 14) * 14)
    if y < min_val:
        y = int(np.ceil(x / 14) * 14)
    return y


def get_resize_hw(h0: int, w0: int, input_size: int = 518):
    """Network input size for an (h0, w0) video: aspect guard (video_depth.py:167-171) then Resize.get_size with
    keep_aspect_ratio / lower_bound / multiple of 14 (util/transform.py:62-107)."""
    ratio = max(h0, w0) / min(h0, w0)
    if ratio > 1.78:
        input_size = int(input_size * 1.777 / ratio)
        input_size = round(input_size / 14) * 14
    scale_h, scale_w = input_size / h0, input_size / w0
    if scale_w > scale_h:
        scale_h = scale_w
    else:
        scale_w = scale_h
    return _constrain(scale_h * h0, input_size), _constrain(scale_w * w0, input_size)


def preprocess_frames(frames: np.ndarray, indices: Iterable[int], input_size: int = 518) -> np.ndarray:
    """uint8 [N,H0,W0,3] RGB -> float32 [len(indices),3,nh,nw]: /255, cv2 INTER_CUBIC resize, ImageNet normalise in
    float64, CHW (util/transform.py:109-158 as composed at video_depth.py:173-185,198).  Each source frame is
    processed once (the reference redoes the overlap frames for every window)."""
    import cv2
    indices = list(indices)
    h0, w0 = frames.shape[1:3]
    nh, nw = get_resize_hw(h0, w0, input_size)
    out = np.empty((len(indices), 3, nh, nw), dtype=np.float32)
    for j, i in enumerate(indices):
        img = frames[i].astype(np.float32) / 255.0
        img = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_CUBIC)
        img = (img - _MEAN) / _STD
        out[j] = np.transpose(img, (2, 0, 1))
    return out


def plan_feature_cache(windows, n_slots: int = INFER_LEN):
    """Slot bookkeeping of the encoder-feature cache for a sequence of windows (lists of source frames): frames the
    current window does not use are evicted (a window only ever reuses frames of its predecessor), frames not cached yet
    get free slots.  Returns, per window, (new_frames, their_slots, slot_of_each_window_position).  Pure host logic
    (the device side is video_depth.FeatureCache); at most `n_slots` frames are live because a window has 32 slots."""
    where, free, plan = {}, list(range(n_slots)), []
    for src in windows:
        need = set(src)
        for f in [f for f in where if f not in need]:
            free.append(where.pop(f))
        missing = sorted(need - where.keys())
        if len(missing) > len(free):
            raise ValueError("feature cache too small for this window")
        slots = [free.pop() for _ in missing]
        where.update(zip(missing, slots))
        plan.append((missing, slots, [where[f] for f in src]))
    return plan


# ------------------------------------------------------------------------------------------------
# bounded device buffers of the long-video driver (pure index logic; the device side is video_depth.FrameUploader /
# video_depth.WindowAligner, the property tests are tests/test_parallel_cpu.py::test_*_ring_*)
# ------------------------------------------------------------------------------------------------
def aligner_ring_len(n_frames: int, ring_segs: int = 4) -> int:
    """Length of the WindowAligner's ring of aligned frames: `ring_segs` segments of 22 frames, fewer for short videos
    (never less than the 44 frames window 0 needs behind the 12-frame shift)."""
    seg = INFER_LEN - OVERLAP
    k = -(-n_frames // seg)
    total = k * seg + OVERLAP
    return seg * min(ring_segs, (total + seg - OVERLAP) // seg)


def aligner_ring_pos(a: int, ring_len: int) -> int:
    """Ring position of aligned frame `a`: shifted by 12 so that window 0 (32 frames) ends, and every later window's 22
    new frames start, on a segment boundary -- a window's new frames and its 8-frame cross-fade tail never wrap."""
    return (a + INFER_LEN - 2 * OVERLAP) % ring_len


def upload_ring_plan(n_needed: int, chunk: int = 64, ring_chunks: int = 4):
    """Device slots and upload chunks of the FrameUploader for `n_needed` frames in upload order.  Returns
    (ring, n_slots, slot_of_position, chunks) with chunks = [(lo, hi)] in upload order.  With the ring on, position 0 (source
    frame 0, read by every window) owns slot 0 and is uploaded alone; positions >= 1 share `ring_chunks * chunk` slots and
    are uploaded in chunks aligned to the ring regions, so chunk m replaces chunk m - ring_chunks."""
    cap = chunk * ring_chunks
    ring = n_needed > 1 + cap
    if not ring:
        return False, n_needed, list(range(n_needed)), [(lo, min(lo + chunk, n_needed)) for lo in range(0, n_needed, chunk)]
    slots = [0] + [1 + (j - 1) % cap for j in range(1, n_needed)]
    chunks = [(0, 1)] + [(lo, min(lo + chunk, n_needed)) for lo in range(1, n_needed, chunk)]
    return True, 1 + cap, slots, chunks
