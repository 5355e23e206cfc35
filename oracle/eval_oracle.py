"""TEST INFRASTRUCTURE ONLY (imported by tests/ and the golden generator; never by the product path).

CPU restatement of the reference's sequence evaluation (benchmark/eval/eval.py:67-122 `eval_depthcrafter`, with
benchmark/eval/metric.py:3-13 `abs_relative_difference`, :30-41 `rmse_linear`, :68-84 `threshold_percentage` /
`delta1_acc`): masked least-squares scale/shift alignment of the predicted disparity to 1/gt over the whole sequence in
float64, clip, disparity -> depth, clip to the dataset's max depth, then per-frame masked means averaged over the frames
that have valid pixels.  File reading, cropping and the resize of the prediction stay with the caller.  Pinned against
the reference function itself by oracle/make_golden_eval.py -> tests/golden/eval_*.npz."""
from __future__ import annotations

import numpy as np


def align_disparity(infs: np.ndarray, gts: np.ndarray, max_depth: float):
    """eval.py:86-96.  infs float32 [T,H,W] (predicted disparity), gts float [T,H,W] (depth, <= 0 = invalid).
    Returns (scale, shift, valid_mask)."""
    valid = np.logical_and(gts > 1e-3, gts < max_depth)
    gt_disp = 1.0 / (gts[valid].reshape(-1, 1).astype(np.float64) + 1e-8)
    pred = np.clip(infs, a_min=1e-3, a_max=None)[valid].reshape(-1, 1).astype(np.float64)
    A = np.concatenate([pred, np.ones_like(pred)], axis=-1)
    X = np.linalg.lstsq(A, gt_disp, rcond=None)[0]
    return float(X[0, 0]), float(X[1, 0]), valid


def eval_sequence(infs: np.ndarray, gts: np.ndarray, max_depth: float):
    """eval.py:67-122 -> [abs_relative_difference, rmse_linear, delta1_acc] (float64 arithmetic, as the reference's
    tensors are float64)."""
    scale, shift, valid = align_disparity(infs, gts, max_depth)
    aligned = np.clip(scale * np.clip(infs, 1e-3, None).astype(np.float64) + shift, a_min=1e-3, a_max=None)
    pred_depth = np.clip(1.0 / aligned, a_min=1e-3, a_max=max_depth)       # depth2disparity of a positive map
    gt = gts.astype(np.float64)
    n = valid.sum((-1, -2))
    keep = n > 0
    pred_depth, gt, valid, n = pred_depth[keep], gt[keep], valid[keep], n[keep].astype(np.float64)
    absrel = np.where(valid, np.abs(pred_depth - gt) / np.where(valid, gt, 1.0), 0.0).sum((-1, -2)) / n
    mse = np.where(valid, (pred_depth - gt) ** 2, 0.0).sum((-1, -2)) / n
    ratio = np.maximum(pred_depth / np.where(valid, gt, 1.0), np.where(valid, gt, 1.0) / pred_depth)
    d1 = np.where(valid, ratio < 1.25, False).sum((-1, -2)) / n
    return [float(absrel.mean()), float(np.sqrt(mse).mean()), float(d1.mean())]


def tae_pair(depth1: np.ndarray, depth2: np.ndarray, R_2_1: np.ndarray, T_2_1: np.ndarray, K: np.ndarray, mask: np.ndarray):
    """benchmark/eval/eval_tae.py:60-107 `tae_torch` in NumPy float64: un-project depth1, move it to frame 2, project,
    round (half to even, as torch.round), scatter (`depth_proj[valid_Y, valid_X] = valid_Z`: NumPy keeps the LAST
    assignment for repeated indices -- the reference's single-threaded index_put behaves the same), masked AbsRel
    against depth2.  Returns 0 where the reference returns 0."""
    H, W = depth1.shape
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    xx, yy = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    X = (xx - cx) * depth1 / fx
    Y = (yy - cy) * depth1 / fy
    pts = np.stack([X.ravel(), Y.ravel(), depth1.ravel()], axis=1)
    ptw = pts @ R_2_1.T + T_2_1
    with np.errstate(divide="ignore", invalid="ignore"):
        xp = np.rint(ptw[:, 0] * fx / ptw[:, 2] + cx)
        yp = np.rint(ptw[:, 1] * fy / ptw[:, 2] + cy)
    ok = (xp >= 0) & (xp < W) & (yp >= 0) & (yp < H)          # (NaN / inf compare false)
    if ok.sum() == 0:
        return 0.0
    proj = np.zeros((H, W), dtype=np.float64)
    proj[yp[ok].astype(np.int64), xp[ok].astype(np.int64)] = ptw[ok, 2]
    valid = (proj > 0) & (depth2 > 0) & mask
    if valid.sum() == 0:
        return 0.0
    return float(np.mean(np.abs(depth2[valid] - proj[valid]) / depth2[valid]))


def eval_tae(infs: np.ndarray, gts: np.ndarray, Ks: np.ndarray, poses: np.ndarray, max_depth: float, masks=None):
    """benchmark/eval/eval_tae.py:109-213 `eval_TAE` after the file loading (inputs already cropped / resized)."""
    scale, shift, _ = align_disparity(infs, gts, max_depth)
    aligned = np.clip(scale * np.clip(infs, 1e-3, None).astype(np.float64) + shift, a_min=1e-3, a_max=None)
    depth = np.clip(1.0 / aligned, a_min=1e-3, a_max=max_depth)
    T = depth.shape[0]
    total = 0.0
    for i in range(T - 1):
        T_2_1 = np.linalg.inv(poses[i + 1]) @ poses[i]
        m1 = np.ones(depth[i].shape, bool) if masks is None else masks[i] > 0
        m2 = np.ones(depth[i].shape, bool) if masks is None else masks[i + 1] > 0
        total += tae_pair(depth[i], depth[i + 1], T_2_1[:3, :3], T_2_1[:3, 3], Ks[i], m2)
        T_1_2 = np.linalg.inv(T_2_1)
        total += tae_pair(depth[i + 1], depth[i], T_1_2[:3, :3], T_1_2[:3, 3], Ks[i], m1)
    return total / (2 * (T - 1)) * 100.0
