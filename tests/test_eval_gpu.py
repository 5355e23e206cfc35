"""GPU: the device evaluation (vda_eval_sequence through video_depth_anything_b200.evaluate) against the reference's own
metrics (tests/golden/eval_*.npz, produced by benchmark/eval/eval.py) and against the oracle restatement on larger
seeded sequences.  Integer-like work (counts) must agree exactly; the float64 sums to 1e-9 relative."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAN = json.load(open(os.path.join(GOLD, "EVAL_MANIFEST.json")))["cases"]
TOL = 1e-9


def _close(a, b):
    return all(abs(x - y) <= TOL * max(1.0, abs(y)) for x, y in zip(a, b))


@pytest.mark.parametrize("name", sorted(MAN))
def test_eval_vs_reference_golden(name):
    from video_depth_anything_b200.evaluate import eval_sequence
    g = np.load(os.path.join(GOLD, name + ".npz"))
    got = eval_sequence(g["inf"], g["gt"], MAN[name]["max_depth"])
    ref = g["metrics"].tolist()
    assert abs(got[0] - ref[0]) <= 1e-9 and abs(got[1] - ref[1]) <= 1e-8, (got, ref)
    assert abs(got[2] - ref[2]) <= 1e-7, (got, ref)          # the reference rounds delta1 through float32


@pytest.mark.parametrize("T,H,W,md,gt_dtype", [(12, 375, 1242, 80.0, np.float64), (7, 120, 160, 10.0, np.float32),
                                               (3, 5, 7, 70.0, np.float64)])
def test_eval_vs_oracle_seeded(T, H, W, md, gt_dtype):
    from oracle import eval_oracle as E
    from video_depth_anything_b200.evaluate import eval_sequence
    rng = np.random.default_rng(T * 1000 + H)
    gt = (rng.random((T, H, W)) * md * 1.1 + 0.01).astype(gt_dtype)
    gt[rng.random((T, H, W)) < 0.3] = -1
    inf = (0.8 / np.maximum(gt, 0.01) + 0.05 + 0.01 * rng.standard_normal((T, H, W))).astype(np.float32)
    got, ss = eval_sequence(inf, gt, md, return_alignment=True)
    ref = E.eval_sequence(inf, gt, md)
    s, t, _ = E.align_disparity(inf, gt, md)
    assert abs(ss[0] - s) <= 1e-9 * abs(s) and abs(ss[1] - t) <= 1e-9 * max(abs(t), 1e-3), (ss, s, t)
    assert _close(got, ref), (got, ref)
    assert eval_sequence(inf, gt, md) == got                 # bit-reproducible


def test_eval_edge_cases():
    from video_depth_anything_b200.evaluate import eval_sequence
    gt = np.full((2, 8, 8), -1.0)
    assert eval_sequence(np.ones((2, 8, 8), np.float32), gt, 10.0) == [0.0, 0.0, 0.0]      # nothing valid
    with pytest.raises(ValueError):
        eval_sequence(np.ones((2, 8, 8), np.float32), np.ones((2, 8, 9)), 10.0)
    with pytest.raises(RuntimeError):
        eval_sequence(np.ones((2, 8, 8), np.float32), gt, 10.0, device="cpu")


# ------------------------------------------------------------------------------------------------ temporal alignment error
TAE = json.load(open(os.path.join(GOLD, "EVAL_MANIFEST.json"))).get("tae_cases", {})


@pytest.mark.parametrize("name", sorted(TAE))
def test_tae_vs_reference_golden(name):
    """Device TAE (vda_eval_tae: exact integer last-writer scatter + float64 gather) against the number the reference's
    own eval_TAE / tae_torch produced (benchmark/eval/eval_tae.py, run single-threaded on the CPU by
    oracle/make_golden_eval.py)."""
    from video_depth_anything_b200.evaluate import eval_tae
    g = np.load(os.path.join(GOLD, name + ".npz"))
    masks = g["masks"] if TAE[name]["masked"] else None
    got = eval_tae(g["inf"], g["gt"], g["Ks"], g["poses"], TAE[name]["max_depth"], masks)
    ref = float(g["tae"])
    assert abs(got - ref) <= 1e-9 * max(1.0, abs(ref)), (got, ref)
    assert eval_tae(g["inf"], g["gt"], g["Ks"], g["poses"], TAE[name]["max_depth"], masks) == got      # bit-reproducible


@pytest.mark.parametrize("T,H,W,motion,masked", [(6, 240, 320, 2.0, False), (4, 464, 618, 1.0, True), (3, 33, 47, 8.0, False)])
def test_tae_vs_oracle_seeded(T, H, W, motion, masked):
    """Larger seeded sequences (ScanNet's cropped 464 x 618 among them) against the oracle restatement; several jobs per
    launch and a chunked job list (jobs_per_call=3) must give the same number."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    from make_golden_eval import synth_tae_case
    from oracle import eval_oracle as E
    from video_depth_anything_b200.evaluate import eval_tae
    inf, gt, Ks, poses, masks = synth_tae_case(100 + T, T, H, W, 10.0, motion, masked)
    ref = E.eval_tae(inf, gt, Ks, poses, 10.0, masks)
    got = eval_tae(inf, gt, Ks, poses, 10.0, masks)
    assert abs(got - ref) <= 1e-9 * max(1.0, abs(ref)), (got, ref)
    assert eval_tae(inf, gt, Ks, poses, 10.0, masks, jobs_per_call=3) == got


def test_tae_edge_cases():
    from video_depth_anything_b200.evaluate import eval_tae
    K = np.array([[50.0, 0, 16], [0, 50.0, 12], [0, 0, 1]])
    Ks, eye = np.stack([K] * 3), np.stack([np.eye(4)] * 3)
    inf = np.full((3, 24, 32), 0.5, np.float32)
    gt = np.full((3, 24, 32), 2.0)
    assert eval_tae(inf, gt, Ks, eye, 10.0) == 0.0                      # static camera, constant depth: perfect consistency
    far = eye.copy()
    far[1, :3, 3] = 1e6                                                  # nothing re-projects into the frame: tae_torch returns 0
    assert eval_tae(inf, gt, Ks, far, 10.0) == 0.0
    with pytest.raises(ValueError):
        eval_tae(inf[:1], gt[:1], Ks[:1], eye[:1], 10.0)
    with pytest.raises(RuntimeError):
        eval_tae(inf, gt, Ks, eye, 10.0, device="cpu")
