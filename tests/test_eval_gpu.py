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
