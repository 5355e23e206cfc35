"""Pin oracle/eval_oracle.py against the UNMODIFIED reference evaluation (benchmark/eval/eval.py `eval_depthcrafter`,
imported from /root/reference, build container only): synthetic sequences are written as .npy files, evaluated by the
reference function (its tensors moved to the CPU: `device` patched from 'cuda' to 'cpu'), and the inputs + metrics are
committed as tests/golden/eval_*.npz.

    python oracle/make_golden_eval.py
Shims: `matplotlib` (hard import at eval.py:3, unused on this path) is stubbed."""
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference_eval():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import benchmark.eval.eval as ev           # `from .metric import *`
    import benchmark.eval.metric as metric
    ev.metric = metric                         # eval.py:109 looks the functions up on a module called `metric`
    ev.device = "cpu"
    return ev


def synth_case(seed, T, H, W, max_depth, holes):
    rng = np.random.default_rng(seed)
    gt = (rng.random((T, H, W)) * (max_depth * 1.2) + 0.05).astype(np.float64)
    gt[rng.random((T, H, W)) < holes] = 0.0                    # missing ground truth
    if T > 2:
        gt[1] = 0.0                                            # a frame without any valid pixel
    disp_true = 1.0 / np.maximum(gt, 0.05)
    inf = (0.37 * disp_true + 0.11 + 0.02 * rng.standard_normal((T, H, W))).astype(np.float32)
    inf[rng.random((T, H, W)) < 0.01] = -0.5                   # below the 1e-3 clip
    return inf, gt


def main():
    from oracle import eval_oracle as E
    ev = import_reference_eval()
    man = {}
    for name, (seed, T, H, W, md, holes) in {"eval_T5_24x40": (0, 5, 24, 40, 80.0, 0.2),
                                             "eval_T3_37x53_dense": (1, 3, 37, 53, 10.0, 0.0),
                                             "eval_T9_16x16_sparse": (2, 9, 16, 16, 70.0, 0.9)}.items():
        inf, gt = synth_case(seed, T, H, W, md, holes)
        with tempfile.TemporaryDirectory() as d:
            ip, gp = [], []
            for t in range(T):
                ip.append(os.path.join(d, f"i{t}.npy"))
                gp.append(os.path.join(d, f"g{t}.npy"))
                np.save(ip[-1], inf[t])
                np.save(gp[-1], gt[t])
            args = types.SimpleNamespace(max_eval_len=T, max_depth_eval=md, a=0, b=H, c=0, d=W)
            ref = ev.eval_depthcrafter(ip, gp, [1.0] * T, args)
        mine = E.eval_sequence(inf, np.where(gt == 0, -1.0, gt), md)     # get_gt: zeros become -1 (eval.py:47)
        err = max(abs(a - b) for a, b in zip(ref, mine))
        print(name, ref, mine, err)
        assert err < 1e-7, err
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), inf=inf, gt=np.where(gt == 0, -1.0, gt),
                            metrics=np.array(ref, dtype=np.float64))
        man[name] = {"max_depth": md, "metrics": ref, "oracle_vs_ref_max_abs": err}
    json.dump({"generator": "oracle/make_golden_eval.py", "cases": man}, open(os.path.join(GOLD, "EVAL_MANIFEST.json"), "w"),
              indent=1)


if __name__ == "__main__":
    main()
