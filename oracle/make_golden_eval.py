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


def import_reference_tae():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import torch
    torch.set_num_threads(1)                   # index_put with repeated indices: sequential = last writer wins
    import benchmark.eval.eval_tae as tae
    tae.device = torch.device("cpu")
    return tae


def synth_tae_case(seed, T, H, W, max_depth, motion, masked):
    """A smooth positive depth sequence seen by a slowly moving pinhole camera (poses camera-to-world)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    gts, poses, Ks = [], [], []
    for t in range(T):
        gt = 2.5 + 1.2 * np.sin(xx / 9.0 + 0.3 * t) * np.cos(yy / 7.0) + 0.05 * rng.standard_normal((H, W))
        gt[rng.random((H, W)) < 0.05] = 0.0                       # holes in the ground truth
        gts.append(gt)
        ang = motion * 0.03 * t + 0.01 * rng.standard_normal(3)
        Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
        Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
        Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
        P = np.eye(4)
        P[:3, :3] = Rz @ Ry @ Rx
        P[:3, 3] = motion * np.array([0.05 * t, -0.02 * t, 0.04 * t]) + 0.01 * rng.standard_normal(3)
        poses.append(P)
        Ks.append(np.array([[0.9 * W, 0, W / 2.0 - 0.5], [0, 0.9 * W, H / 2.0 - 0.5], [0, 0, 1.0]]))
    gts = np.stack(gts)
    inf = (0.8 / np.maximum(gts, 0.3) + 0.07 + 0.01 * rng.standard_normal(gts.shape)).astype(np.float32)
    masks = (rng.random(gts.shape) > 0.3) if masked else None
    return inf, gts, np.stack(Ks), np.stack(poses), masks


def make_tae_goldens(man):
    import cv2
    from oracle import eval_oracle as E
    tae = import_reference_tae()
    for name, (seed, T, H, W, md, motion, masked) in {"tae_T4_48x64": (10, 4, 48, 64, 10.0, 1.0, False),
                                                       "tae_T3_37x53_fast": (11, 3, 37, 53, 10.0, 6.0, False),
                                                       "tae_T5_40x40_masked": (12, 5, 40, 40, 4.0, 2.0, True)}.items():
        inf, gt, Ks, poses, masks = synth_tae_case(seed, T, H, W, md, motion, masked)
        with tempfile.TemporaryDirectory() as d:
            ip, gp, mp_ = [], [], []
            for t in range(T):
                ip.append(os.path.join(d, f"i{t}.npy"))
                gp.append(os.path.join(d, f"g{t}.npy"))
                np.save(ip[-1], inf[t])
                np.save(gp[-1], gt[t])
                if masked:
                    mp_.append(os.path.join(d, f"m{t}.png"))
                    cv2.imwrite(mp_[-1], (masks[t] * 255).astype(np.uint8))
            args = types.SimpleNamespace(max_depth_eval=md, a=0, b=H, c=0, d=W, mask=masked, hard_crop=False)
            ref = float(tae.eval_TAE(ip, gp, [1.0] * T, mp_, list(Ks), list(poses), args))
        mine = E.eval_tae(inf, gt, Ks, poses, md, masks)
        err = abs(ref - mine)
        print(name, ref, mine, err)
        assert err < 1e-9 * max(1.0, abs(ref)), err
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), inf=inf, gt=gt, Ks=Ks, poses=poses,
                            masks=(masks if masked else np.zeros(0, bool)), tae=np.array(ref, dtype=np.float64))
        man[name] = {"max_depth": md, "tae": ref, "masked": masked, "oracle_vs_ref_abs": err}


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
    tman = {}
    make_tae_goldens(tman)
    json.dump({"generator": "oracle/make_golden_eval.py", "cases": man, "tae_cases": tman},
              open(os.path.join(GOLD, "EVAL_MANIFEST.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
