"""CPU: libvda.so builds, loads, and exports every symbol include/vda.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from video_depth_anything_b200 import build
    path = build.build()
    return ctypes.CDLL(path)


def _header_names():
    hdr = open(os.path.join(ROOT, "include", "vda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(vda_[a-z0-9_]+)\s*\(", hdr))


def test_header_symbols_exported(lib):
    names = _header_names()
    assert len(names) >= 15
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/vda.h but not exported by libvda.so"


def test_binding_matches_header(lib):
    from video_depth_anything_b200 import _lib
    assert _header_names() == set(_lib.EXPORTS)
    assert _lib.load().vda_version() == 100


def test_gemm_params_struct_layout():
    """ctypes mirror of vda_gemm_params must have the C layout (checked against a tiny C probe)."""
    import subprocess
    import tempfile
    from video_depth_anything_b200._lib import GemmParams
    src = ('#include "vda.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu", sizeof(vda_gemm_params), '
           '__builtin_offsetof(vda_gemm_params, out), __builtin_offsetof(vda_gemm_params, tail_b));'
           'printf(" %zu %zu", __builtin_offsetof(vda_gemm_params, out16), __builtin_offsetof(vda_gemm_params, a_k));return 0;}')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o",
                               os.path.join(d, "p")])
        size, off_out, off_tb, off_o16, off_ak = map(int, subprocess.check_output([os.path.join(d, "p")]).split())
    assert ctypes.sizeof(GemmParams) == size
    assert GemmParams.out.offset == off_out and GemmParams.tail_b.offset == off_tb
    assert GemmParams.out16.offset == off_o16 and GemmParams.a_k.offset == off_ak


def test_no_cpu_fallback():
    import numpy as np
    import torch
    from video_depth_anything_b200 import MODEL_CONFIGS, VideoDepthAnything
    m = VideoDepthAnything(**MODEL_CONFIGS["vits"])
    with pytest.raises(RuntimeError):
        m.forward(torch.zeros(1, 2, 3, 28, 28))
    with pytest.raises(RuntimeError):
        m.infer_video_depth(np.zeros((3, 28, 28, 3), np.uint8), 24, device="cpu")


def test_state_dict_contract():
    import torch
    from video_depth_anything_b200 import MODEL_CONFIGS, VideoDepthAnything, synth_state_dict
    m = VideoDepthAnything(**MODEL_CONFIGS["vits"])
    sd = synth_state_dict(**MODEL_CONFIGS["vits"], seed=3)
    assert len(sd) == 351
    m.load_state_dict(sd, strict=True)
    assert torch.equal(m.state_dict()["head.projects.0.weight"], sd["head.projects.0.weight"])
    bad = dict(sd)
    bad.pop("pretrained.norm.weight")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)


def test_windows_match_oracle():
    from oracle import vda_oracle as O
    from video_depth_anything_b200 import windows as Wn
    for n in (1, 5, 22, 23, 33, 50, 131, 2048):
        assert Wn.window_source_indices(n) == O.window_source_indices_literal(n)
    for hw in ((518, 518), (720, 1280), (60, 80), (1080, 1920), (480, 2000), (2000, 480)):
        assert Wn.get_resize_hw(*hw, 518) == O.get_resize_hw(*hw, 518)


def test_feature_cache_plan_properties():
    """Host logic of the encoder-feature reuse (windows.plan_feature_cache): replaying the plan on a fake slot store gives
    every window position the features of its own source frame; contiguous window runs encode every frame exactly once;
    arbitrary window subsets (multi-GPU blocks, non-contiguous ids) stay correct; never more than 32 slots."""
    from video_depth_anything_b200 import windows as Wn
    for n in (1, 5, 31, 32, 33, 70, 131, 2048):
        wins = Wn.window_source_indices(n)
        for ids in (list(range(len(wins))), list(range(len(wins) // 2, len(wins))), list(range(0, len(wins), 3))):
            if not ids:
                continue
            seq = [wins[k] for k in ids]
            store, encoded = {}, []
            for (missing, slots, positions), src in zip(Wn.plan_feature_cache(seq), seq):
                assert len(set(slots)) == len(slots) and all(0 <= s < Wn.INFER_LEN for s in slots)
                for f, s in zip(missing, slots):
                    store[s] = f                      # "encode frame f into slot s"
                encoded += missing
                assert [store[s] for s in positions] == list(src)
            if ids == list(range(ids[0], ids[0] + len(ids))):          # contiguous run: nothing is encoded twice
                assert len(encoded) == len(set(encoded)) == len({f for w in seq for f in w})
    steady = Wn.plan_feature_cache(Wn.window_source_indices(200))
    assert [len(m) for m, _, _ in steady][:4] == [32, 22, 22, 22]
