"""A compiled C program as the host of libvda's handle-level API (tests/c_host/host_forward.c): built with gcc against
include/vda.h and the in-tree libvda.so.  CPU: it compiles, links and reports the library version (no GPU needed).  GPU:
its depth map for a synthetic state dict and window is bit-identical to the Python engine's."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_host", "host_forward.c")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    from video_depth_anything_b200 import build
    lib = build.build()
    exe = str(tmp_path / "host_forward")
    libdir = os.path.dirname(lib)
    subprocess.check_call(["gcc", "-O1", "-std=c99", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), SRC,
                           "-o", exe, "-L", libdir, "-l:libvda.so", "-L", os.path.join(CUDA, "lib64"), "-lcudart",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(CUDA, "lib64")])
    return exe


def test_c_host_builds_and_links(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.check_output([exe, "--version"]).decode().strip()
    assert out == "100"


@pytest.mark.gpu
@pytest.mark.parametrize("enc,shape", [("vits", (1, 8, 56, 70)), ("vits", (1, 32, 518, 518))])
def test_c_host_forward_is_bit_identical(tmp_path, enc, shape):
    import torch
    from e2e_checks import build_model
    from video_depth_anything_b200 import MODEL_CONFIGS
    exe = _build(tmp_path)
    m, sd = build_model(enc, 0, torch.bfloat16)
    B, T, H, W = shape
    x = torch.randn(B, T, 3, H, W, generator=torch.Generator().manual_seed(11))
    ref = m.forward(x.cuda()).cpu().numpy()
    wpath, ipath, opath = (str(tmp_path / n) for n in ("weights.bin", "input.bin", "output.bin"))
    with open(wpath, "wb") as f:
        f.write(struct.pack("<i", len(sd)))
        for k, v in sd.items():
            t = v.detach().to("cpu", torch.float32).contiguous().numpy()
            name = k.encode()
            f.write(struct.pack("<i", len(name)) + name + struct.pack("<i", t.ndim) + struct.pack(f"<{t.ndim}q", *t.shape))
            f.write(t.tobytes())
    with open(ipath, "wb") as f:
        f.write(struct.pack("<4i", B, T, H, W))
        f.write(x.numpy().tobytes())
    cfg = MODEL_CONFIGS[enc]
    args = [exe, enc, str(cfg["features"]), *map(str, cfg["out_channels"]), "0", wpath, ipath, opath]
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    out = np.fromfile(opath, dtype=np.float32).reshape(ref.shape)
    assert np.array_equal(out, ref), f"max abs diff {np.abs(out - ref).max():.3e}"
