"""GPU: the handle-level C API (vda_create ... vda_forward, csrc/model.cu) against the Python engine on the same weights
and inputs.  Both issue the same libvda operator calls with the same packed weights, so the outputs must be bit-identical;
the C side additionally never allocates inside vda_forward (it runs inside a CUDA graph capture here)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from e2e_checks import build_model  # noqa: E402
from video_depth_anything_b200 import MODEL_CONFIGS  # noqa: E402
from video_depth_anything_b200.cmodel import CModel  # noqa: E402


@pytest.mark.parametrize("enc,shape,dtype", [("vits", (1, 8, 3, 56, 70), torch.bfloat16),
                                             ("vits", (2, 4, 3, 42, 42), torch.float16),
                                             ("vits", (1, 32, 3, 518, 518), torch.bfloat16),
                                             ("vitl", (1, 4, 3, 518, 518), torch.bfloat16)])
def test_c_forward_is_bit_identical_to_the_python_engine(enc, shape, dtype):
    m, sd = build_model(enc, 0, dtype)
    x = torch.randn(shape, generator=torch.Generator().manual_seed(5)).cuda()
    ref = m.forward(x)
    cm = CModel(**MODEL_CONFIGS[enc], dtype=dtype)
    cm.load_state_dict(sd)
    out = cm.forward(x)
    torch.cuda.synchronize()
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert torch.equal(out, ref), f"max abs diff {(out - ref).abs().max().item():.3e}"
    assert cm.workspace_bytes(*shape[:2], *shape[3:]) > 0
    cm.close()


def test_c_forward_inside_a_cuda_graph():
    m, sd = build_model("vits", 0, torch.bfloat16)
    x = torch.randn(1, 8, 3, 56, 70, generator=torch.Generator().manual_seed(6)).cuda()
    cm = CModel(**MODEL_CONFIGS["vits"], dtype=torch.bfloat16)
    cm.load_state_dict(sd)
    eager = cm.forward(x)                      # also creates the resampled position embedding of this grid
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        captured = cm.forward(x)
    captured.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(captured, eager)
    cm.close()


def test_c_api_errors():
    cm = CModel(**MODEL_CONFIGS["vits"])
    with pytest.raises(Exception):
        cm.load_state_dict({"pretrained.cls_token": torch.zeros(1, 1, 384)})      # strict: everything else is missing
    with pytest.raises(Exception):
        CModel("vitg", 64, [48, 96, 192, 384])
    cm.close()
