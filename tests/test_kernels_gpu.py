"""GPU parity: each libvda kernel vs the plain PyTorch fp32 op it replaces (tests/kernel_checks.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from kernel_checks import CHECKS  # noqa: E402


@pytest.mark.parametrize("name,fn", CHECKS, ids=[c[0] for c in CHECKS])
def test_kernel(name, fn):
    assert torch.cuda.is_available()
    err, tol = fn()
    assert err == err and err <= tol, f"{name}: max abs err {err:.3e} > tol {tol:.3e}"
