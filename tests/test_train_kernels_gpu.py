"""GPU parity of the training-only kernels against torch fp32 on the same fp16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, tol=2e-3):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, f"max err {err:.4e} vs scale {scale:.4e}"


def _rand(*shape, s=1.0):
    return (torch.randn(*shape, device="cuda") * s).half()


@pytest.mark.parametrize("M,P,Q", [(4096, 320, 128), (1000, 128, 320), (16384, 1280, 128), (300, 64, 64), (2048, 640, 640),
                                   (8192, 128, 2560), (77 * 4, 8, 768)])
def test_wgrad_tn(M, P, Q):
    """dW = A^T B over the token dim (LoRA up/down grads, zero-conv grads): MN-major UMMA operands."""
    from ctrlora_b200 import ops
    torch.manual_seed(0)
    a, b = _rand(M, P), _rand(M, Q, s=M ** -0.5)
    out = ops.wgrad_tn(a, b)
    ref = a.float().t() @ b.float()
    _close(out, ref, 1e-3)
    out2 = ops.wgrad_tn(a, b, out=out.clone(), alpha=0.5, beta=2.0)
    _close(out2, 0.5 * ref + 2.0 * out, 1e-3)
