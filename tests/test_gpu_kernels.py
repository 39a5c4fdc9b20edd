"""-m gpu: every C-ABI kernel against the CPU oracle on the same seeded inputs (through ctypes)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "gpu tests need a B200"


def test_simt_kernels_vs_oracle():
    from tools import gpu_check
    assert gpu_check.check_simt()


@pytest.mark.parametrize("B,N,heads,mode,masked", [
    (1, 128, 1, "vones", False), (1, 128, 1, "quniform", False), (1, 128, 1, "random", False),
    (1, 256, 2, "random", False), (2, 300, 8, "random", True), (1, 40, 8, "random", True),
    (8, 1050, 8, "random", True)])
def test_enc_attn_fwd(B, N, heads, mode, masked):
    from tools import gpu_check
    assert gpu_check.attn_case(B, N, heads, mode, masked)


@pytest.mark.parametrize("B,N,heads,masked", [(1, 128, 1, False), (1, 256, 2, False), (2, 300, 8, True),
                                              (3, 54, 8, True), (8, 1050, 8, True)])
def test_enc_attn_bwd(B, N, heads, masked):
    from tools import gpu_check
    assert gpu_check.attn_bwd_case(B, N, heads, masked)
