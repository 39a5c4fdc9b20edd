"""-m gpu: destr_select_queries (csrc/query_select.cu) against the reference's golden vectors and the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import query_select_oracle as QO
from test_query_select_oracle import cases, check_against_reference

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("i", range(5))
def test_kernel_matches_reference_golden(i):
    from object_detection_destr_b200 import query_select as QS
    c = cases()[i]
    k = int(c["k"])
    B, N, _ = c["scores"].shape
    t = lambda a: torch.from_numpy(a).cuda()
    if "sel" in c:
        sel, cen, idx, status = QS.select_queries(t(c["scores"]), t(c["mask"]), t(c["cls_feat"]), t(c["reg_feat"]),
                                                  t(c["coords"]), top_k=10 ** 6, valid0=k)
        QS.check_status(status)
        assert np.array_equal(sel.cpu().numpy(), c["sel"]) and np.array_equal(cen.cpu().numpy(), c["cen"])
    else:
        bi, flat = QS.get_topk_index(t(c["scores"]), k, t(c["mask"]))
        idx = flat.view(B, k)
        assert bi.tolist() == [b for b in range(B) for _ in range(k)]
    idx = idx.cpu().numpy()
    check_against_reference(c, idx)
    assert np.array_equal(idx, QO.topk_index(c["scores"], k, c["mask"]))  # and bit-exact against the oracle, ties included


@pytest.mark.parametrize("B,N,C,k,pad", [(8, 1050, 91, 100, 37), (2, 4200, 91, 300, 1000), (3, 64, 1, 64, 20)])
def test_kernel_matches_oracle(B, N, C, k, pad):
    """Config-2 / config-4 shapes with the real feature width (256), heavy ties (quantised scores) included."""
    from object_detection_destr_b200 import query_select as QS
    g = torch.Generator().manual_seed(N + C)
    scores = (torch.rand(B, N, C, generator=g) * 64).floor() / 64  # many exact ties
    mask = torch.zeros(B, N, dtype=torch.bool)
    for b in range(1, B):
        mask[b, N - pad * b:] = True
    scores = scores.masked_fill(mask.unsqueeze(-1), 0.0)
    cf, rf = torch.randn(B, N, 256, generator=g), torch.randn(B, N, 256, generator=g)
    co = torch.rand(B, N, 4, generator=g)
    ei, es, ec = QO.select_queries(scores.numpy(), mask.numpy(), cf.numpy(), rf.numpy(), co.numpy(), k)
    sel, cen, idx, _ = QS.select_queries(scores.cuda(), mask.cuda(), cf.cuda(), rf.cuda(), co.cuda(), top_k=k)
    assert np.array_equal(idx.cpu().numpy(), ei)
    assert np.array_equal(sel.cpu().numpy(), es) and np.array_equal(cen.cpu().numpy(), ec)
    sel16 = QS.select_queries(scores.cuda(), mask.cuda(), cf.cuda(), rf.cuda(), co.cuda(), top_k=k, want_bf16=True).selected_objects
    assert torch.equal(sel16.cpu(), torch.from_numpy(es).bfloat16())


def test_no_valid_position_is_reported():
    from object_detection_destr_b200 import query_select as QS
    scores = torch.zeros(2, 32, 3).cuda()
    mask = torch.zeros(2, 32, dtype=torch.bool)
    mask[1] = True
    z = torch.zeros(2, 32, 8).cuda()
    r = QS.select_queries(scores, mask.cuda(), z, z, torch.zeros(2, 32, 4).cuda(), top_k=5)
    assert r.status.tolist() == [0, 1]
    with pytest.raises(ZeroDivisionError):
        QS.check_status(r.status)
    with pytest.raises(ZeroDivisionError):
        QS.get_topk_index(scores, 5, mask.cuda())
