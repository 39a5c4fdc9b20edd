"""-m gpu: in-kernel dropout.  Every kernel recomputes its keep-mask from (seed, site, row, col) with a counter-based
hash; oracle/dropout_mask.py is the numpy twin, so the tests apply THE SAME mask with plain torch arithmetic
(torch.nn.functional.dropout semantics: zero with probability p, scale the rest by 1/(1-p)) and compare."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
P = 0.3


def _mask(seed, site, M, C, rows=None):
    from oracle.dropout_mask import keep_mask, thr16_of, scale_of
    t = thr16_of(P)
    rows = np.arange(M) if rows is None else rows
    return torch.from_numpy(keep_mask(seed, site, rows, np.arange(C), t)).float() * scale_of(t)


def _drop(seed, site):
    from object_detection_destr_b200 import ops
    return (torch.tensor([seed], dtype=torch.int32, device="cuda"), ops.drop_thr16(P), site)


def test_mask_statistics_and_independence():
    m1, m2, m3 = _mask(1, 7, 4000, 512) > 0, _mask(2, 7, 4000, 512) > 0, _mask(1, 8, 4000, 512) > 0
    assert abs(float(m1.float().mean()) - (1 - P)) < 2e-3
    for other in (m2, m3):  # different seed / site -> independent masks: agreement = (1-p)^2 + p^2
        assert abs(float((m1 == other).float().mean()) - ((1 - P) ** 2 + P ** 2)) < 3e-3


def test_dropout_inplace_matches_numpy_twin():
    from object_detection_destr_b200 import ops
    x = torch.randn(801, 1024, generator=torch.Generator().manual_seed(0)).bfloat16()
    y = ops.dropout_inplace(x.clone().cuda(), _drop(11, 3)).cpu().float()
    ref = (x.float() * _mask(11, 3, 801, 1024)).bfloat16().float()
    assert torch.equal(y, ref)


@pytest.mark.parametrize("M,D", [(8400, 256), (803, 512)])
def test_add_layernorm_dropout_fwd_bwd(M, D):
    from object_detection_destr_b200 import ops
    g = torch.Generator().manual_seed(M)
    a, b, dy = (torch.randn(M, D, generator=g).bfloat16() for _ in range(3))
    gam, bet = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    res = torch.randn(M, D, generator=g).bfloat16()
    mk = _mask(5, 9, M, D)
    af, bf = a.float().requires_grad_(), b.float().requires_grad_()
    gr, br = gam.clone().requires_grad_(), bet.clone().requires_grad_()
    y_ref = torch.nn.functional.layer_norm(af + bf * mk, (D,), gr, br, 1e-5)
    y_ref.backward(dy.float())
    drop = _drop(5, 9)
    y, mean, rstd = ops.add_layernorm(a.cuda(), b.cuda(), gam.cuda(), bet.cuda(), save_stats=True, drop=drop)
    assert float((y.cpu().float() - y_ref).abs().max()) < 3e-2
    dbias = torch.zeros(D, device="cuda")
    dx, dg, db, ro = ops.add_layernorm_bwd(dy.cuda(), a.cuda(), b.cuda(), gam.cuda(), mean, rstd, dbias=dbias,
                                           res_in=res.cuda(), drop=drop)
    assert float((dx.cpu().float() - bf.grad).abs().max()) < 3e-2            # gradient w.r.t. b: masked
    assert float((ro.cpu().float() - (af.grad + res.float())).abs().max()) < 5e-2  # residual stream: not masked
    assert torch.allclose(dbias.cpu(), bf.grad.sum(0), rtol=2e-2, atol=0.3)
    assert torch.allclose(dg.cpu(), gr.grad, rtol=2e-2, atol=0.3) and torch.allclose(db.cpu(), br.grad, rtol=2e-2, atol=0.3)
    # un-masked sum without a residual input
    dx2, _, _, ro2 = ops.add_layernorm_bwd(dy.cuda(), a.cuda(), b.cuda(), gam.cuda(), mean, rstd, drop=drop, want_sum=True)
    assert float((ro2.cpu().float() - af.grad).abs().max()) < 3e-2 and torch.equal(dx2, dx)


def test_relu_bwd_scale():
    from object_detection_destr_b200 import ops
    g = torch.Generator().manual_seed(3)
    dy = torch.randn(800, 1024, generator=g).bfloat16()
    h = (torch.randn(800, 1024, generator=g).clamp(min=0) * _mask(4, 2, 800, 1024)).bfloat16()
    dbias = torch.zeros(1024, device="cuda")
    from oracle.dropout_mask import scale_of, thr16_of
    s = scale_of(thr16_of(P))
    dpre = ops.relu_bwd_colsum(dy.cuda(), h.cuda(), dbias, scale=s).cpu().float()
    ref = (dy.float() * (h.float() > 0) * s)
    assert float((dpre - ref.bfloat16().float()).abs().max()) == 0.0
    assert torch.allclose(dbias.cpu(), ref.sum(0), rtol=1e-3, atol=0.05)


@pytest.mark.parametrize("B,N,heads", [(2, 300, 8), (8, 600, 8), (1, 130, 2)])
def test_encoder_attention_dropout_fwd_bwd(B, N, heads):
    """Attention-probability dropout inside the flash kernels == softmax -> mask -> scale -> PV with the twin mask
    (nn.MultiheadAttention(dropout=p) arithmetic, torch functional.py:6645), forward and all three gradients."""
    import math
    from object_detection_destr_b200 import ops
    g = torch.Generator().manual_seed(B * N)
    C = heads * 32
    qk = torch.randn(B * N, 2 * C, generator=g).bfloat16()
    v = torch.randn(B * N, C, generator=g).bfloat16()
    dout = torch.randn(B * N, C, generator=g).bfloat16()
    kpm = torch.zeros(B, N, dtype=torch.bool)
    kpm[B - 1, N - 37:] = True
    site, seed = 21, 77
    mk = _mask(seed, site, B * heads * N, N).view(B, heads, N, N)
    scale = 1 / math.sqrt(32)
    split = lambda t: t.reshape(B, N, heads, 32).transpose(1, 2)
    qf, kf, vf = (t.float().requires_grad_() for t in (qk[:, :C], qk[:, C:], v))
    s = torch.einsum("bhqd,bhkd->bhqk", split(qf), split(kf)) * scale
    s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    ref = torch.einsum("bhqk,bhkd->bhqd", torch.softmax(s, -1) * mk, split(vf)).transpose(1, 2).reshape(B * N, C)
    ref.backward(dout.float())
    drop = _drop(seed, site)
    bits = ops.pack_key_mask(kpm.cuda(), B, N)
    qk_d, v_d = qk.cuda(), v.cuda()
    out, lse = ops.enc_attn_fwd(qk_d[:, :C], qk_d[:, C:], v_d, bits, B, N, heads, scale, drop=drop)
    assert float((out.cpu().float() - ref).abs().max()) < 3e-2, float((out.cpu().float() - ref).abs().max())
    dqk, dv = ops.enc_attn_bwd(qk_d[:, :C], qk_d[:, C:], v_d, bits, out, dout.cuda(), lse, B, N, heads, scale, drop=drop)
    for got, exp, name in ((dv, vf.grad, "dV"), (dqk[:, C:], kf.grad, "dK"), (dqk[:, :C], qf.grad, "dQ")):
        err = float((got.cpu().float() - exp).abs().max())
        assert err < 3e-2 * float(exp.abs().max()) + 2e-2, (name, err, float(exp.abs().max()))
    # and it really dropped something: the no-dropout output differs
    out0, _ = ops.enc_attn_fwd(qk_d[:, :C], qk_d[:, C:], v_d, bits, B, N, heads, scale)
    assert float((out0.float() - out.float()).abs().max()) > 0.05
