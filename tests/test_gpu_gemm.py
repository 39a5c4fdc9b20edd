"""-m gpu: the tcgen05 GEMM family (csrc/gemm_tc.cu) against plain fp32 torch on the same bf16 inputs.
Tolerance: bf16 output rounding (2^-8 relative) on top of fp32 accumulation -> |err| <= 1e-2 * absmax(ref) + 1e-2
for bf16 outputs, 2e-3 relative for the fp32 weight gradient; LayerNorm statistics 1e-4."""
import numpy as np
import pytest
import torch

from parity_log import record

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _close(got, ref, what, rel=1e-2, ab=1e-2):
    err = float((got.float() - ref.float()).abs().max())
    tol = rel * float(ref.float().abs().max()) + ab
    assert err <= tol, (what, err, tol)
    return err


def _mask(seed, site, M, C):
    from oracle.dropout_mask import keep_mask, scale_of, thr16_of
    t = thr16_of(0.3)
    return torch.from_numpy(keep_mask(seed, site, np.arange(M), np.arange(C), t)).float().cuda() * scale_of(t)


@pytest.mark.parametrize("M,N,K,b_kn", [(8400, 256, 256, False), (8400, 512, 256, False), (8400, 2048, 256, False),
                                        (8400, 256, 2048, False), (8400, 256, 2048, True), (8400, 2048, 256, True),
                                        (8400, 256, 512, True), (800, 1536, 512, False), (37, 64, 72, False),
                                        (131, 96, 200, True), (8400, 3072, 256, False)])
def test_gemm_store_plain(M, N, K, b_kn):
    from object_detection_destr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g, device="cuda").to(BF)
    w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).to(BF)   # nn.Linear layout [N, K]
    bias = torch.randn(N, generator=g, device="cuda")
    ref = a.float() @ w.float().t()
    out = ops.gemm(a, w.t().contiguous() if b_kn else w, b_kn=b_kn)
    e = _close(out, ref, "plain")
    record(f"gemm_M{M}_N{N}_K{K}_{'KN' if b_kn else 'NK'}", "plain.max_abs", e)
    out = ops.gemm(a, w.t().contiguous() if b_kn else w, b_kn=b_kn, bias=bias, relu=True)
    _close(out, torch.relu(ref + bias), "bias+relu")


def test_gemm_store_strided_operands_and_epilogue_chain():
    """x + pos * (h W^T + b) (encoder_block.py:38,95), beta = 1 accumulation in place, two outputs, dropout."""
    from object_detection_destr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 1050 * 3 + 7, 256, 256
    big = torch.randn(M, 3 * K, generator=g, device="cuda").to(BF)
    a = big[:, K:2 * K]                                            # strided A (row pitch 768)
    wfull = (torch.randn(3 * N, K, generator=g, device="cuda") / 16).to(BF)
    w = wfull[N:2 * N]                                             # a slice of a packed weight
    bias = torch.randn(N, generator=g, device="cuda")
    x = torch.randn(M, N, generator=g, device="cuda").to(BF)
    pos = torch.randn(M, N, generator=g, device="cuda").to(BF)
    acc = a.float() @ w.float().t() + bias
    out = ops.gemm(a, w, bias=bias, mul=pos, add=x)
    _close(out, x.float() + pos.float() * acc, "x + pos*s")
    # backward pair: ds = dxq * pos, dx_out = dx_in + dxq   (dxq = dqk W, W as [K,N])
    wkn = (torch.randn(K, N, generator=g, device="cuda") / 16).to(BF)
    dxq = a.float() @ wkn.float()
    ds, dx = ops.gemm(a, wkn, b_kn=True, mul=pos, add2=x, out2=True)
    _close(ds, dxq * pos.float(), "ds")
    _close(dx, dxq + x.float(), "dx")
    # in-place accumulation: out aliases add
    acc_buf = x.clone()
    ops.gemm(a, wkn, b_kn=True, add=acc_buf, out=acc_buf)
    _close(acc_buf, dxq + x.float(), "in-place beta=1")
    # dropout in the epilogue = the mask of dropout_inplace
    seed = torch.tensor([77], dtype=torch.int32, device="cuda")
    out = ops.gemm(a, w, bias=bias, relu=True, drop=(seed, ops.drop_thr16(0.3), 9))
    _close(out, torch.relu(acc) * _mask(77, 9, M, N), "dropout")


@pytest.mark.parametrize("M,N,K", [(8400, 2048, 256), (800, 1024, 256), (333, 256, 256)])
def test_gemm_relu_bwd(M, N, K):
    from object_detection_destr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M)
    dy = torch.randn(M, K, generator=g, device="cuda").to(BF)
    w = (torch.randn(K, N, generator=g, device="cuda") / 16).to(BF)   # fc2.weight [256, 2048]
    h = torch.relu(torch.randn(M, N, generator=g, device="cuda")).to(BF)
    h = h * (torch.rand(M, N, generator=g, device="cuda") > 0.3)       # dropped activation: zeros carry both masks
    dbias = torch.full((N,), 3.0, device="cuda")
    scale = 1.0 / 0.7
    dpre = ops.gemm_relu_bwd(dy, w, h.to(BF), scale, dbias)
    ref = (dy.float() @ w.float()) * (h > 0).float() * scale
    _close(dpre, ref, "dpre")
    cs = ref.to(BF).float().sum(0) + 3.0
    err = float((dbias - cs).abs().max())
    record(f"gemm_relu_bwd_M{M}_N{N}", "dbias.max_abs", err, 2e-3 * float(cs.abs().max()) + 0.05)
    assert err <= 2e-3 * float(cs.abs().max()) + 0.05, err


@pytest.mark.parametrize("M,K,double,p", [(8400, 256, False, 0.0), (8400, 2048, True, 0.0), (8400, 2048, True, 0.3),
                                          (777, 256, False, 0.3), (8400, 256, True, 0.0)])
def test_gemm_res_ln(M, K, double, p):
    from object_detection_destr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = torch.randn(M, K, generator=g, device="cuda").to(BF)
    w = (torch.randn(256, K, generator=g, device="cuda") / K ** 0.5).to(BF)
    bias = torch.randn(256, generator=g, device="cuda") * 0.1
    res = torch.randn(M, 256, generator=g, device="cuda").to(BF)
    res2 = torch.randn(M, 256, generator=g, device="cuda").to(BF)
    gam, bet = 1 + 0.1 * torch.randn(256, generator=g, device="cuda"), 0.1 * torch.randn(256, generator=g, device="cuda")
    gam2, bet2 = 1 + 0.1 * torch.randn(256, generator=g, device="cuda"), 0.1 * torch.randn(256, generator=g, device="cuda")
    drop = None
    lin = a.float() @ w.float().t() + bias
    if p:
        seed = torch.tensor([31], dtype=torch.int32, device="cuda")
        drop = (seed, ops.drop_thr16(p), 5)
        lin = lin * _mask(31, 5, M, 256)
    zr = res.float() + lin
    yr = torch.nn.functional.layer_norm(zr, (256,), gam, bet, 1e-5)
    r = ops.gemm_res_ln(a, w, bias, res, gam, bet, drop=drop, res2=res2 if double else None, gamma2=gam2 if double else None,
                        beta2=bet2 if double else None)
    y, z, mean, rstd = r[:4]
    tag = f"gemm_res_ln_M{M}_K{K}_{'double' if double else 'single'}_p{p}"
    record(tag, "y.max_abs", _close(y, yr, "y", rel=1e-2, ab=2e-2))
    _close(z, zr, "z")
    assert float((mean - zr.mean(1)).abs().max()) < 1e-4
    assert float((rstd - (zr.var(1, unbiased=False) + 1e-5).rsqrt()).abs().max()) < 1e-3
    if double:
        y2, mean2, rstd2 = r[4:]
        z2 = res2.float() + y.float()          # the kernel normalises res2 + y as stored (bf16)
        y2r = torch.nn.functional.layer_norm(z2, (256,), gam2, bet2, 1e-5)
        record(tag, "y2.max_abs", _close(y2, y2r, "y2", rel=1e-2, ab=2e-2))
        assert float((mean2 - z2.mean(1)).abs().max()) < 1e-4


@pytest.mark.parametrize("M,Nout,Kin", [(8400, 2048, 256), (8400, 256, 2048), (8400, 512, 256), (8400, 256, 256),
                                        (800, 1536, 512), (77, 72, 40), (50400, 256, 256)])
def test_gemm_dw(M, Nout, Kin):
    from object_detection_destr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + Nout)
    dy = torch.randn(M, Nout, generator=g, device="cuda").to(BF)
    x = torch.randn(M, Kin, generator=g, device="cuda").to(BF)
    big = torch.full((Nout + 3, Kin + 8), 0.5, device="cuda")
    dw = big[1:Nout + 1, 4:Kin + 4]                                  # a view with a row pitch and an offset
    ops.gemm_dw(dy, x, dw)
    ref = dy.float().t() @ x.float() + 0.5
    err = float((dw - ref).abs().max()) / float(ref.abs().max())
    record(f"gemm_dw_M{M}_{Nout}x{Kin}", "max_err_over_absmax", err, 2e-3)
    assert err <= 2e-3, err
    assert float((big[0] - 0.5).abs().max()) == 0 and float((big[:, :4] - 0.5).abs().max()) == 0  # nothing outside the view
