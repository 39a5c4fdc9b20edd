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
    (8, 1050, 8, "random", True), (5, 2100, 8, "random", True)])  # last two: several items per persistent CTA
def test_enc_attn_fwd(B, N, heads, mode, masked):
    from tools import gpu_check
    assert gpu_check.attn_case(B, N, heads, mode, masked)


@pytest.mark.parametrize("B,N,heads,masked", [(1, 128, 1, False), (1, 256, 2, False), (2, 300, 8, True),
                                              (3, 54, 8, True), (8, 1050, 8, True), (5, 2100, 8, True)])
def test_enc_attn_bwd(B, N, heads, masked):
    from tools import gpu_check
    assert gpu_check.attn_bwd_case(B, N, heads, masked)


@pytest.mark.parametrize("M,C,with_relu", [(800, 1024, True), (8400, 2048, True), (8400, 512, False), (37, 256, True)])
def test_relu_bwd_colsum(M, C, with_relu):
    """dpre = dy * (h > 0), dbias += colsum(dpre); both block shapes (short / FFN-sized) and the plain column sum."""
    from object_detection_destr_b200 import ops
    g = torch.Generator().manual_seed(M + C)
    dy = torch.randn(M, C, generator=g).bfloat16()
    h = torch.randn(M, C, generator=g).bfloat16().clamp(min=0) if with_relu else None
    dbias = torch.full((C,), 0.5, dtype=torch.float32, device="cuda")
    dpre = ops.relu_bwd_colsum(dy.cuda(), None if h is None else h.cuda(), dbias)
    ref = dy.float() * (h.float() > 0) if with_relu else dy.float()
    if with_relu:
        assert torch.equal(dpre.cpu().float(), ref)
    else:
        assert dpre is None
    exp = 0.5 + ref.sum(0)
    assert torch.allclose(dbias.cpu(), exp, rtol=1e-4, atol=1e-2 * float(ref.abs().sum(0).max()) / M ** 0.5 + 1e-3)


def test_sine_pos2d_wide_and_padded():
    from oracle import destr_oracle as O
    from object_detection_destr_b200 import ops
    mask = torch.zeros(2, 50, 84, dtype=torch.bool)
    mask[1, 40:, :] = True
    mask[1, :, 70:] = True
    pf, _ = ops.sine_pos2d(mask.cuda(), want_f32=True, want_bf16=False)
    ref = O.sine_pos2d(mask).flatten(2).transpose(1, 2)
    assert float((pf.cpu() - ref).abs().max()) < 2e-5


@pytest.mark.parametrize("M,C", [(800, 91), (37, 7), (8, 128)])
def test_heads_fwd_bwd(M, C):
    """csrc/heads.cu against the plain torch fp32 statement of model.py:120-131 (same bf16 decoder output)."""
    import torch
    from object_detection_destr_b200 import ops
    g = torch.Generator().manual_seed(M + C)
    dec = (torch.randn(M, 512, generator=g) * 0.7).bfloat16()
    centers = torch.rand(M, 2, generator=g) * 0.98 + 0.01
    centers[0, 0] = 0.0  # clamped by inverse_sigmoid (misc.py:59-62)
    P = [torch.randn(C, 256, generator=g) * 0.06, torch.randn(C, generator=g) * 0.1,
         torch.randn(256, 256, generator=g) * 0.06, torch.randn(256, generator=g) * 0.1,
         torch.randn(4, 256, generator=g) * 0.06, torch.randn(4, generator=g) * 0.1]
    dl, db = torch.randn(M, C, generator=g), torch.randn(M, 4, generator=g)

    def ref(dec_f, Wc, bc, W1, b1, W2, b2):
        logits = dec_f[:, :256] @ Wc.t() + bc
        delta = torch.relu(dec_f[:, 256:] @ W1.t() + b1) @ W2.t() + b2
        inv = -torch.log(1.0 / centers.double().clamp(min=1e-6) - 1.0)
        return logits, torch.cat([delta[:, :2] + inv, delta[:, 2:]], -1).sigmoid()

    dec_r = dec.double().requires_grad_()
    Pr = [p.double().requires_grad_() for p in P]
    lr, br = ref(dec_r, *Pr)
    ((lr * dl.double()).sum() + (br * db.double()).sum()).backward()

    dec_d = dec.cuda().requires_grad_()
    Pd = [p.cuda().requires_grad_() for p in P]
    lo, bo = ops.heads(dec_d, centers.cuda(), *Pd)
    ((lo * dl.cuda()).sum() + (bo * db.cuda()).sum()).backward()
    assert torch.allclose(lo.cpu().double(), lr, rtol=1e-5, atol=1e-5)
    assert torch.allclose(bo.cpu().double(), br, rtol=1e-5, atol=1e-6)
    # d_dec leaves in bf16 (the decoder's gradient dtype): 2^-8 relative
    assert torch.allclose(dec_d.grad.cpu().double(), dec_r.grad, rtol=1e-2, atol=2e-3)
    for got, exp in zip(Pd, Pr):
        assert torch.allclose(got.grad.cpu().double(), exp.grad, rtol=1e-4, atol=1e-4 * float(exp.grad.abs().max()))


@pytest.mark.parametrize("M,N,K,p", [(8400, 2048, 256, 0.3), (800, 1024, 256, 0.3), (37, 256, 256, 0.0), (33600, 2048, 256, 0.3)])
def test_linear_bias_relu_dropout(M, N, K, p):
    """tcgen05 GEMM + bias + ReLU + dropout epilogue (csrc/gemm_bias_relu.cu) vs fp32 torch with the numpy twin's mask."""
    import numpy as np
    import torch
    from object_detection_destr_b200 import ops
    from oracle.dropout_mask import keep_mask, scale_of, thr16_of
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16()
    b = torch.randn(N, generator=g) * 0.5
    ref = torch.relu(x.float() @ w.float().t() + b)
    drop = None
    if p > 0:
        t = thr16_of(p)
        ref = ref * torch.from_numpy(keep_mask(5, 77, np.arange(M), np.arange(N), t)).float() * scale_of(t)
        drop = (torch.tensor([5], dtype=torch.int32, device="cuda"), ops.drop_thr16(p), 77)
    out = ops.linear_bias_relu_dropout(x.cuda(), w.cuda(), b.cuda(), drop).cpu().float()
    assert torch.isfinite(out).all()
    # bf16 output of an fp32-accumulated product: 2^-8 relative
    assert torch.allclose(out, ref, rtol=1e-2, atol=1e-2), float((out - ref).abs().max())
    if p > 0:  # exactly the twin's zero pattern (ReLU zeros aside)
        assert torch.equal((out == 0) | (ref == 0), ref == 0) or float(((out == 0) != (ref == 0)).float().mean()) < 1e-3
    # a strided input view (row pitch > K) gives the same result
    xs = torch.zeros(M, K + 64, dtype=torch.bfloat16)
    xs[:, :K] = x
    out2 = ops.linear_bias_relu_dropout(xs.cuda()[:, :K], w.cuda(), b.cuda(), drop).cpu().float()
    assert torch.equal(out2, out)


@pytest.mark.parametrize("M,p", [(8400, 0.0), (8400, 0.3), (803, 0.3)])
def test_add_layernorm2_fwd_bwd(M, p):
    """Two chained LayerNorms in one pass, y1 = LN1(a + dropout(b)), y2 = LN2(c + y1), and its one-pass backward,
    vs torch autograd on the same bf16 inputs (the dropout mask from the numpy twin of the kernels' hash)."""
    import numpy as np
    from object_detection_destr_b200 import ops
    from oracle.dropout_mask import keep_mask, scale_of, thr16_of
    from parity_log import record
    g = torch.Generator().manual_seed(M)
    a, b, c, dy = (torch.randn(M, 256, generator=g).bfloat16() for _ in range(4))
    g1, b1, g2, b2 = (1 + 0.1 * torch.randn(256, generator=g), 0.1 * torch.randn(256, generator=g),
                      1 + 0.1 * torch.randn(256, generator=g), 0.1 * torch.randn(256, generator=g))
    mk = torch.ones(M, 256)
    drop = None
    if p:
        t = thr16_of(p)
        mk = torch.from_numpy(keep_mask(9, 4, np.arange(M), np.arange(256), t)).float() * scale_of(t)
        drop = (torch.tensor([9], dtype=torch.int32, device="cuda"), ops.drop_thr16(p), 4)
    af, bf, cf = (t.float().requires_grad_() for t in (a, b, c))
    p1 = [t.clone().requires_grad_() for t in (g1, b1, g2, b2)]
    LN = torch.nn.functional.layer_norm
    y1r = LN(af + bf * mk, (256,), p1[0], p1[1], 1e-5)
    y2r = LN(cf + y1r, (256,), p1[2], p1[3], 1e-5)
    y2r.backward(dy.float())
    dev = [t.cuda() for t in (g1, b1, g2, b2)]
    y1, m1, r1, y2, m2, r2 = ops.add_layernorm2(a.cuda(), b.cuda(), dev[0], dev[1], c.cuda(), dev[2], dev[3], drop=drop)
    assert float((y1.cpu().float() - y1r).abs().max()) < 3e-2 and float((y2.cpu().float() - y2r).abs().max()) < 4e-2
    grads = [torch.zeros(256, device="cuda") for _ in range(5)]  # dgamma2, dbeta2, dgamma1, dbeta1, dbias
    d3, dxb, dsum = ops.add_layernorm2_bwd(dy.cuda(), c.cuda(), y1, dev[2], m2, r2, a.cuda(), b.cuda(), dev[0], m1, r1,
                                           grads[0], grads[1], grads[2], grads[3], dbias=grads[4], drop=drop, want_sum=True)
    e3 = float((d3.cpu().float() - cf.grad).abs().max())      # gradient of the residual stream c (= d(c + y1))
    eb = float((dxb.cpu().float() - bf.grad).abs().max())     # through b's dropout mask
    es = float((dsum.cpu().float() - af.grad).abs().max())    # un-masked
    record(f"add_layernorm2_M{M}_p{p}", "d3/dxb/dsum.max_abs", max(e3, eb, es), 5e-2)
    assert max(e3, eb, es) < 5e-2, (e3, eb, es)
    for got, ref in ((grads[0], p1[2].grad), (grads[1], p1[3].grad), (grads[2], p1[0].grad), (grads[3], p1[1].grad),
                     (grads[4], bf.grad.sum(0))):
        assert torch.allclose(got.cpu(), ref, rtol=3e-2, atol=0.5), float((got.cpu() - ref).abs().max())


@pytest.mark.gpu
def test_copy_many_matches_copy():
    """destr_copy_many (the engine's batch hand-over): every pair copied bit-exactly, aligned and unaligned, 1 B .. 4 MB."""
    from object_detection_destr_b200 import ops
    g = torch.Generator().manual_seed(0)
    sizes = [1, 36, 6400, 8400, 4300800, 17, 1 << 20, 16, 15, 800 * 512 * 2]
    base = [torch.randint(0, 256, (n + 64,), generator=g, dtype=torch.uint8).cuda() for n in sizes]
    srcs = [b[(k % 3):(k % 3) + n] for k, (b, n) in enumerate(zip(base, sizes))]      # some views start off 16-byte alignment
    dsts = [torch.zeros(n + 32, dtype=torch.uint8, device="cuda")[(k % 5):(k % 5) + n] for k, n in enumerate(sizes)]
    ops.copy_many(dsts, srcs)
    torch.cuda.synchronize()
    for d, s in zip(dsts, srcs):
        assert torch.equal(d, s)
    f = torch.randn(8400, 256, generator=g).bfloat16().cuda()
    out = torch.empty_like(f)
    ops.copy_many([out], [f])
    assert torch.equal(out, f)
    with pytest.raises(ValueError):
        ops.copy_many([out], [f[:100]])
    with pytest.raises(RuntimeError):
        ops.copy_many([out.cpu()], [f.cpu()])
