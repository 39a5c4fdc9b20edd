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


class TwinDropper:
    """Feeds the oracle the masks the kernels generate (site numbering of runtime.py, rows/cols as in the kernels)."""

    def __init__(self, seed, p=P):
        from oracle.dropout_mask import thr16_of, scale_of
        self.seed, self.thr = seed, thr16_of(p)
        self.scale = scale_of(self.thr)

    @staticmethod
    def _site(prefix, name):
        import re
        from object_detection_destr_b200.runtime import dec_site, enc_site
        l = int(re.search(r"\.(\d+)\.", prefix).group(1))
        if prefix.startswith("_encoder."):
            return enc_site(l, name)
        if "_branch." in prefix:
            if name == "ca":
                return dec_site(l, "ca")
            return dec_site(l, ("b0." if "_cls_branch" in prefix else "b1.") + name)
        return dec_site(l, name)

    def _m(self, site, rows, ncols, shape):
        from oracle.dropout_mask import keep_mask
        return (torch.from_numpy(keep_mask(self.seed, site, rows, np.arange(ncols), self.thr)).float() * self.scale).view(shape)

    def rows(self, prefix, name, x):
        B, S, C = x.shape
        return x * self._m(self._site(prefix, name), np.arange(B * S), C, (B, S, C)).to(x.device)

    def attn(self, prefix, name, p):
        B, H, Q, K = p.shape
        if name == "ca":  # (B,1,Q,N): mask row = (b, branch, query)
            br = 0 if "_cls_branch" in prefix else 1
            rows = ((np.arange(B)[:, None] * 2 + br) * Q + np.arange(Q)[None, :]).reshape(-1)
        else:             # mask row = (b, head, query)
            rows = np.arange(B * H * Q)
        return p * self._m(self._site(prefix, name), rows, K, (B, H, Q, K)).to(p.device)


@pytest.mark.parametrize("runtime", [True, False])
def test_transformer_half_with_dropout_fwd_bwd(runtime):
    """The whole hot path in TRAINING mode with the reference's default dropout (p = 0.3 at all 16 sites per layer
    pair): outputs and every parameter gradient vs the fp32 oracle run with the SAME masks."""
    from argparse import Namespace
    from oracle import destr_oracle as O
    from object_detection_destr_b200.hotpath import TransformerHalf
    L, B, H, W, Q, C = 2, 2, 25, 42, 100, 91
    g = torch.Generator().manual_seed(19)
    enc_sd, dec_sd = O.make_encoder_weights(L, seed=41), O.make_decoder_weights(L, seed=42)
    cls_sd, bbox_sd = O.make_head_weights(C, seed=43)
    feats = torch.randn(B, 256, H, W, generator=g)
    mask = torch.zeros(B, H, W, dtype=torch.bool)
    mask[1, :, 35:] = True
    sel = torch.randn(B, Q, 512, generator=g)
    centers = 0.05 + 0.9 * torch.rand(B, Q, 2, generator=g)
    seed = 4242
    req = lambda sd: {k: v.clone().requires_grad_() for k, v in sd.items()}
    r_e, r_d, r_c, r_b = req(enc_sd), req(dec_sd), req(cls_sd), req(bbox_sd)
    with O.dropout(TwinDropper(seed)):
        pos = O.sine_pos2d(mask)
        enc = O.encoder_forward(feats, mask, pos, r_e, L)
        fine = O.fine_pos_tokens(enc, pos, r_e)
        dec, coords = O.decoder_forward(sel, enc.flatten(2).transpose(1, 2), mask.flatten(1), fine,
                                        O.query_sine_embed(centers, 256), centers, r_d, r_b, L, return_coords=True)
        ref = O.heads_forward(dec, centers, r_c, r_b)
    ref_pairs = [O.get_pairs(c.detach()).int().cuda() for c in coords]
    gcls, gbox = torch.randn(B, Q, C, generator=g), torch.randn(B, Q, 4, generator=g)
    (ref["pred_class"] * gcls).sum().add((ref["pred_boxes"] * gbox).sum()).backward()

    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=L, num_decoder_blocks=L, num_cls=C),
                            runtime=runtime)  # hand-scheduled runtime AND the module-level autograd path
    model._encoder.load_state_dict(enc_sd)
    model._decoder.load_state_dict(dec_sd)
    model._cls_embed.load_state_dict(cls_sd)
    model._bbox_embed.load_state_dict(bbox_sd)
    model.cuda().train()  # dropout stays at the reference defaults
    model.set_dropout_seed(seed)
    out, _ = model(feats.cuda(), mask.cuda(), sel.cuda(), centers.cuda(), pairs_override=ref_pairs)
    (out["pred_class"] * gcls.cuda()).sum().add((out["pred_boxes"] * gbox.cuda()).sum()).backward()
    rel = lambda a, b: float((a.float().cpu() - b).norm() / b.norm())
    assert rel(out["pred_class"], ref["pred_class"].detach()) < 3e-2
    assert float((out["pred_boxes"].cpu() - ref["pred_boxes"].detach()).abs().max()) < 3e-2
    named = dict(model.named_parameters())
    worst = 0.0
    n = 0
    for prefix, r_sd in (("_encoder.", r_e), ("_decoder.", r_d), ("_cls_embed.", r_c), ("_bbox_embed.", r_b)):
        for k, v in r_sd.items():
            if v.grad is None or named[prefix + k].grad is None:
                continue
            e = rel(named[prefix + k].grad, v.grad)
            worst = max(worst, e)
            n += 1
            assert e < 8e-2, (prefix + k, e)
    print(f"{n} parameter gradients with dropout, worst rel-fro error {worst:.3e}")
    assert n > 100
    # a different seed gives a different forward (the masks really depend on it)
    model.set_dropout_seed(seed + 1)
    with torch.no_grad():
        out2, _ = model(feats.cuda(), mask.cuda(), sel.cuda(), centers.cuda(), pairs_override=ref_pairs)
    assert float((out2["pred_class"] - out["pred_class"]).abs().max()) > 1e-2


def test_standalone_modules_apply_the_reference_dropout():
    """Drop-in modules used on their own: EncoderBlock in train mode drops (and is reproducible under a pinned seed),
    in eval mode it does not; SelfAttention drops ALWAYS, like the reference's inline nn.Dropout (self_attention.py:40)."""
    from object_detection_destr_b200.decoder import SelfAttention
    from object_detection_destr_b200.encoder import EncoderBlock, set_dropout_seed
    g = torch.Generator().manual_seed(5)
    blk = EncoderBlock().cuda()
    x, pos = torch.randn(300, 2, 256, generator=g).cuda(), torch.randn(300, 2, 256, generator=g).cuda()
    blk.eval()
    with torch.no_grad():
        y_eval = blk(x, pos_embed=pos)
        assert torch.equal(y_eval, blk(x, pos_embed=pos))
        blk.train()
        set_dropout_seed(blk, 9)
        y1, y2 = blk(x, pos_embed=pos), blk(x, pos_embed=pos)
        assert torch.equal(y1, y2) and float((y1 - y_eval).abs().max()) > 0.05
        set_dropout_seed(blk, None)          # released: the seed advances on every forward
        assert not torch.equal(blk(x, pos_embed=pos), blk(x, pos_embed=pos))
        sa = SelfAttention(heads_num=8).cuda().eval()
        q, k, v = (torch.randn(2, 8, 100, 64, generator=g).cuda() for _ in range(3))
        assert not torch.equal(sa(q, k, v), sa(q, k, v))          # stochastic even in eval
        sa._dropout_prob = 0.0
        assert torch.equal(sa(q, k, v), sa(q, k, v))


def test_two_forwards_then_backward_uses_each_forwards_own_masks():
    """the module-level autograd path used to keep a reference to the module's LIVE seed counter, so a
    second forward before backward (gradient accumulation, two views per step) made the first forward's backward
    recompute different masks.  Each forward now snapshots its seed: backward of forward #1 after forward #2 equals
    backward of forward #1 alone."""
    from object_detection_destr_b200.encoder import EncoderBlock, set_dropout_seed
    g = torch.Generator().manual_seed(6)
    blk = EncoderBlock().cuda().train()
    x = torch.randn(200, 2, 256, generator=g).cuda()
    pos = torch.randn(200, 2, 256, generator=g).cuda()
    dy = torch.randn(200, 2, 256, generator=g).cuda()

    def grads(second_forward: bool):
        set_dropout_seed(blk, 100)
        set_dropout_seed(blk, None)       # released at a known value: every forward advances the counter
        blk._drop_seed_t.fill_(100)
        blk.zero_grad()
        y1 = blk(x, pos_embed=pos)
        if second_forward:
            with torch.no_grad():
                blk(x * 0.5, pos_embed=pos)    # moves the module's seed before y1's backward runs
        y1.backward(dy)
        return y1.detach().clone(), {k: v.grad.clone() for k, v in blk.named_parameters() if v.grad is not None}

    ya, ga = grads(False)
    yb, gb = grads(True)
    assert torch.equal(ya, yb)
    assert len(ga) >= 12
    for k in ga:   # (LayerNorm-affine / bias gradients are accumulated with atomics: equal up to summation order;
        #            a different mask would change them by O(1))
        assert float((ga[k] - gb[k]).abs().max()) <= 1e-3 * float(ga[k].abs().max()) + 1e-6, k
