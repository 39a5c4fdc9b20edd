"""-m gpu: drop-in decoder modules, matcher and the whole transformer half vs golden / oracle."""
from argparse import Namespace

import pytest
import torch

from oracle import destr_oracle as O

pytestmark = pytest.mark.gpu


def _cmp(got, ref, max_tol, mean_tol, what):
    err = (got.float().cpu() - ref).abs()
    print(f"{what}: max {err.max():.3e} mean {err.mean():.3e} (ref absmax {ref.abs().max():.3e})")
    assert torch.isfinite(got).all()
    assert err.max() <= max_tol and err.mean() <= mean_tol, what


def test_decoder_kernels():
    from tools import gpu_check
    assert gpu_check.check_decoder_kernels(2, 100, 300)
    assert gpu_check.check_decoder_kernels(1, 300, 1050)
    assert gpu_check.check_decoder_kernels(3, 40, 54)


def _golden_decoder(golden):
    from object_detection_destr_b200.decoder import build_decoder
    from object_detection_destr_b200.encoder import disable_dropout
    L = golden["dec_layers"]
    dec = build_decoder(Namespace(hidden_dim=256, num_decoder_blocks=L))
    dec.load_state_dict(O.make_decoder_weights(L, seed=golden["dec_seed"]), strict=True)
    _, bbox_sd = O.make_head_weights(5, seed=golden["head_seed"])
    bbox = torch.nn.Sequential(torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, 4))
    bbox.load_state_dict(bbox_sd)
    return disable_dropout(dec).cuda(), bbox.cuda(), bbox_sd


def test_decoder_golden(golden):
    dec, bbox, bbox_sd = _golden_decoder(golden)
    centers = golden["dec_centers"]
    pos_embed = O.query_sine_embed(centers, 256)
    with torch.no_grad():
        out = dec(golden["dec_x"].cuda(), golden["dec_enc_tok"].cuda(), golden["enc_mask"].flatten(1).cuda(),
                  golden["dec_fine_pos"].cuda(), pos_embed.cuda(), centers.cuda(), bbox)
    assert out.shape == golden["dec_out"].shape
    _cmp(out, golden["dec_out"], 1e-1, 8e-3, "decoder vs reference golden")


def test_decoder_block_and_attention_modules(golden):
    """Module-level swap points: SelfAttention / PairSelfAttention with the reference's own
    (B,H,S,d) calling convention."""
    from object_detection_destr_b200.decoder import PairSelfAttention, SelfAttention
    sa = SelfAttention(heads_num=8, dropout_prob=0.0, hidden_dim=256)
    _cmp(sa(golden["sa_q"].cuda(), golden["sa_k"].cuda(), golden["sa_v"].cuda()), golden["sa_out"], 3e-2, 4e-3, "SelfAttention")
    ca = SelfAttention(heads_num=1, dropout_prob=0.0)
    _cmp(ca(golden["ca_q"].cuda(), golden["ca_k"].cuda(), golden["ca_v"].cuda(), key_padding_mask=golden["ca_kpm"].cuda()),
         golden["ca_out"], 3e-2, 4e-3, "cross SelfAttention")
    pa = PairSelfAttention(heads_num=8)
    _cmp(pa(golden["pa_q"].cuda(), golden["pa_k"].cuda(), golden["pa_v"].cuda(), golden["pa_coords"].cuda()),
         golden["pa_out"], 1e-2, 1e-3, "PairSelfAttention")


def test_matcher_bit_exact(golden):
    from object_detection_destr_b200.matcher import HungarianMatcher, HungarianMatcherWoL1
    outputs = {"pred_class": golden["m_logits"].cuda(), "pred_boxes": golden["m_boxes"].cuda()}
    t_int = [{"labels": l.cuda(), "boxes": b.cuda()} for l, b in zip(golden["m_labels"], golden["m_tboxes"])]
    t_oh = [{"labels": torch.nn.functional.one_hot(l, 2).cuda(), "boxes": b.cuda()}
            for l, b in zip(golden["m_labels"], golden["m_tboxes"])]
    for key, m, t in (("m_hung_121", HungarianMatcher(1.0, 2.0, 1.0), t_oh), ("m_hung_default", HungarianMatcher(), t_oh),
                      ("m_wol1", HungarianMatcherWoL1(0.5, 0.5), t_int)):
        got = m(outputs, t)
        for (gi, gj), (ri, rj) in zip(got, golden[key]):
            assert gi.device.type == "cpu" and gi.dtype == torch.int64
            assert torch.equal(gi, ri) and torch.equal(gj, rj), key


def test_matcher_c5_shape_bit_exact():
    """config-5 upper end: B=64, Q=300, C=91 -- assignments identical to the oracle's fp32 cost + scipy."""
    from object_detection_destr_b200.matcher import HungarianMatcherWoL1
    g = torch.Generator().manual_seed(11)
    B, Q, C = 64, 300, 91
    logits = torch.randn(B, Q, C, generator=g)
    boxes = torch.cat([0.1 + 0.8 * torch.rand(B, Q, 2, generator=g), 0.03 + 0.4 * torch.rand(B, Q, 2, generator=g)], -1)
    labels, tboxes = O.make_targets(B, seed=9)
    ref = O.hungarian_match(O.match_cost_blocks(logits, boxes, labels, tboxes, 0.5, 0.0, 0.5, with_l1=False))
    got = HungarianMatcherWoL1(0.5, 0.5)({"pred_class": logits.cuda(), "pred_boxes": boxes.cuda()},
                                         [{"labels": l.cuda(), "boxes": b.cuda()} for l, b in zip(labels, tboxes)])
    assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(got, ref))


@pytest.mark.parametrize("runtime", [False, True])
def test_transformer_half_fwd_bwd(runtime):
    """Whole hot path (2+2 layers, N=1050 padded, Q=100): forward vs oracle with the oracle's pairing injected
    (SURVEY 7.3-3), gradients vs stock-torch bf16 autocast yardstick."""
    from object_detection_destr_b200.encoder import disable_dropout
    from object_detection_destr_b200.hotpath import TransformerHalf
    from object_detection_destr_b200 import functional as Fn
    L, B, H, W, Q, C = 2, 2, 25, 42, 100, 91
    g = torch.Generator().manual_seed(17)
    enc_sd, dec_sd = O.make_encoder_weights(L, seed=31), O.make_decoder_weights(L, seed=32)
    cls_sd, bbox_sd = O.make_head_weights(C, seed=33)
    feats = torch.randn(B, 256, H, W, generator=g)
    mask = torch.zeros(B, H, W, dtype=torch.bool)
    mask[1, :, 33:] = True
    sel = torch.randn(B, Q, 512, generator=g)
    centers = 0.05 + 0.9 * torch.rand(B, Q, 2, generator=g)

    def oracle(sd_e, sd_d, sd_c, sd_b, f, dev):
        pos = O.sine_pos2d(mask.to(dev))
        enc = O.encoder_forward(f, mask.to(dev), pos, sd_e, L)
        fine = O.fine_pos_tokens(enc, pos, sd_e)
        c = centers.to(dev)
        dec, coords = O.decoder_forward(sel.to(dev), enc.flatten(2).transpose(1, 2), mask.flatten(1).to(dev), fine,
                                        O.query_sine_embed(c, 256), c, sd_d, sd_b, L, return_coords=True)
        return O.heads_forward(dec, c, sd_c, sd_b), coords

    req = lambda sd, dev="cpu": {k: v.clone().to(dev).requires_grad_() for k, v in sd.items()}
    r_e, r_d, r_c, r_b = req(enc_sd), req(dec_sd), req(cls_sd), req(bbox_sd)
    ref, ref_coords = oracle(r_e, r_d, r_c, r_b, feats, "cpu")
    ref_pairs = [O.get_pairs(c.detach()).int().cuda() for c in ref_coords]
    gcls = torch.randn(B, Q, C, generator=g)
    gbox = torch.randn(B, Q, 4, generator=g)
    (ref["pred_class"] * gcls).sum().add((ref["pred_boxes"] * gbox).sum()).backward()

    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=L, num_decoder_blocks=L, num_cls=C),
                            runtime=runtime)
    model._encoder.load_state_dict(enc_sd)
    model._decoder.load_state_dict(dec_sd)
    model._cls_embed.load_state_dict(cls_sd)
    model._bbox_embed.load_state_dict(bbox_sd)
    disable_dropout(model).cuda()
    # inject the oracle's pairing so the comparison is not dominated by argmax flips under bf16
    out, _ = model(feats.cuda(), mask.cuda(), sel.cuda(), centers.cuda(), pairs_override=ref_pairs)
    (out["pred_class"] * gcls.cuda()).sum().add((out["pred_boxes"] * gbox.cuda()).sum()).backward()
    _cmp(out["pred_class"], ref["pred_class"].detach(), 1.5e-1, 1.5e-2, "pred_class (logits)")
    _cmp(out["pred_boxes"], ref["pred_boxes"].detach(), 2e-2, 2e-3, "pred_boxes")
    # own pairing (no injection) must agree with the oracle's on (almost) every query
    with torch.no_grad():
        aux2 = []
        model(feats.cuda(), mask.cuda(), sel.cuda(), centers.cuda(), aux=aux2)
    for l, (coords, pairs) in enumerate(aux2):
        agree = (pairs.cpu() == ref_pairs[l].cpu()).all(-1).float().mean()
        cerr = (coords.view(B, Q, 4).cpu() - ref_coords[l].detach()).abs().max()
        print(f"layer {l}: pairing agreement {agree:.3f}, coords max abs err {cerr:.2e}")
        assert agree >= 0.9 and cerr <= 2e-2

    # gradient yardstick: the oracle graph under stock torch bf16 autocast on the GPU
    y_e, y_d, y_c, y_b = req(enc_sd, "cuda"), req(dec_sd, "cuda"), req(cls_sd, "cuda"), req(bbox_sd, "cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yard, _ = oracle(y_e, y_d, y_c, y_b, feats.cuda(), "cuda")
    (yard["pred_class"].float() * gcls.cuda()).sum().add((yard["pred_boxes"].float() * gbox.cuda()).sum()).backward()

    def rel(a, b):
        return float((a.float().cpu() - b).norm() / b.norm())

    named = dict(model.named_parameters())
    # EVERY parameter of the hot path (weights, biases, LayerNorm affine, shared and hoisted groups)
    checks = []
    for prefix, r_sd, y_sd in (("_encoder.", r_e, y_e), ("_decoder.", r_d, y_d), ("_cls_embed.", r_c, y_c),
                               ("_bbox_embed.", r_b, y_b)):
        for k in r_sd:
            if "_proj_to_q." in k or "_proj_to_k." in k or "_proj_to_v." in k:
                if prefix == "_encoder.":
                    assert named[prefix + k].grad is None  # dead parameters
                    continue
            checks.append((prefix + k, r_sd[k], y_sd[k]))
    assert len(checks) > 100
    for name, r, y in checks:
        ours, yd = rel(named[name].grad, r.grad), rel(y.grad, r.grad)
        if ours > max(1.5 * yd, 2e-2):
            print(f"grad {name}: rel-fro ours {ours:.3e}  torch-autocast {yd:.3e}")
        assert ours <= max(2.5 * yd, 4e-2), name
    print(f"checked {len(checks)} parameter gradients")
