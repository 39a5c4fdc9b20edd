"""Generate golden input/output vectors by running the UNMODIFIED reference on CPU fp32.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/golden.pt (inputs + reference outputs; weights are regenerated from seeds by
oracle.destr_oracle.make_*_weights and guarded by checksums stored in the file).
Dropout is neutralised exactly as SURVEY.md section 8(c) prescribes: `.eval()` for nn.Dropout /
nn.MultiheadAttention and `SelfAttention._dropout_prob = 0` (self_attention.py:15,40 builds its
Dropout inline, so eval() alone does not disable it).
"""
import os
import sys
from argparse import Namespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("DESTR_REF", "/root/reference"))

from oracle import destr_oracle as O  # noqa: E402

from src.model.attention.self_attention import SelfAttention  # noqa: E402
from src.model.attention.pair_self_attention import PairSelfAttention, _get_pairs  # noqa: E402
from src.model.blocks.encoder_block import build_encoder  # noqa: E402
from src.model.blocks.decoder_block import build_decoder  # noqa: E402
from src.utils.positional_embedding import gen_sineembed_for_position  # noqa: E402
from src.utils.position_encoding_cdetr import build_position_encoding_fix  # noqa: E402
from src.utils.misc import NestedTensor, inverse_sigmoid, sigmoid_focal_loss  # noqa: E402
from src.utils.bbox_utils import complete_iou, from_cxcyhw_to_xyxy, from_xyxy_to_cxcyhw, get_iou  # noqa: E402
from src.utils.matcher import HungarianMatcher, HungarianMatcherWoL1  # noqa: E402
from src.utils.criterion import SetCriterion, CompleteIOULoss  # noqa: E402


def neutralise(m):
    m.eval()
    for s in m.modules():
        if isinstance(s, SelfAttention):
            s._dropout_prob = 0.0
    return m


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def padded_mask(B, H, W, g):
    mask = torch.zeros(B, H, W, dtype=torch.bool)
    for b in range(1, B):
        hh = int(torch.randint(H // 2 + 1, H + 1, (1,), generator=g))
        ww = int(torch.randint(W // 2 + 1, W + 1, (1,), generator=g))
        mask[b, hh:, :] = True
        mask[b, :, ww:] = True
    return mask


def main():
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1234)
    G = {}

    # ---- known-answer vectors from the reference's own __main__ snippets ----
    kat_boxes = torch.tensor([[4, 8, 2, 2], [7, 4, 4, 4], [2, 10, 2, 8]], dtype=torch.float32)[None] / 20
    G["kat_pairs_in"] = kat_boxes
    G["kat_pairs_out"] = _get_pairs(kat_boxes)
    G["kat_ciou_p"] = from_cxcyhw_to_xyxy(torch.tensor([[0.3, 0.4, 0.4, 0.4]]))
    G["kat_ciou_g"] = torch.tensor([[.15, .25, .55, .7], [.6, .6, .9, .9]])
    G["kat_ciou_out"] = complete_iou(torch.tensor([[.1, .2, .5, .6]]), G["kat_ciou_g"])
    G["kat_sine_in"] = torch.tensor([[[0.25, 0.5]]])
    G["kat_sine_out"] = gen_sineembed_for_position(G["kat_sine_in"], 256)
    G["kat_invsig_in"] = torch.tensor([0.0, 0.5, 1.0, 0.3, 1e-7])
    G["kat_invsig_out"] = inverse_sigmoid(G["kat_invsig_in"])

    # ---- a1: image-grid sine embedding on a padded mask ----
    m = padded_mask(3, 6, 9, g)
    G["pos2d_mask"] = m
    G["pos2d_out"] = build_position_encoding_fix()(NestedTensor(torch.zeros(3, 1, 6, 9), m))

    # ---- a11 random ----
    c = torch.rand(2, 7, 2, generator=g)
    G["qsine_in"], G["qsine_out"] = c, gen_sineembed_for_position(c, 256)

    # ---- a12 random boxes ----
    def rand_xyxy(n):
        xy = torch.rand(n, 2, generator=g) * 0.6
        wh = 0.02 + torch.rand(n, 2, generator=g) * 0.35
        return torch.cat([xy, xy + wh], -1)

    def rand_cxcyhw(*shape):
        cxy = 0.1 + 0.8 * torch.rand(*shape, 2, generator=g)
        hw = 0.03 + 0.4 * torch.rand(*shape, 2, generator=g)
        return torch.cat([cxy, hw], -1)

    pb, gb = rand_cxcyhw(17), rand_xyxy(9)
    G["box_pred_cxcyhw"], G["box_gt_xyxy"] = pb, gb
    G["box_pred_xyxy"] = from_cxcyhw_to_xyxy(pb)
    G["box_back_cxcyhw"] = from_xyxy_to_cxcyhw(G["box_pred_xyxy"])
    G["box_iou"] = get_iou(G["box_pred_xyxy"], gb)
    G["box_ciou_cost"] = complete_iou(G["box_pred_xyxy"], gb)

    # ---- a10: pairs ----
    pc = rand_cxcyhw(3, 20)
    pc[1, 5] = pc[1, 11]  # duplicate box
    pc[2, :, 2:] *= 0.15  # tiny boxes: many disjoint (fake positive intersections), self-pairs
    G["pairs_in"], G["pairs_out"] = pc, _get_pairs(pc)

    # ---- a7: SelfAttention (decoder SA shape, cross-attn shape with kpm) ----
    sa = neutralise(SelfAttention(heads_num=8, dropout_prob=0.3, hidden_dim=256))
    q, k, v = (torch.randn(2, 8, 13, 64, generator=g) for _ in range(3))
    G["sa_q"], G["sa_k"], G["sa_v"], G["sa_out"] = q, k, v, sa(q, k, v)
    ca = neutralise(SelfAttention(heads_num=1))
    cq = torch.randn(2, 1, 11, 512, generator=g) * 0.3
    ck = torch.randn(2, 1, 37, 512, generator=g) * 0.3
    cv = torch.randn(2, 1, 37, 256, generator=g)
    ckpm = torch.zeros(2, 37, dtype=torch.bool)
    ckpm[1, 20:] = True
    ckpm[1, 3] = True
    G["ca_q"], G["ca_k"], G["ca_v"], G["ca_kpm"] = cq, ck, cv, ckpm
    G["ca_out"] = ca(cq, ck, cv, key_padding_mask=ckpm)

    # ---- a9: PairSelfAttention ----
    pa = PairSelfAttention(heads_num=8)
    pq, pk, pv = (torch.randn(3, 8, 20, 64, generator=g) * 0.5 for _ in range(3))
    G["pa_q"], G["pa_k"], G["pa_v"], G["pa_coords"] = pq, pk, pv, pc
    G["pa_out"] = pa(pq, pk, pv, pc)

    # ---- a2-a4: encoder, 2 layers, padded ----
    L = 2
    args = Namespace(hidden_dim=256, num_encoder_blocks=L, num_decoder_blocks=L)
    enc = neutralise(build_encoder(args))
    enc_sd = O.make_encoder_weights(L, seed=11)
    enc.load_state_dict(enc_sd, strict=True)
    B, H, W = 3, 6, 9
    ex = torch.randn(B, 256, H, W, generator=g)
    emask = m
    epos = G["pos2d_out"]
    with torch.no_grad():
        eout = enc(ex, emask, epos)
    G["enc_layers"], G["enc_seed"], G["enc_ck"] = L, 11, checksum(enc_sd)
    G["enc_x"], G["enc_mask"], G["enc_pos"], G["enc_out"] = ex, emask, epos, eout

    # one EncoderBlock alone (seq-first API) incl. the MHA
    blk = enc._encoder[0]
    xs = ex.flatten(2).permute(2, 0, 1)
    ps = epos.flatten(2).permute(2, 0, 1) * 0.5
    with torch.no_grad():
        G["encblk_out"] = blk(xs, key_mask=emask.flatten(1), pos_embed=ps)
        G["encmha_out"] = blk.self_attn(query=xs + ps, key=xs + ps, value=xs,
                                        key_padding_mask=emask.flatten(1))[0]
    G["encblk_pos"] = ps

    # ---- a5,a6,a8: decoder, 2 layers ----
    dec = neutralise(build_decoder(args))
    dec_sd = O.make_decoder_weights(L, seed=12)
    dec.load_state_dict(dec_sd, strict=True)
    cls_sd, bbox_sd = O.make_head_weights(5, seed=13)
    bbox = torch.nn.Sequential(torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, 4))
    bbox.load_state_dict(bbox_sd)
    Q = 10
    dx = torch.randn(B, Q, 512, generator=g)
    centers = 0.05 + 0.9 * torch.rand(B, Q, 2, generator=g)
    enc_tok = eout.flatten(2).transpose(1, 2).contiguous()
    with torch.no_grad():
        fine = (epos.flatten(2).permute(2, 0, 1) * enc._pos_scale(eout.flatten(2).permute(2, 0, 1))).permute(1, 0, 2).contiguous()
        pos_embed = gen_sineembed_for_position(centers, 256)
        dout = dec(selected_objects=dx, encoder_output=enc_tok, mask=emask.flatten(1), fine_pos=fine,
                   selected_objects_pos_embed=pos_embed, selected_centers=centers, bbox_embed=bbox)
    G["dec_layers"], G["dec_seed"], G["dec_ck"], G["head_seed"] = L, 12, checksum(dec_sd), 13
    G["dec_x"], G["dec_centers"], G["dec_fine_pos"], G["dec_out"] = dx, centers, fine, dout
    G["dec_enc_tok"] = enc_tok

    # ---- a13: matchers;  a14: criterion (2-class, as the reference supports) ----
    Bm, Qm, Cm = 3, 12, 2
    logits = torch.randn(Bm, Qm, Cm, generator=g)
    boxes = rand_cxcyhw(Bm, Qm)
    sizes = [4, 1, 15]  # the last image has more targets than queries
    tl = [torch.randint(0, Cm, (t,), generator=g) for t in sizes]
    tb = [rand_xyxy(t) for t in sizes]
    outputs = {"pred_class": logits, "pred_boxes": boxes}
    t_int = [{"labels": l, "boxes": b} for l, b in zip(tl, tb)]
    t_oh = [{"labels": torch.nn.functional.one_hot(l, Cm), "boxes": b} for l, b in zip(tl, tb)]
    G["m_logits"], G["m_boxes"], G["m_labels"], G["m_tboxes"] = logits, boxes, tl, tb
    G["m_hung_121"] = HungarianMatcher(1.0, 2.0, 1.0)(outputs, t_oh)
    G["m_hung_default"] = HungarianMatcher()(outputs, t_oh)
    G["m_wol1"] = HungarianMatcherWoL1(0.5, 0.5)(outputs, t_int)
    crit = SetCriterion(num_classes=2, matcher=HungarianMatcherWoL1(0.5, 0.5),
                        loss_fn={"class": sigmoid_focal_loss, "bbox": torch.nn.L1Loss(), "ciou": CompleteIOULoss()})
    # the reference criterion can only run when every image has T_i <= Q ... it can run anyway
    G["crit_out"] = {k: v.detach() for k, v in crit(outputs, t_int).items()}

    torch.save(G, os.path.join(HERE, "golden.pt"))
    n = sum(v.numel() * v.element_size() for v in G.values() if isinstance(v, torch.Tensor))
    print(f"wrote golden.pt: {len(G)} entries, ~{n/1e6:.2f} MB of tensors")


if __name__ == "__main__":
    main()
