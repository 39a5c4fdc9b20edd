"""Drives the UNMODIFIED reference modules (staged by oracle/stage_ref.py into oracle/_ref/) over the hot path.

TEST / BENCH INFRASTRUCTURE, not product code: only tests/, __graft_entry__.smoke() and bench.py's reference legs
import this.  Nothing here reimplements reference arithmetic -- every class used on the path is the reference's own:

    build_encoder / build_decoder            src/model/blocks/encoder_block.py:115-124, decoder_block.py:263-274
    PositionEmbeddingSine                    src/utils/position_encoding_cdetr.py:144-150
    gen_sineembed_for_position, inverse_sigmoid
    HungarianMatcherWoL1, SetCriterion, CompleteIOULoss, sigmoid_focal_loss, reduce_dict

`RefTransformerHalf.forward` is `ObjDetSplitTransformer.forward` (src/model/model.py:84-131) with the two
out-of-scope stages replaced by their outputs: the `reduce_dim(backbone(img))` features come in as an argument and
so do the mini-detector's `selected_objects` / `selected_centers` (model.py:100-102) -- the same inputs
`hotpath.TransformerHalf` takes.  Parameter names are the reference's (`_encoder.*`, `_decoder.*`, `_cls_embed.*`,
`_bbox_embed.*`).

One harness shim, applied identically on both arms: `SetCriterion._get_class_loss` hard-codes a
2-class one-hot (criterion.py:45) and raises for COCO-shaped 91-class labels; `SetCriterionC` widens that one-hot
to the logits width and changes nothing else.
"""
from __future__ import annotations

import os
import sys
import time
from argparse import Namespace
from typing import Dict, List

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_ROOT, "MANIFEST.json"))


_R = None


def ref():
    """Import the staged reference package (`src.*`) once; returns a namespace of its hot-path symbols."""
    global _R
    if _R is not None:
        return _R
    if not available():
        raise ImportError("oracle/_ref is missing: run `python oracle/stage_ref.py` in the authoring container "
                          "(it copies /root/reference/src; the copy travels to the GPU box with the snapshot)")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from src.model.blocks.encoder_block import build_encoder
    from src.model.blocks.decoder_block import build_decoder
    from src.model.attention.self_attention import SelfAttention
    from src.model.attention.pair_self_attention import PairSelfAttention
    from src.utils.position_encoding_cdetr import build_position_encoding_fix
    from src.utils.positional_embedding import gen_sineembed_for_position
    from src.utils.misc import NestedTensor, inverse_sigmoid, reduce_dict, sigmoid_focal_loss
    from src.utils.matcher import HungarianMatcher, HungarianMatcherWoL1
    from src.utils.criterion import CompleteIOULoss, SetCriterion
    _R = Namespace(build_encoder=build_encoder, build_decoder=build_decoder, SelfAttention=SelfAttention,
                   PairSelfAttention=PairSelfAttention, build_position_encoding_fix=build_position_encoding_fix,
                   gen_sineembed_for_position=gen_sineembed_for_position, NestedTensor=NestedTensor,
                   inverse_sigmoid=inverse_sigmoid, reduce_dict=reduce_dict, sigmoid_focal_loss=sigmoid_focal_loss,
                   HungarianMatcher=HungarianMatcher, HungarianMatcherWoL1=HungarianMatcherWoL1,
                   CompleteIOULoss=CompleteIOULoss, SetCriterion=SetCriterion)
    return _R


def neutralise_dropout(m: nn.Module) -> nn.Module:
    """`.eval()` for nn.Dropout / nn.MultiheadAttention and `_dropout_prob = 0` for the inline
    nn.Dropout of SelfAttention (self_attention.py:40), which eval() does not reach."""
    m.eval()
    for s in m.modules():
        if isinstance(s, ref().SelfAttention):
            s._dropout_prob = 0.0
    return m


def zero_dropout(m: nn.Module) -> nn.Module:
    """p = 0 at every dropout site while staying in train mode (the p = 0 training configuration of bench.py)."""
    for s in m.modules():
        if isinstance(s, nn.Dropout):
            s.p = 0.0
        elif isinstance(s, nn.MultiheadAttention):
            s.dropout = 0.0
        elif isinstance(s, ref().SelfAttention):
            s._dropout_prob = 0.0
    return m


class RefTransformerHalf(nn.Module):
    """The reference's transformer half: its own encoder, decoder and heads, wired as model.py:84-131 wires them."""

    def __init__(self, args: Namespace):
        super().__init__()
        R = ref()
        d = args.hidden_dim
        self._hidden_dim = d
        self._encoder = R.build_encoder(args)
        self._decoder = R.build_decoder(args)
        self._cls_embed = nn.Linear(d, args.num_cls)                                     # model.py:30-32
        self._bbox_embed = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Linear(d, 4))    # model.py:33-39
        self._pos = R.build_position_encoding_fix()

    def forward(self, features: torch.Tensor, mask: torch.Tensor, selected_objects: torch.Tensor,
                selected_centers: torch.Tensor):
        R = ref()
        B, C, H, W = features.shape
        pos = self._pos(R.NestedTensor(features, mask)).to(features.dtype)               # backbone.py:159
        x = self._encoder(features, mask, pos)                                           # model.py:84
        enc = x
        fine_pos = pos.flatten(2).permute(2, 0, 1)                                       # model.py:89-97
        fine_pos = fine_pos * self._encoder._pos_scale(x.flatten(2).permute(2, 0, 1).contiguous())
        fine_pos = fine_pos.view(H, W, B, -1).permute(2, 3, 0, 1).contiguous()
        pos_embed = R.gen_sineembed_for_position(selected_centers, self._hidden_dim)     # model.py:104-106
        y = self._decoder(selected_objects=selected_objects,                             # model.py:108-118
                          encoder_output=enc.flatten(2).transpose(1, 2).contiguous(),
                          mask=mask.flatten(1).contiguous(),
                          fine_pos=fine_pos.flatten(2).transpose(1, 2).contiguous(),
                          selected_objects_pos_embed=pos_embed, selected_centers=selected_centers,
                          bbox_embed=self._bbox_embed)
        cls_x, reg_x = torch.split(y, [self._hidden_dim, self._hidden_dim], dim=-1)      # model.py:120-131
        cls = self._cls_embed(cls_x)
        tmp = self._bbox_embed(reg_x)
        tmp[..., :2] += R.inverse_sigmoid(selected_centers)
        return {"pred_class": cls, "pred_boxes": tmp.sigmoid()}, enc


def make_criterion(num_cls: int, cost_class: float = 0.5, cost_ciou: float = 0.5):
    """train.py:253-262: SetCriterion(HungarianMatcherWoL1, {focal, L1, CIoU}) + the one-hot-width shim."""
    R = ref()

    class SetCriterionC(R.SetCriterion):
        def _get_class_loss(self, pred_logits, pred_idx, gt_class):  # criterion.py:29-49, one-hot width = logits width
            selected = pred_logits.index_select(index=pred_idx, dim=0)
            keep = torch.ones(pred_logits.size(0), dtype=torch.bool, device=pred_logits.device)
            keep[pred_idx] = False
            ordered = torch.concat([selected, pred_logits[keep]], dim=0)
            dummy = torch.ones((pred_logits.size(0) - pred_idx.size(0),), device=pred_logits.device).long()
            oh = nn.functional.one_hot(torch.concat([gt_class, dummy], dim=0), num_classes=pred_logits.size(1))
            return self._loss_fns["class"](ordered, oh, ordered.size(0))

    return SetCriterionC(num_classes=num_cls, matcher=R.HungarianMatcherWoL1(cost_class=cost_class, cost_ciou=cost_ciou),
                         loss_fn={"class": R.sigmoid_focal_loss, "bbox": nn.L1Loss(), "ciou": R.CompleteIOULoss()})


class RefTrainer:
    """One reference training step of the hot path: forward -> SetCriterion (matcher inside: `C.cpu()` + scipy,
    matcher.py:107-112) -> backward -> AdamW (train.py:164-178, 236-244), on `device`, optionally under
    `torch.autocast(bfloat16)` (the comparator SURVEY 8(d) names for config 2)."""

    def __init__(self, cfg: Dict, device="cpu", dropout: bool = True, autocast: bool = False, lr: float = 1e-5,
                 weights=None, optimizer: bool = True, seed: int = 0):
        R = ref()
        torch.manual_seed(seed)
        self.cfg, self.dev, self.autocast = cfg, torch.device(device), autocast
        args = Namespace(hidden_dim=256, num_encoder_blocks=cfg["L"], num_decoder_blocks=cfg["L"], num_cls=cfg["C"])
        self.model = RefTransformerHalf(args).to(self.dev).train()
        if not dropout:
            zero_dropout(self.model)
        self.crit = make_criterion(cfg["C"]).to(self.dev)
        self.w = weights or {"class": 0.5, "bbox": 0.0, "ciou": 0.5}      # arg_parser.py:41-61 defaults
        self.opt = torch.optim.AdamW(self.model.parameters(), lr=lr) if optimizer else None
        self.R = R

    def to_dev(self, batch):
        feats, mask, sel, centers, labels, boxes = batch
        d = self.dev
        tg = [{"labels": l.to(d), "boxes": b.to(d)} for l, b in zip(labels, boxes)]
        return feats.to(d), mask.to(d), sel.to(d), centers.to(d), tg

    def step(self, batch, on_device: bool = False) -> torch.Tensor:
        feats, mask, sel, centers, tg = batch if on_device else self.to_dev(batch)
        if self.opt is not None:
            self.opt.zero_grad()
        else:
            for p in self.model.parameters():
                p.grad = None
        with torch.autocast(self.dev.type, dtype=torch.bfloat16, enabled=self.autocast):
            out, _ = self.model(feats, mask, sel, centers)
        out = {k: v.float() for k, v in out.items()}
        loss = self.R.reduce_dict(self.crit(out, tg), weights=self.w)
        loss.sum().backward()
        if self.opt is not None:
            self.opt.step()
        return loss.detach()


def time_trainer(tr: RefTrainer, batches: List, steps: int, warmup: int, on_device: bool = False):
    """-> (images/s, s/step).  CUDA: events on the current stream around `steps` steps (loss read back each step,
    as a training loop logging its loss does: train.py:169-170); CPU: wall clock."""
    B = batches[0][0].shape[0]
    cuda = tr.dev.type == "cuda"
    if on_device:
        batches = [tr.to_dev(b) for b in batches]
    for s in range(warmup):
        float(tr.step(batches[s % len(batches)], on_device).sum())
    if cuda:
        torch.cuda.synchronize()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
    t0 = time.perf_counter()
    for s in range(steps):
        float(tr.step(batches[s % len(batches)], on_device).sum())
    if cuda:
        en.record()
        torch.cuda.synchronize()
        dt = st.elapsed_time(en) / 1e3
    else:
        dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps
