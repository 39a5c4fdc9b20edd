"""The DESTR transformer half as one module: everything between the backbone's 1x1 `reduce_dim`
conv and the loss (src/model/model.py:84-131 minus the mini-detector), on the B200 kernels.

Parameter names follow the reference `ObjDetSplitTransformer` (`_encoder.*`, `_decoder.*`,
`_cls_embed.*`, `_bbox_embed.*`) so a reference checkpoint's matching entries load directly.
The mini-detector (query selection, out of scope) is replaced by explicit `selected_objects` /
`selected_centers` inputs, which is exactly what it hands to the decoder (model.py:100-118).
"""
from __future__ import annotations

from argparse import Namespace

import torch
import torch.nn.functional as F
from torch import nn

from . import functional as Fn
from . import ops
from .decoder import build_decoder
from .encoder import _params, build_encoder

BF16 = torch.bfloat16


def inverse_sigmoid(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """reference: misc.py:59-62."""
    return -torch.log(1.0 / x.clamp(min=eps) - 1.0)


class TransformerHalf(nn.Module):
    """`runtime=True` (default) executes encoder+decoder through the hand-scheduled HotPathRuntime
    (flat parameter buffers, explicit backward); `runtime=False` goes through the module-level autograd
    path of functional.py.  Both compute the same function (tests/test_gpu_decoder.py)."""

    def __init__(self, args: Namespace, runtime: bool = True):
        super().__init__()
        self.use_runtime = runtime
        self._rt = None
        self._encoder = build_encoder(args)
        self._decoder = build_decoder(args)
        d = args.hidden_dim
        self._hidden_dim = d
        self._cls_embed = nn.Linear(d, args.num_cls)
        self._bbox_embed = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Linear(d, 4))

    def runtime(self):
        """The hand-scheduled executor (created on first use, after the module sits on its device; it
        re-points the encoder/decoder parameters at views of its flat master buffer)."""
        if self._rt is None:
            from .runtime import HotPathRuntime
            self._rt = HotPathRuntime(self._encoder, self._decoder, self._bbox_embed, self._cls_embed.weight.device,
                                      cls_embed=self._cls_embed)
        return self._rt

    def make_optimizer(self, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        """AdamW for this module (call after .to(device)).  With the hand-scheduled runtime: one flat kernel for
        every parameter -- encoder, decoder and the prediction heads, which share the runtime's flat buffer --
        (+ bf16 shadow refresh); otherwise a
        plain fused torch.optim.AdamW.  Same arithmetic either way (tests/test_gpu_engine.py)."""
        if not self.use_runtime:
            return torch.optim.AdamW(self.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                     fused=True, capturable=True)
        from .runtime import FlatAdamW
        P = self.runtime().P
        lo, hi = P.m32.data_ptr(), P.m32.data_ptr() + P.m32.numel() * 4
        extra = [p for p in self.parameters() if not (lo <= p.data_ptr() < hi)]
        return FlatAdamW(P, extra, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)

    def set_dropout_seed(self, value: int):
        """Seed of the in-kernel dropout masks.  Runtime path: the device counter of the runtime (the engine bumps it
        every step); module-level path: pins the per-module seed (None releases it)."""
        if self.use_runtime:
            self.runtime().seed.fill_(int(value))
        else:
            from .encoder import set_dropout_seed
            set_dropout_seed(self, value)

    def after_optimizer_step(self):
        """Refresh the bf16 weight shadows (call after every optimizer step when runtime=True)."""
        if self._rt is not None:
            self._rt.P.refresh()

    def forward(self, features: torch.Tensor, mask: torch.Tensor, selected_objects: torch.Tensor,
                selected_centers: torch.Tensor, pairs_override=None, aux=None):
        """features (B,256,H,W) fp32 = reduce_dim(backbone) output; mask (B,H,W) bool;
        selected_objects (B,Q,512); selected_centers (B,Q,2) in (0,1).
        Returns {"pred_class": (B,Q,C) fp32, "pred_boxes": (B,Q,4) fp32} (model.py:120-131)."""
        B, C, H, W = features.shape
        N = H * W
        Q = selected_objects.shape[1]
        x = features.flatten(2).transpose(1, 2).reshape(B * N, C).to(BF16).contiguous()  # (B = 1: the reshape is a column-major view)
        _, pos = ops.sine_pos2d(mask, want_f32=False, want_bf16=True)  # K1 (position_encoding_cdetr.py:39-63)
        pos = pos.view(B * N, C)
        kpm = mask.flatten(1).contiguous()
        bits = ops.pack_key_mask(kpm, B, N, device=features.device)
        centers = selected_centers.reshape(B * Q, 2).float().contiguous()
        _, pos_embed = ops.query_sine_embed(centers, want_f32=False, want_bf16=True)  # model.py:104-106
        sel = selected_objects.reshape(B * Q, 512).to(BF16)
        if self.use_runtime:
            from .runtime import _RuntimeFn
            rt = self.runtime()  # dropout (incl. the decoder's always-on attention dropout) is applied in the kernels
            dec, enc = _RuntimeFn.apply(rt, rt.anchor, x, pos, bits, kpm, sel, pos_embed, pos_embed, centers, B, N, Q,
                                        pairs_override, aux)
        else:
            from .encoder import drop_scope
            with drop_scope(self, features.device):  # one seed for the whole path (sites differ per layer / call)
                enc = Fn.encoder_tokens(x, pos, bits, _params(self._encoder), len(self._encoder._encoder), B, N)
                # fine_pos = pos * encoder._pos_scale(enc_out)  (model.py:89-92)
                fine_pos = Fn._MulConst.apply(Fn.mlp2(enc, _params(self._encoder), "_pos_scale."), pos)
                dec = Fn.decoder_tokens(sel, enc, bits, kpm, fine_pos, pos_embed, centers, _params(self._decoder),
                                        _params(self._bbox_embed), len(self._decoder._decoder), B, Q, N,
                                        pairs_override=pairs_override, aux=aux)
        # class + box heads, fp32 as in the reference (model.py:120-131): one kernel (csrc/heads.cu)
        hp = (self._cls_embed.weight, self._cls_embed.bias, self._bbox_embed[0].weight, self._bbox_embed[0].bias,
              self._bbox_embed[2].weight, self._bbox_embed[2].bias)
        # with the runtime the head parameters live in its flat buffer: the backward kernel writes their gradients
        # straight into the flat gradient buffer (.grad are views of it)
        grad_out = None
        if self.use_runtime and self._rt.heads_in_flat and torch.is_grad_enabled():
            grad_out = tuple(self._rt.P.g("h." + n) for n in ("cls_w", "cls_b", "box0_w", "box0_b", "box2_w", "box2_b"))
        cls, boxes = ops.heads(dec, centers, *hp, grad_out=grad_out,
                               off_path=self._rt.P.off_path if grad_out is not None else None)
        return {"pred_class": cls.view(B, Q, -1), "pred_boxes": boxes.view(B, Q, 4)}, enc.view(B, N, C)
