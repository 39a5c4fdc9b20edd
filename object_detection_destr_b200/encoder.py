"""Drop-in replacements for the reference encoder modules (src/model/blocks/encoder_block.py).

Same class names, constructor arguments, forward signatures, return shapes and parameter names
(so `state_dict()`s are interchangeable with the reference, including the dead `_proj_to_q/k/v`
parameters), but the forward runs on the B200 kernels: bf16 token-major activations, fused tcgen05
flash attention, fused residual+LayerNorm, cuBLAS bf16 GEMMs for the plain projections.
"""
from __future__ import annotations

import copy
from typing import Optional

import torch
from torch import nn

from . import functional as Fn
from . import ops

BF16 = torch.bfloat16


def _params(module: nn.Module):
    return dict(module.named_parameters())


class EncoderBlock(nn.Module):
    """reference: encoder_block.py:47-112."""

    def __init__(self, hidden_dim: int = 256, heads_num: int = 8, d_k: int = 256, d_v: int = 256):
        super().__init__()
        if hidden_dim != 256 or heads_num != 8:
            raise ValueError("the B200 encoder kernels are built for hidden_dim=256, 8 heads (d_head=32), "
                             "the only configuration the reference can run (encoder_block.py:17-22)")
        self.self_attn = nn.MultiheadAttention(embed_dim=hidden_dim, num_heads=heads_num, dropout=0.3,
                                               kdim=hidden_dim, vdim=hidden_dim)
        self.fc1 = nn.Linear(hidden_dim, 2048)
        self.fc2 = nn.Linear(2048, hidden_dim)
        self.dropout1 = nn.Dropout(0.3)
        self.dropout2 = nn.Dropout(0.3)
        self.dropout3 = nn.Dropout(0.3)
        self.norm1 = nn.LayerNorm(hidden_dim)
        self.norm2 = nn.LayerNorm(hidden_dim)
        self._heads_num = heads_num
        self._hidden_dim = hidden_dim
        # dead parameters of the reference (encoder_block.py:76-82), kept for state_dict parity
        self._proj_to_q = nn.Linear(hidden_dim, d_k, bias=False)
        self._proj_to_k = nn.Linear(hidden_dim, d_k, bias=False)
        self._proj_to_v = nn.Linear(d_k, d_v, bias=False)

    @property
    def heads_num(self):
        return self._heads_num

    def forward(self, inputs, mask: Optional[torch.Tensor] = None, key_mask: Optional[torch.Tensor] = None,
                pos_embed: Optional[torch.Tensor] = None):
        """inputs, pos_embed: (N, B, 256) seq-first as in the reference; key_mask (B, N) bool."""
        if mask is not None:
            raise NotImplementedError("attn_mask is never used by the reference encoder (encoder_block.py:35-39)")
        _check_dropout(self)
        N, B, d = inputs.shape
        x = inputs.transpose(0, 1).reshape(B * N, d).to(BF16)
        pos = pos_embed.transpose(0, 1).reshape(B * N, d).to(BF16)
        bits = ops.pack_key_mask(key_mask, B, N, device=inputs.device)
        p = {"blk." + k: v for k, v in self.named_parameters()}
        y = _block_only(x, pos, bits, p, "blk.", B, N)
        return y.view(B, N, d).transpose(0, 1).to(inputs.dtype)


def _block_only(x, pos, bits, p, lp, B, N):
    """EncoderBlock.forward proper (without the Encoder-level pos scaling / outer norm)."""
    xq = x + pos
    W, bias = p[lp + "self_attn.in_proj_weight"], p[lp + "self_attn.in_proj_bias"]
    d = x.shape[-1]
    qk = Fn.linear(xq, W[: 2 * d], bias[: 2 * d])
    v = Fn.linear(x, W[2 * d:], bias[2 * d:])
    a = Fn.enc_attn(qk, v, bits, B, N, 8)
    o = Fn.linear(a, p[lp + "self_attn.out_proj.weight"], p[lp + "self_attn.out_proj.bias"])
    x1 = Fn.add_layernorm(x, o, p[lp + "norm1.weight"], p[lp + "norm1.bias"])
    f = Fn.linear(torch.relu(Fn.linear(x1, p[lp + "fc1.weight"], p[lp + "fc1.bias"])), p[lp + "fc2.weight"],
                  p[lp + "fc2.bias"])
    return Fn.add_layernorm(x1, f, p[lp + "norm2.weight"], p[lp + "norm2.bias"])


def _check_dropout(module: nn.Module):
    if module.training:
        for m in module.modules():
            p_drop = m.p if isinstance(m, nn.Dropout) else (m.dropout if isinstance(m, nn.MultiheadAttention) else
                                                            getattr(m, "_dropout_prob", 0.0))
            if p_drop > 0:
                raise NotImplementedError(
                    "train-mode dropout is not implemented in the B200 kernels yet: set every nn.Dropout.p = 0 "
                    "(object_detection_destr_b200.disable_dropout(model)) or call .eval(); see DESIGN.md")


class Encoder(nn.Module):
    """reference: encoder_block.py:8-44.  `_pos_scale` and `norm` are shared by all layers."""

    def __init__(self, encoder_block: nn.Module, num_encoder_blocks: int = 6):
        super().__init__()
        self._encoder = nn.ModuleList([copy.deepcopy(encoder_block) for _ in range(num_encoder_blocks)])
        self._num_enc = num_encoder_blocks
        self._pos_scale = nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 256))
        self.norm = nn.LayerNorm(256)

    def forward_tokens(self, x: torch.Tensor, pos: torch.Tensor, bits: torch.Tensor, B: int, N: int):
        """Token-major fast path: x, pos bf16 [B*N, 256] -> bf16 [B*N, 256]."""
        _check_dropout(self)
        return Fn.encoder_tokens(x, pos, bits, _params(self), self._num_enc, B, N)

    def forward(self, inputs, mask, pos_embed):
        """inputs, pos_embed (B,256,H,W); mask (B,H,W) bool -> (B,256,H,W)  (reference signature)."""
        B, C, H, W = inputs.shape
        N = H * W
        x = inputs.flatten(2).transpose(1, 2).reshape(B * N, C).to(BF16)
        pos = pos_embed.flatten(2).transpose(1, 2).reshape(B * N, C).to(BF16)
        bits = ops.pack_key_mask(mask.flatten(1), B, N, device=inputs.device)
        y = self.forward_tokens(x, pos, bits, B, N)
        return y.view(B, N, C).transpose(1, 2).reshape(B, C, H, W).to(inputs.dtype)


def build_encoder(args):
    """reference: encoder_block.py:115-124."""
    return Encoder(encoder_block=EncoderBlock(d_k=args.hidden_dim, d_v=args.hidden_dim),
                   num_encoder_blocks=args.num_encoder_blocks)


def disable_dropout(model: nn.Module) -> nn.Module:
    """Set every dropout probability to 0 (nn.Dropout, nn.MultiheadAttention and the reference-style
    inline `_dropout_prob`), the configuration parity is defined on."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
        if hasattr(m, "_dropout_prob"):
            m._dropout_prob = 0.0
    return model
