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


class EncoderMHA(nn.MultiheadAttention):
    """`EncoderBlock.self_attn` (encoder_block.py:57-63, called at :97-103) as a swap point of its own: the same
    parameters and call signature as nn.MultiheadAttention, but forward runs the fused tcgen05 attention instead of
    torch's math path.  Supports what the reference calls it with: seq-first (N, B, 256) inputs, query is key,
    a boolean key_padding_mask, no attn_mask; the attention weights are not materialised (the reference discards
    them) -> returns (out, None)."""

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None,
                average_attn_weights=True, is_causal=False):
        if attn_mask is not None or is_causal:
            raise NotImplementedError("attn_mask is never used by the reference encoder (encoder_block.py:35-39)")
        if self.embed_dim != 256 or self.num_heads != 8 or query.dim() != 3 or self.batch_first:
            raise NotImplementedError("fused encoder attention: (N, B, 256) seq-first inputs, 8 heads x 32")
        if not query.is_cuda:
            raise RuntimeError("destr_b200 has no CPU path")
        N, B, d = query.shape
        tok = lambda t: t.transpose(0, 1).reshape(B * N, d).to(BF16)
        W, bias = self.in_proj_weight, self.in_proj_bias
        xq, xk, xv = tok(query), (tok(key) if key is not query else None), tok(value)
        if xk is None:
            qk = Fn.linear(xq, W[:2 * d], bias[:2 * d])
        else:
            qk = torch.cat([Fn.linear(xq, W[:d], bias[:d]), Fn.linear(xk, W[d:2 * d], bias[d:2 * d])], dim=-1)
        v = Fn.linear(xv, W[2 * d:], bias[2 * d:])
        bits = ops.pack_key_mask(key_padding_mask, B, N, device=query.device)
        with drop_scope(self, query.device):
            a = Fn.enc_attn(qk, v, bits, B, N, 8, drop=Fn._dr("e.attn", Fn._enc_site(0, "attn")))
        o = Fn.linear(a, self.out_proj.weight, self.out_proj.bias)
        return o.view(B, N, d).transpose(0, 1).to(query.dtype), None


class EncoderBlock(nn.Module):
    """reference: encoder_block.py:47-112."""

    def __init__(self, hidden_dim: int = 256, heads_num: int = 8, d_k: int = 256, d_v: int = 256):
        super().__init__()
        if hidden_dim != 256 or heads_num != 8:
            raise ValueError("the B200 encoder kernels are built for hidden_dim=256, 8 heads (d_head=32), "
                             "the only configuration the reference can run (encoder_block.py:17-22)")
        self.self_attn = EncoderMHA(embed_dim=hidden_dim, num_heads=heads_num, dropout=0.3, kdim=hidden_dim,
                                    vdim=hidden_dim)
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
        N, B, d = inputs.shape
        x = inputs.transpose(0, 1).reshape(B * N, d).to(BF16)
        pos = pos_embed.transpose(0, 1).reshape(B * N, d).to(BF16)
        bits = ops.pack_key_mask(key_mask, B, N, device=inputs.device)
        p = {"blk." + k: v for k, v in self.named_parameters()}
        with drop_scope(self, inputs.device):
            y = _block_only(x, pos, bits, p, "blk.", B, N)
        return y.view(B, N, d).transpose(0, 1).to(inputs.dtype)


def _block_only(x, pos, bits, p, lp, B, N):
    """EncoderBlock.forward proper (without the Encoder-level pos scaling / outer norm)."""
    xq = x + pos
    W, bias = p[lp + "self_attn.in_proj_weight"], p[lp + "self_attn.in_proj_bias"]
    d = x.shape[-1]
    qk = Fn.linear(xq, W[: 2 * d], bias[: 2 * d])
    v = Fn.linear(x, W[2 * d:], bias[2 * d:])
    S = Fn._enc_site
    a = Fn.enc_attn(qk, v, bits, B, N, 8, drop=Fn._dr("e.attn", S(0, "attn")))
    o = Fn.linear(a, p[lp + "self_attn.out_proj.weight"], p[lp + "self_attn.out_proj.bias"])
    x1 = Fn.add_layernorm(x, o, p[lp + "norm1.weight"], p[lp + "norm1.bias"], Fn._dr("e.d1", S(0, "d1")))
    h = Fn.dropout(torch.relu(Fn.linear(x1, p[lp + "fc1.weight"], p[lp + "fc1.bias"])), Fn._dr("e.d2", S(0, "d2")))
    f = Fn.linear(h, p[lp + "fc2.weight"], p[lp + "fc2.bias"])
    return Fn.add_layernorm(x1, f, p[lp + "norm2.weight"], p[lp + "norm2.bias"], Fn._dr("e.d3", S(0, "d3")))


_SEED_STREAMS = 0


def initial_dropout_seed() -> int:
    """Start value of a dropout seed counter: derived from torch's generator seed (so `torch.manual_seed` controls
    it, as it controls the reference's nn.Dropout), the data-parallel rank (independent masks per worker) and a
    per-process stream number (separately constructed modules do not replay each other's masks)."""
    global _SEED_STREAMS
    _SEED_STREAMS += 1
    import os
    rank = int(os.environ.get("RANK", "0"))
    x = (torch.initial_seed() * 0x9E3779B97F4A7C15 + rank * 0xD1B54A32D192ED03 + _SEED_STREAMS * 0x94D049BB133111EB)
    x &= (1 << 64) - 1
    x ^= x >> 31
    return int(x & 0x3FFFFFFF)


def drop_scope(module: nn.Module, device, seed: Optional[torch.Tensor] = None):
    """Context that installs the dropout in force for `module`'s forward (functional.dropout_ctx): the probabilities
    are read from the module tree the way the reference applies them -- nn.Dropout and nn.MultiheadAttention(dropout=)
    only in training mode, the SelfAttention attention-probability dropout ALWAYS (it builds its nn.Dropout inline,
    self_attention.py:40).  The seed is a per-module device counter (not a registered buffer: state_dict keys stay the
    reference's), bumped on every forward; `set_dropout_seed(module, v)` pins it."""
    t = ops.drop_thr16
    thr = {}
    tr = module.training
    for m in module.modules():
        cls = type(m).__name__
        if isinstance(m, nn.MultiheadAttention):
            thr["e.attn"] = t(m.dropout) if tr else 0
        elif cls == "EncoderBlock":
            thr.update({"e.d1": t(m.dropout1.p) if tr else 0, "e.d2": t(m.dropout2.p) if tr else 0,
                        "e.d3": t(m.dropout3.p) if tr else 0})
        elif cls == "DecoderBlock":
            thr.update({"d.sa": t(m._self_attn._dropout_prob), "d.d1": t(m.dropout1.p) if tr else 0})
        elif cls == "ClsRegBranch":
            thr.update({"d.ca": t(m.cross_attn._dropout_prob), "d.br": t(m.dropout.p) if tr else 0})
        elif cls == "SelfAttention" and m is module:  # stand-alone use
            thr.update({"d.sa": t(m._dropout_prob), "d.ca": t(m._dropout_prob)})
    if seed is None:
        seed = getattr(module, "_drop_seed_t", None)
        if seed is None or seed.device != torch.device(device):
            seed = torch.full((1,), initial_dropout_seed(), dtype=torch.int32, device=device)
            object.__setattr__(module, "_drop_seed_t", seed)
        if any(thr.values()) and not getattr(module, "_drop_seed_pinned", False):
            seed.add_(1)
    return Fn.dropout_ctx(seed, thr)


def set_dropout_seed(module: nn.Module, value: int, device=None):
    """Pin the dropout seed of a drop-in module (tests, reproducibility): no automatic bump until reset with None."""
    dev = device or next(module.parameters()).device
    if value is None:
        object.__setattr__(module, "_drop_seed_pinned", False)
        return
    object.__setattr__(module, "_drop_seed_t", torch.full((1,), int(value), dtype=torch.int32, device=dev))
    object.__setattr__(module, "_drop_seed_pinned", True)


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
        with drop_scope(self, x.device):
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
