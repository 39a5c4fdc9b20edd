"""Autograd-aware functional layer over the C-ABI kernels.

Token-major bf16 activations ([rows, channels]); fp32 master parameters are cast to bf16 for the
tensor-core GEMMs (cuBLAS via torch, a plain library GEMM) while every non-GEMM op on the hot path
is one of our kernels with a hand-written backward.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ops

Tensor = torch.Tensor
BF16 = torch.bfloat16


# ----------------------------------------------------------------------------------------------
# dropout context: the modules install the probabilities in force (as 16-bit thresholds, keys as in
# runtime.HotPathRuntime._drop_config) and a device seed; the functions below turn a (kind, site) into the `drop`
# argument of the kernels.  No context / threshold 0 -> no dropout.
# ----------------------------------------------------------------------------------------------
_CTX = None


class dropout_ctx:
    def __init__(self, seed: Optional[Tensor], thr: Optional[Dict[str, int]]):
        self.thr = thr or {}
        # every autograd Function of this forward keeps (seed tensor, thr, site) for its backward, and the kernels read
        # the seed from DEVICE memory: it must be this forward's own copy -- the module's live counter advances on the
        # next forward (gradient accumulation, a second view, an eval pass before backward ...)
        self.seed = seed.clone() if (seed is not None and any(self.thr.values())) else seed

    def __enter__(self):
        global _CTX
        self._prev, _CTX = _CTX, (self if any(self.thr.values()) else None)
        return self

    def __exit__(self, *exc):
        global _CTX
        _CTX = self._prev


def _dr(kind: str, site: int, site2: Optional[int] = None):
    if _CTX is None or not _CTX.thr.get(kind, 0):
        return None
    t = _CTX.thr[kind]
    return (_CTX.seed, t, site) if site2 is None else (_CTX.seed, t, site, site2)


def _enc_site(layer: int, name: str) -> int:
    from .runtime import enc_site
    return enc_site(layer, name)


def _dec_site(layer: int, name: str) -> int:
    from .runtime import dec_site
    return dec_site(layer, name)


class _PosMulAdd(torch.autograd.Function):
    """y = x + pos * s   (encoder_block.py:38,95)."""

    @staticmethod
    def forward(ctx, x, pos, s):
        ctx.save_for_backward(pos)
        return ops.pos_mul_add(x, pos, s)

    @staticmethod
    def backward(ctx, dy):
        (pos,) = ctx.saved_tensors
        dy = dy.contiguous()
        return dy, None, ops.pos_mul_add_bwd(dy, pos)


class _AddLayerNorm(torch.autograd.Function):
    """y = LayerNorm(a + b)."""

    @staticmethod
    def forward(ctx, a, b, gamma, beta, drop=None):
        y, mean, rstd = ops.add_layernorm(a, b, gamma, beta, save_stats=True, drop=drop)
        ctx.save_for_backward(a, b, gamma, mean, rstd)
        ctx.drop = drop
        return y

    @staticmethod
    def backward(ctx, dy):
        a, b, gamma, mean, rstd = ctx.saved_tensors
        if ctx.drop is None:
            dx, dg, db = ops.add_layernorm_bwd(dy, a, b, gamma, mean, rstd)
            return dx, dx, dg, db, None
        # y = LN(a + dropout(b)): d(b) goes through the mask, d(a) does not
        dxb, dg, db, dsum = ops.add_layernorm_bwd(dy, a, b, gamma, mean, rstd, drop=ctx.drop, want_sum=True)
        return dsum, dxb, dg, db, None


class _Dropout(torch.autograd.Function):
    """nn.Dropout on a bf16 [M,C] activation with the kernels' counter-based mask (the mask is linear: the backward
    applies the same mask to the gradient)."""

    @staticmethod
    def forward(ctx, x, drop):
        ctx.drop = drop
        return ops.dropout_inplace(x.contiguous().clone(), drop)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout_inplace(dy.contiguous().clone(), ctx.drop), None


class _EncAttn(torch.autograd.Function):
    """Fused encoder attention; qk = [q | k] projection output, v projection output."""

    @staticmethod
    def forward(ctx, qk, v, bits, B, N, heads, scale, drop=None):
        C = heads * 32
        rb = cb = None
        if drop is not None and drop[1]:  # both orientations of the mask now: the seed may move before backward runs
            rb, cb = ops.attn_dropout_bits(drop, B * heads, N, qk.device)
        out, lse = ops.enc_attn_fwd(qk[:, :C], qk[:, C:], v, bits, B, N, heads, scale, drop=drop, rowbits=rb)
        ctx.colbits = cb
        ctx.save_for_backward(qk, v, bits, out, lse)
        ctx.dims = (B, N, heads, scale)
        ctx.drop = drop
        return out

    @staticmethod
    def backward(ctx, dout):
        qk, v, bits, out, lse = ctx.saved_tensors
        B, N, heads, scale = ctx.dims
        C = heads * 32
        dqk, dv = ops.enc_attn_bwd(qk[:, :C], qk[:, C:], v, bits, out, dout, lse, B, N, heads, scale, drop=ctx.drop,
                                   colbits=ctx.colbits)
        return dqk, dv, None, None, None, None, None, None


def pos_mul_add(x, pos, s):
    return _PosMulAdd.apply(x, pos, s)


def add_layernorm(a, b, gamma, beta, drop=None):
    return _AddLayerNorm.apply(a, b, gamma, beta, drop)


def dropout(x, drop):
    return x if drop is None else _Dropout.apply(x, drop)


def enc_attn(qk, v, bits, B, N, heads=8, drop=None):
    return _EncAttn.apply(qk, v, bits, B, N, heads, 1.0 / math.sqrt(32), drop)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    """bf16 tensor-core GEMM (cuBLAS) with fp32 master weights."""
    return F.linear(x, w.to(BF16), None if b is None else b.to(BF16))


def mlp2(x: Tensor, p: Dict[str, Tensor], prefix: str) -> Tensor:
    return linear(torch.relu(linear(x, p[prefix + "0.weight"], p[prefix + "0.bias"])),
                  p[prefix + "2.weight"], p[prefix + "2.bias"])


# ----------------------------------------------------------------------------------------------
# encoder  (reference: src/model/blocks/encoder_block.py)
# ----------------------------------------------------------------------------------------------
def encoder_layer(x: Tensor, pos: Tensor, bits: Tensor, p: Dict[str, Tensor], lp: str, B: int, N: int,
                  heads: int = 8, layer: int = 0) -> Tensor:
    """One `x = norm(x + EncoderBlock(x, pos*pos_scale(x)))` step (encoder_block.py:33-40, 88-112).
    x, pos: bf16 [B*N, 256]."""
    s = mlp2(x, p, "_pos_scale.")
    xq = pos_mul_add(x, pos, s)
    W, bias = p[lp + "self_attn.in_proj_weight"], p[lp + "self_attn.in_proj_bias"]
    d = x.shape[-1]
    qk = linear(xq, W[: 2 * d], bias[: 2 * d])
    v = linear(x, W[2 * d:], bias[2 * d:])
    a = enc_attn(qk, v, bits, B, N, heads, drop=_dr("e.attn", _enc_site(layer, "attn")))
    o = linear(a, p[lp + "self_attn.out_proj.weight"], p[lp + "self_attn.out_proj.bias"])
    x1 = add_layernorm(x, o, p[lp + "norm1.weight"], p[lp + "norm1.bias"], _dr("e.d1", _enc_site(layer, "d1")))
    h = dropout(torch.relu(linear(x1, p[lp + "fc1.weight"], p[lp + "fc1.bias"])), _dr("e.d2", _enc_site(layer, "d2")))
    f = linear(h, p[lp + "fc2.weight"], p[lp + "fc2.bias"])
    x2 = add_layernorm(x1, f, p[lp + "norm2.weight"], p[lp + "norm2.bias"], _dr("e.d3", _enc_site(layer, "d3")))
    return add_layernorm(x, x2, p["norm.weight"], p["norm.bias"])


def encoder_tokens(x: Tensor, pos: Tensor, bits: Tensor, p: Dict[str, Tensor], num_layers: int, B: int, N: int):
    """Encoder.forward on token-major bf16 activations."""
    for l in range(num_layers):
        x = encoder_layer(x, pos, bits, p, f"_encoder.{l}.", B, N, layer=l)
    return x


# ----------------------------------------------------------------------------------------------
# decoder  (reference: src/model/blocks/decoder_block.py)
# ----------------------------------------------------------------------------------------------
class _MulConst(torch.autograd.Function):
    """y = a * c with c constant (sin_embed = sine * pos_scale(x_reg), decoder_block.py:49)."""

    @staticmethod
    def forward(ctx, a, c):
        ctx.save_for_backward(c)
        return ops.mul(a, c)

    @staticmethod
    def backward(ctx, dy):
        (c,) = ctx.saved_tensors
        return ops.mul(dy.contiguous(), c), None


class _DualLnMix(torch.autograd.Function):
    """lam*LN1(x+o1) + (1-lam)*LN2(x+o2eff) with the pair-attention slot masking fused."""

    @staticmethod
    def forward(ctx, x, o1, o2, pairs, g1, b1, g2, b2, lam, Q, drop=None):
        out, stats = ops.dual_ln_mix(x, o1, o2, pairs, g1, b1, g2, b2, lam, Q, drop=drop)
        ctx.save_for_backward(x, o1, o2, pairs, g1, g2, stats)
        ctx.lam, ctx.Q, ctx.drop = lam, Q, drop
        return out

    @staticmethod
    def backward(ctx, dout):
        x, o1, o2, pairs, g1, g2, stats = ctx.saved_tensors
        dx, do1, do2, dg1, db1, dg2, db2 = ops.dual_ln_mix_bwd(dout, x, o1, o2, pairs, g1, g2, stats, ctx.lam, ctx.Q,
                                                               drop=ctx.drop)
        return dx, do1, do2, None, dg1, db1, dg2, db2, None, None, None


class _DecQkvPrep(torch.autograd.Function):
    """q/k position add + left/right pair gathers (decoder_block.py:167-177, pair_self_attention.py:47-89)."""

    @staticmethod
    def forward(ctx, qkv_obj, qk_pos, pairs, B, Q):
        qkv, cat = ops.dec_qkv_prep(qkv_obj, qk_pos, pairs, B, Q)
        ctx.save_for_backward(pairs)
        ctx.dims = (B, Q)
        return qkv, cat

    @staticmethod
    def backward(ctx, d_qkv, d_cat):
        (pairs,) = ctx.saved_tensors
        B, Q = ctx.dims
        return ops.dec_qkv_prep_bwd(d_qkv.contiguous(), d_cat.contiguous(), pairs, B, Q) + (None, None, None)


class _DecSelfPairAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, cat, B, Q, drop=None):
        o1, o2, lse1, lse2 = ops.dec_self_pair_attn_fwd(qkv, cat, B, Q, drop=drop)
        ctx.save_for_backward(qkv, cat, o1, o2, lse1, lse2)
        ctx.dims = (B, Q)
        ctx.drop = drop
        return o1, o2

    @staticmethod
    def backward(ctx, do1, do2):
        qkv, cat, o1, o2, lse1, lse2 = ctx.saved_tensors
        B, Q = ctx.dims
        hm = lambda t, d: t.reshape(B, Q, 8, d).transpose(1, 2).contiguous()  # token-major -> head-major
        d1, d2 = hm(do1, 64), hm(do2, 128)
        delta1 = (d1.float() * hm(o1, 64).float()).sum(-1)
        delta2 = (d2.float() * hm(o2, 128).float()).sum(-1)
        d_qkv, d_cat = ops.dec_self_pair_attn_bwd(qkv, cat, d1, d2, lse1, lse2, delta1, delta2, B, Q, drop=ctx.drop)
        return d_qkv, d_cat, None, None, None


class _SplitCrossAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_obj, q_pos, k_enc, k_pos, v, bits, kpm, B, Q, N, drop=None):
        out, lse = ops.split_cross_attn_fwd(q_obj, q_pos, k_enc, k_pos, v, bits, B, Q, N, drop=drop)
        ctx.save_for_backward(q_obj, q_pos, k_enc, k_pos, v, bits, out, lse)
        ctx.dims = (B, Q, N)
        ctx.drop = drop
        return out

    @staticmethod
    def backward(ctx, dout):
        q_obj, q_pos, k_enc, k_pos, v, bits, out, lse = ctx.saved_tensors
        B, Q, N = ctx.dims
        g = ops.split_cross_attn_bwd(q_obj, q_pos, k_enc, k_pos, v, bits, out, dout.contiguous(), lse, B, Q, N,
                                     drop=ctx.drop)
        return g + (None, None, None, None, None, None)


def decoder_hoisted_projections(enc_out: Tensor, fine_pos: Tensor, pos_embed: Tensor, p: Dict[str, Tensor],
                                num_layers: int):
    """The projections of decoder_block.py:167-177,189-193 that depend only on layer weights and
    layer-invariant inputs, for ALL layers as three packed GEMMs:
      kv_all   [B*N, L*512]: per layer [W_k_enc enc | W_v_enc enc]
      kpos_all [B*N, L*256]: per layer  W_k_pos fine_pos
      qkpos_all[B*Q, L*512]: per layer [W_sa_q_pos pos | W_sa_k_pos pos]"""
    wkv = torch.cat([torch.cat([p[f"_decoder.{l}._ca_proj_to_k_enc.weight"], p[f"_decoder.{l}._ca_proj_to_v_enc.weight"]])
                     for l in range(num_layers)])
    wkp = torch.cat([p[f"_decoder.{l}._ca_proj_to_k_pos.weight"] for l in range(num_layers)])
    wqk = torch.cat([torch.cat([p[f"_decoder.{l}._sa_proj_to_q_pos.weight"], p[f"_decoder.{l}._sa_proj_to_k_pos.weight"]])
                     for l in range(num_layers)])
    return linear(enc_out, wkv), linear(fine_pos, wkp), linear(pos_embed, wqk)


def decoder_block_core(x: Tensor, sin_embed: Tensor, pairs: Tensor, qk_pos: Tensor, k_enc: Tensor, k_pos: Tensor,
                       v: Tensor, bits: Tensor, kpm: Optional[Tensor], p: Dict[str, Tensor], lp: str, B: int, Q: int,
                       N: int, lam: float = 0.5, layer: int = 0) -> Tensor:
    """DecoderBlock.forward (decoder_block.py:157-220) given the projected keys/values.
    x bf16 [B*Q,512]; sin_embed bf16 [B*Q,256]; pairs int32 [B,Q,2]; qk_pos bf16 [B*Q,512] = [W_q_pos p | W_k_pos p];
    k_enc, k_pos, v bf16 [B*N,256] views.  Returns cat(cls, reg) bf16 [B*Q,512]."""
    wqkv = torch.cat([p[lp + "_sa_proj_to_q_obj.weight"], p[lp + "_sa_proj_to_k_obj.weight"],
                      p[lp + "_sa_proj_to_v_obj.weight"]])
    qkv_obj = linear(x, wqkv)
    qkv, cat = _DecQkvPrep.apply(qkv_obj, qk_pos, pairs, B, Q)
    o1, o2 = _DecSelfPairAttn.apply(qkv, cat, B, Q, _dr("d.sa", _dec_site(layer, "sa")))
    o = _DualLnMix.apply(x, o1, o2, pairs, p[lp + "norm1.weight"], p[lp + "norm1.bias"], p[lp + "norm2.weight"],
                         p[lp + "norm2.bias"], lam, Q, _dr("d.d1", _dec_site(layer, "d1a"), _dec_site(layer, "d1b")))
    q_obj = linear(o, p[lp + "_ca_proj_to_q_obj.weight"])
    q_pos = linear(sin_embed, p[lp + "_ca_proj_to_q_pos.weight"])
    ca = _SplitCrossAttn.apply(q_obj, q_pos, k_enc, k_pos, v, bits, kpm, B, Q, N, _dr("d.ca", _dec_site(layer, "ca")))
    outs = []
    for i, br in enumerate(("_cls_branch.", "_reg_branch.")):
        bp = lp + br
        xb = add_layernorm(o[:, i * 256:(i + 1) * 256], ca[:, i * 256:(i + 1) * 256], p[bp + "norm1.weight"],
                           p[bp + "norm1.bias"], _dr("d.br", _dec_site(layer, f"b{i}.d_ca")))
        h = dropout(torch.relu(linear(xb, p[bp + "fc1.weight"], p[bp + "fc1.bias"])),
                    _dr("d.br", _dec_site(layer, f"b{i}.d_relu")))
        f = linear(h, p[bp + "fc2.weight"], p[bp + "fc2.bias"])
        outs.append(add_layernorm(xb, f, p[bp + "norm2.weight"], p[bp + "norm2.bias"],
                                  _dr("d.br", _dec_site(layer, f"b{i}.d_fc2"))))
    return torch.cat(outs, dim=-1)


def decoder_layer(x: Tensor, l: int, kv_all: Tensor, kpos_all: Tensor, qkpos_all: Tensor, sine: Tensor,
                  centers: Tensor, bits: Tensor, kpm: Optional[Tensor], p: Dict[str, Tensor],
                  bbox_p: Dict[str, Tensor], B: int, Q: int, N: int, lam: float = 0.5,
                  pairs_override: Optional[Tensor] = None):
    """x = norm(x + DecoderBlock(x, ...)) (decoder_block.py:43-65).  x bf16 [B*Q, 512].
    `pairs_override` (int32 [B,Q,2]) injects the discrete pairing (parity tests)."""
    x_reg = x[:, 256:]
    sin_embed = _MulConst.apply(mlp2(x_reg, p, "_pos_scale."), sine)
    with torch.no_grad():  # coords only select the pairing (an argmax): fp32, no gradient
        xr32 = x_reg.float()
        delta = F.linear(torch.relu(F.linear(xr32, bbox_p["0.weight"], bbox_p["0.bias"])), bbox_p["2.weight"],
                         bbox_p["2.bias"])
        coords = ops.box_refine(delta, centers)
        pairs = ops.pair_indices(coords.view(B, Q, 4)) if pairs_override is None else pairs_override
    y = decoder_block_core(x, sin_embed, pairs, qkpos_all[:, l * 512:(l + 1) * 512],
                           kv_all[:, l * 512:l * 512 + 256], kpos_all[:, l * 256:(l + 1) * 256],
                           kv_all[:, l * 512 + 256:(l + 1) * 512], bits, kpm, p, f"_decoder.{l}.", B, Q, N, lam, layer=l)
    return add_layernorm(x, y, p["norm.weight"], p["norm.bias"]), (coords, pairs)


def decoder_tokens(x: Tensor, enc_out: Tensor, bits: Tensor, kpm: Tensor, fine_pos: Tensor, pos_embed: Tensor,
                   centers: Tensor, p: Dict[str, Tensor], bbox_p: Dict[str, Tensor], num_layers: int, B: int, Q: int,
                   N: int, pairs_override=None, aux: Optional[list] = None):
    """Decoder.forward (decoder_block.py:28-67) on token-major bf16 activations.
    x [B*Q,512], enc_out/fine_pos [B*N,256], pos_embed [B*Q,256] bf16; centers fp32 [B*Q,2]."""
    kv_all, kpos_all, qkpos_all = decoder_hoisted_projections(enc_out, fine_pos, pos_embed, p, num_layers)
    _, sine = ops.query_sine_embed(centers, want_f32=False, want_bf16=True)  # loop-invariant (decoder_block.py:45-47)
    for l in range(num_layers):
        x, a = decoder_layer(x, l, kv_all, kpos_all, qkpos_all, sine, centers, bits, kpm, p, bbox_p, B, Q, N,
                             pairs_override=None if pairs_override is None else pairs_override[l])
        if aux is not None:
            aux.append(a)
    return x
