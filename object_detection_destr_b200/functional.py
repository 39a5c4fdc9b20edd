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
    def forward(ctx, a, b, gamma, beta):
        y, mean, rstd = ops.add_layernorm(a, b, gamma, beta, save_stats=True)
        ctx.save_for_backward(a, b, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, b, gamma, mean, rstd = ctx.saved_tensors
        dx, dg, db = ops.add_layernorm_bwd(dy, a, b, gamma, mean, rstd)
        return dx, dx, dg, db


class _EncAttn(torch.autograd.Function):
    """Fused encoder attention; qk = [q | k] projection output, v projection output."""

    @staticmethod
    def forward(ctx, qk, v, bits, B, N, heads, scale):
        C = heads * 32
        out, lse = ops.enc_attn_fwd(qk[:, :C], qk[:, C:], v, bits, B, N, heads, scale)
        ctx.save_for_backward(qk, v, bits, out, lse)
        ctx.dims = (B, N, heads, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        qk, v, bits, out, lse = ctx.saved_tensors
        B, N, heads, scale = ctx.dims
        C = heads * 32
        dqk, dv = ops.enc_attn_bwd(qk[:, :C], qk[:, C:], v, bits, out, dout, lse, B, N, heads, scale)
        return dqk, dv, None, None, None, None, None


def pos_mul_add(x, pos, s):
    return _PosMulAdd.apply(x, pos, s)


def add_layernorm(a, b, gamma, beta):
    return _AddLayerNorm.apply(a, b, gamma, beta)


def enc_attn(qk, v, bits, B, N, heads=8):
    return _EncAttn.apply(qk, v, bits, B, N, heads, 1.0 / math.sqrt(32))


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
                  heads: int = 8) -> Tensor:
    """One `x = norm(x + EncoderBlock(x, pos*pos_scale(x)))` step (encoder_block.py:33-40, 88-112).
    x, pos: bf16 [B*N, 256]."""
    s = mlp2(x, p, "_pos_scale.")
    xq = pos_mul_add(x, pos, s)
    W, bias = p[lp + "self_attn.in_proj_weight"], p[lp + "self_attn.in_proj_bias"]
    d = x.shape[-1]
    qk = linear(xq, W[: 2 * d], bias[: 2 * d])
    v = linear(x, W[2 * d:], bias[2 * d:])
    a = enc_attn(qk, v, bits, B, N, heads)
    o = linear(a, p[lp + "self_attn.out_proj.weight"], p[lp + "self_attn.out_proj.bias"])
    x1 = add_layernorm(x, o, p[lp + "norm1.weight"], p[lp + "norm1.bias"])
    f = linear(torch.relu(linear(x1, p[lp + "fc1.weight"], p[lp + "fc1.bias"])), p[lp + "fc2.weight"],
               p[lp + "fc2.bias"])
    x2 = add_layernorm(x1, f, p[lp + "norm2.weight"], p[lp + "norm2.bias"])
    return add_layernorm(x, x2, p["norm.weight"], p["norm.bias"])


def encoder_tokens(x: Tensor, pos: Tensor, bits: Tensor, p: Dict[str, Tensor], num_layers: int, B: int, N: int):
    """Encoder.forward on token-major bf16 activations."""
    for l in range(num_layers):
        x = encoder_layer(x, pos, bits, p, f"_encoder.{l}.", B, N)
    return x
