"""Backward passes of the three DECODER attention ops, composed from cuBLAS batched GEMMs plus
elementwise torch ops on the GPU.

Status (round 1): the decoder score matrices are small ([B,8,Q,Q] and [B,2,Q,N]; 0.6M / 1.7M
elements at config 2), so unlike the encoder (whose N x N scores must never be materialised and
whose backward is the fused tcgen05 kernel in csrc/enc_attn_bwd.cu) their backward is written as
plain library GEMMs around a recomputed P = 2^(c*S - lse).  Fused tcgen05 backward kernels for these
ops are the next step ("what comes next"); the forward of all three IS a fused kernel.
Everything here runs on the GPU; there is no CPU path.
"""
from __future__ import annotations

import math

import torch

LOG2E = 1.4426950408889634
BF16 = torch.bfloat16


def _bmm_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16 x bf16 -> fp32 batched GEMM (fp32 accumulate, fp32 result)."""
    sh = a.shape[:-2]
    a3, b3 = a.reshape(-1, *a.shape[-2:]), b.reshape(-1, *b.shape[-2:])
    try:
        c = torch.bmm(a3, b3, out_dtype=torch.float32)
    except (TypeError, RuntimeError, NotImplementedError):
        c = torch.bmm(a3.float(), b3.float())
    return c.view(*sh, a.shape[-2], b.shape[-1])


def _attn_bwd_core(q, k, v, do, lse, delta, c_log2, p_scale, ds_scale, key_valid=None):
    """q [..,Sq,dk], k [..,Sk,dk], v [..,Sk,dv], do [..,Sq,dv] bf16; lse/delta fp32 [..,Sq].
    P = p_scale * 2^(c_log2*q.k^T - lse);  O = P.v.   Returns dq, dk, dv (fp32)."""
    s = _bmm_f32(q, k.transpose(-1, -2))
    p = torch.exp2(s * c_log2 - lse[..., None]) * p_scale
    if key_valid is not None:
        p = p * key_valid
    dp = _bmm_f32(do, v.transpose(-1, -2))
    ds = (p * (dp - delta[..., None]) * ds_scale).to(BF16)
    pb = p.to(BF16)
    dv = _bmm_f32(pb.transpose(-1, -2), do)
    dq = _bmm_f32(ds, k)
    dk = _bmm_f32(ds.transpose(-1, -2), q)
    return dq, dk, dv


def dec_self_pair_attn_bwd(qkv, cat, o1, o2, do1, do2, lse1, lse2, B, Q):
    """qkv [3,B,8,Q,64], cat [3,B,8,Q,128] head-major; o/do token-major.
    -> head-major (d_qkv bf16 [3,B,8,Q,64], d_cat bf16 [3,B,8,Q,128])."""
    h = lambda t, d: t.reshape(B, Q, 8, d).transpose(1, 2)  # token-major -> [B,8,Q,d] view
    q, k, v = qkv[0], qkv[1], qkv[2]
    d1 = h(do1, 64)
    delta1 = (d1.float() * h(o1, 64).float()).sum(-1)
    dq, dk, dv = _attn_bwd_core(q, k, v, d1, lse1, delta1, LOG2E / 8.0, 1.0, 1.0 / 8.0)
    d_qkv = torch.stack([dq, dk, dv]).to(BF16)
    r = 1.0 / math.sqrt(128.0)
    qc, kc, vc = cat[0], cat[1], cat[2]
    d2 = h(do2, 128)
    delta2 = (d2.float() * h(o2, 128).float()).sum(-1)
    s = _bmm_f32(qc, kc.transpose(-1, -2))
    p = torch.exp2(s * LOG2E - lse2[..., None])
    dp = _bmm_f32(d2, vc.transpose(-1, -2))
    ds = (p * (dp * r - delta2[..., None])).to(BF16)
    dvc = _bmm_f32((p * r).to(BF16).transpose(-1, -2), d2)
    dqc = _bmm_f32(ds, kc)
    dkc = _bmm_f32(ds.transpose(-1, -2), qc)
    d_cat = torch.stack([dqc, dkc, dvc]).to(BF16)
    return d_qkv, d_cat


def dec_qkv_prep_bwd(d_qkv, d_cat, pairs, B, Q):
    """Backward of destr_dec_qkv_prep (head-major grads in): scatter-add of the left/right gathers +
    position-add fan-in.  -> (d_qkv_obj bf16 [B*Q,1536], d_qk_pos bf16 [B*Q,512])."""
    M = B * Q
    tok = lambda t: t.transpose(1, 2).reshape(M, -1)  # [B,8,Q,d] -> [B*Q, 8*d]
    base = (torch.arange(B, device=pairs.device, dtype=torch.int64) * Q)[:, None]
    gl = (pairs[..., 0].long() + base).reshape(-1)
    gr = (pairs[..., 1].long() + base).reshape(-1)
    parts = []
    for w in range(3):
        acc = tok(d_qkv[w]).float()
        c = tok(d_cat[w]).float().view(M, 8, 128)
        acc = acc.index_add(0, gl, c[:, :, :64].reshape(M, 512))
        acc = acc.index_add(0, gr, c[:, :, 64:].reshape(M, 512))
        parts.append(acc)
    d_obj = torch.cat(parts, dim=-1).to(BF16)
    d_pos = torch.cat([parts[0][:, :256] + parts[0][:, 256:], parts[1][:, :256] + parts[1][:, 256:]], dim=-1).to(BF16)
    return d_obj, d_pos
