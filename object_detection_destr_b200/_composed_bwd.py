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
    """-> (d_qkv bf16 [B*Q,1536], d_cat bf16 [3,B*Q,1024])."""
    h = lambda t, d: t.reshape(B, Q, 8, d).transpose(1, 2)  # [B,8,Q,d]
    back = lambda t: t.transpose(1, 2).reshape(B * Q, -1)
    # self-attention: P = softmax(q.k^T / 8)
    q, k, v = h(qkv[:, :512], 64), h(qkv[:, 512:1024], 64), h(qkv[:, 1024:], 64)
    d1 = h(do1, 64)
    delta1 = (d1.float() * h(o1, 64).float()).sum(-1)
    dq, dk, dv = _attn_bwd_core(q, k, v, d1, lse1, delta1, LOG2E / 8.0, 1.0, 1.0 / 8.0)
    d_qkv = torch.cat([back(dq), back(dk), back(dv)], dim=-1).to(BF16)
    # pair attention: P2 = softmax(qcat.kcat^T) / sqrt(128); O2 = P2.vcat
    r = 1.0 / math.sqrt(128.0)
    qc, kc, vc = h(cat[0], 128), h(cat[1], 128), h(cat[2], 128)
    d2 = h(do2, 128)
    delta2 = (d2.float() * h(o2, 128).float()).sum(-1)  # = sum_j P2_ij dP2_ij
    # dA = softmax o (r*dP - delta2)  ->  use p_scale=1 and fold r:  ds = p*(r*dp - delta2)
    s = _bmm_f32(qc, kc.transpose(-1, -2))
    p = torch.exp2(s * LOG2E - lse2[..., None])
    dp = _bmm_f32(d2, vc.transpose(-1, -2))
    ds = (p * (dp * r - delta2[..., None])).to(BF16)
    dvc = _bmm_f32((p * r).to(BF16).transpose(-1, -2), d2)
    dqc = _bmm_f32(ds, kc)
    dkc = _bmm_f32(ds.transpose(-1, -2), qc)
    d_cat = torch.stack([back(dqc), back(dkc), back(dvc)]).to(BF16)
    return d_qkv, d_cat


def dec_qkv_prep_bwd(d_qkv, d_cat, pairs, B, Q):
    """Backward of destr_dec_qkv_prep: scatter-add of the left/right gathers + position-add fan-in.
    -> (d_qkv_obj bf16 [B*Q,1536], d_qk_pos bf16 [B*Q,512])."""
    M = B * Q
    base = (torch.arange(B, device=pairs.device, dtype=torch.int64) * Q)[:, None]
    gl = (pairs[..., 0].long() + base).reshape(-1)
    gr = (pairs[..., 1].long() + base).reshape(-1)
    parts = []
    for w in range(3):
        acc = d_qkv[:, w * 512:(w + 1) * 512].float()
        c = d_cat[w].float().view(M, 8, 128)
        acc = acc.index_add(0, gl, c[:, :, :64].reshape(M, 512))
        acc = acc.index_add(0, gr, c[:, :, 64:].reshape(M, 512))
        parts.append(acc)
    d_obj = torch.cat(parts, dim=-1).to(BF16)
    d_pos = torch.cat([parts[0][:, :256] + parts[0][:, 256:], parts[1][:, :256] + parts[1][:, 256:]], dim=-1).to(BF16)
    return d_obj, d_pos


def split_cross_attn_bwd(q_obj, q_pos, k_enc, k_pos, v, kpm, out, dout, lse, B, Q, N):
    """-> (dq_obj [B*Q,512], dq_pos [B*Q,256], dk_enc, dk_pos, dv [B*N,256]) bf16."""
    scale = 1.0 / math.sqrt(512.0)
    qo = q_obj.view(B, Q, 2, 256).transpose(1, 2)          # [B,2,Q,256]
    qp = q_pos.view(B, 1, Q, 256).expand(B, 2, Q, 256)
    ke = k_enc.reshape(B, 1, N, 256).expand(B, 2, N, 256)
    kp = k_pos.reshape(B, 1, N, 256).expand(B, 2, N, 256)
    vv = v.reshape(B, 1, N, 256).expand(B, 2, N, 256)
    do = dout.view(B, Q, 2, 256).transpose(1, 2)
    o = out.view(B, Q, 2, 256).transpose(1, 2)
    s = _bmm_f32(qo, ke.transpose(-1, -2)) + _bmm_f32(qp, kp.transpose(-1, -2))
    p = torch.exp2(s * (scale * LOG2E) - lse[..., None])
    if kpm is not None:
        p = p.masked_fill(kpm.bool()[:, None, None, :], 0.0)
    delta = (do.float() * o.float()).sum(-1, keepdim=True)
    dp = _bmm_f32(do, vv.transpose(-1, -2))
    ds = (p * (dp - delta) * scale).to(BF16)
    pb = p.to(BF16)
    dv = _bmm_f32(pb.transpose(-1, -2), do).sum(1)
    dqo = _bmm_f32(ds, ke).transpose(1, 2).reshape(B * Q, 512)
    dqp = _bmm_f32(ds, kp).sum(1).reshape(B * Q, 256)
    dst = ds.transpose(-1, -2)
    dke = _bmm_f32(dst, qo).sum(1).reshape(B * N, 256)
    dkp = _bmm_f32(dst, qp).sum(1).reshape(B * N, 256)
    return dqo.to(BF16), dqp.to(BF16), dke.to(BF16), dkp.to(BF16), dv.reshape(B * N, 256).to(BF16)
