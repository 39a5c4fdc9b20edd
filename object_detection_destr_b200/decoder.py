"""Drop-in replacements for the reference decoder-side modules:
src/model/attention/self_attention.py, src/model/attention/pair_self_attention.py and
src/model/blocks/decoder_block.py -- same class names, constructor arguments, forward signatures,
return shapes and parameter names (state_dict-compatible), running on the B200 kernels.
"""
from __future__ import annotations

import copy
import math
from typing import Optional

import torch
from torch import nn

from . import functional as Fn
from . import ops
from .encoder import _params, drop_scope

BF16 = torch.bfloat16


class SelfAttention(nn.Module):
    """reference: self_attention.py:8-47.  Supports the two shapes the reference uses:
    decoder self-attention (B,8,S,64) without masks, and the single-head cross-attention
    (B,1,Sq,512) x (B,1,Sk,512) -> (B,1,Sk,256) with a key-padding mask."""

    def __init__(self, heads_num: int = 8, dropout_prob: float = 0.3, hidden_dim: int = 256):
        super().__init__()
        self._num_heads = heads_num
        self._dropout_prob = dropout_prob
        self._hidden_dim = hidden_dim

    def forward(self, query, key, value, attn_mask: Optional[torch.Tensor] = None,
                key_padding_mask: Optional[torch.Tensor] = None):
        if attn_mask is not None:
            raise NotImplementedError("attn_mask is never passed by the reference decoder (decoder_block.py:179,246)")
        B, H, Sq, dq = query.shape
        Sk, dv = key.shape[2], value.shape[3]
        tok = lambda t: t.transpose(1, 2).reshape(t.shape[0] * t.shape[2], -1).to(BF16)
        if H == 8 and dq == 64 and dv == 64 and Sq == Sk and key_padding_mask is None:
            qkv = torch.cat([tok(query), tok(key), tok(value)], dim=-1)
            ident = torch.arange(Sq, device=query.device, dtype=torch.int32)[None, :, None].expand(B, Sq, 2).contiguous()
            qkv2, cat = Fn._DecQkvPrep.apply(qkv, torch.zeros(B * Sq, 512, dtype=BF16, device=query.device), ident, B, Sq)
            with drop_scope(self, query.device):  # the inline nn.Dropout of the reference: always on (self_attention.py:40)
                o1, _ = Fn._DecSelfPairAttn.apply(qkv2, cat, B, Sq, Fn._dr("d.sa", Fn._dec_site(0, "sa")))
            return o1.view(B, Sq, 512).to(query.dtype)
        if H == 1 and dq == 512 and dv == 256:
            q = tok(query)
            k, v = tok(key), tok(value)
            bits = ops.pack_key_mask(key_padding_mask, B, Sk, device=query.device)
            kpm = None if key_padding_mask is None else key_padding_mask.contiguous()
            q_obj = torch.cat([q[:, :256], q[:, :256]], dim=-1)
            with drop_scope(self, query.device):
                out = Fn._SplitCrossAttn.apply(q_obj, q[:, 256:].contiguous(), k[:, :256], k[:, 256:], v, bits, kpm, B, Sq,
                                               Sk, Fn._dr("d.ca", Fn._dec_site(0, "ca")))
            return out[:, :256].reshape(B, Sq, 256).to(query.dtype)
        raise NotImplementedError(f"SelfAttention shape (H={H}, d_qk={dq}, d_v={dv}) is not on the DESTR hot path")


class PairSelfAttention(nn.Module):
    """reference: pair_self_attention.py:9-107 (8 heads x 64)."""

    def __init__(self, heads_num: int) -> None:
        super().__init__()
        self._heads_num = heads_num

    @property
    def heads_num(self):
        return self._heads_num

    def forward(self, query, key, value, top_k_centers):
        B, H, S, d = query.shape
        if H != 8 or d != 64:
            raise NotImplementedError("pair attention kernel is built for 8 heads x 64")
        tok = lambda t: t.transpose(1, 2).reshape(B * S, H * d).to(BF16)
        pairs = ops.pair_indices(top_k_centers.detach().float().contiguous())
        qkv = torch.cat([tok(query), tok(key), tok(value)], dim=-1)
        qkv2, cat = Fn._DecQkvPrep.apply(qkv, torch.zeros(B * S, 512, dtype=BF16, device=query.device), pairs, B, S)
        _, o2 = Fn._DecSelfPairAttn.apply(qkv2, cat, B, S)
        me = torch.arange(S, device=query.device, dtype=torch.int32)[None, :, None]
        keep = (pairs == me).to(o2.dtype).view(B * S, 2, 1)
        return (o2.view(B * S, 2, 512) * keep).sum(1).view(B, S, 512).to(query.dtype)


class ClsRegBranch(nn.Module):
    """reference: decoder_block.py:223-260 (parameter container; the fused path lives in
    functional.decoder_layer, this forward serves direct use of the module)."""

    def __init__(self, hidden_dim: int = 256):
        super().__init__()
        self.cross_attn = SelfAttention(heads_num=1)
        self.fc1 = nn.Linear(hidden_dim, hidden_dim * 4)
        self.fc2 = nn.Linear(hidden_dim * 4, hidden_dim)
        self.dropout = nn.Dropout(0.3)
        self.norm1 = nn.LayerNorm(hidden_dim)
        self.norm2 = nn.LayerNorm(hidden_dim)

    def forward(self, inputs, query, key, value, key_mask):
        B, Q, d = inputs.shape
        ca = self.cross_attn(query=query.unsqueeze(1), key=key.unsqueeze(1), value=value.unsqueeze(1),
                             key_padding_mask=key_mask)
        p = _params(self)
        S = Fn._dec_site
        with drop_scope(self, inputs.device):
            xb = Fn.add_layernorm(inputs.reshape(B * Q, d).to(BF16), ca.reshape(B * Q, d).to(BF16), p["norm1.weight"],
                                  p["norm1.bias"], Fn._dr("d.br", S(0, "b0.d_ca")))
            h = Fn.dropout(torch.relu(Fn.linear(xb, p["fc1.weight"], p["fc1.bias"])), Fn._dr("d.br", S(0, "b0.d_relu")))
            f = Fn.linear(h, p["fc2.weight"], p["fc2.bias"])
            y = Fn.add_layernorm(xb, f, p["norm2.weight"], p["norm2.bias"], Fn._dr("d.br", S(0, "b0.d_fc2")))
        return y.view(B, Q, d).to(inputs.dtype)


class DecoderBlock(nn.Module):
    """reference: decoder_block.py:70-220."""

    def __init__(self, lambda_: float = 0.5, hidden_dim: int = 256, heads_num: int = 8) -> None:
        super().__init__()
        if hidden_dim != 256 or heads_num != 8:
            raise ValueError("the B200 decoder kernels are built for hidden_dim=256, 8 heads")
        self._heads_num = heads_num
        self._channels = hidden_dim
        self._lambda = lambda_
        self._self_attn = SelfAttention(heads_num=heads_num, hidden_dim=hidden_dim, dropout_prob=0.3)
        self._pair_attn = PairSelfAttention(heads_num=heads_num)
        self._cls_branch = ClsRegBranch(hidden_dim=hidden_dim)
        self._reg_branch = ClsRegBranch(hidden_dim=hidden_dim)
        c = hidden_dim
        self._sa_proj_to_q_obj = nn.Linear(2 * c, 2 * c, bias=False)
        self._sa_proj_to_q_pos = nn.Linear(c, c, bias=False)
        self._sa_proj_to_k_obj = nn.Linear(2 * c, 2 * c, bias=False)
        self._sa_proj_to_k_pos = nn.Linear(c, c, bias=False)
        self._sa_proj_to_v_obj = nn.Linear(2 * c, 2 * c, bias=False)
        self._ca_proj_to_q_obj = nn.Linear(2 * c, 2 * c, bias=False)
        self._ca_proj_to_q_pos = nn.Linear(c, c, bias=False)
        self._ca_proj_to_k_enc = nn.Linear(c, c, bias=False)
        self._ca_proj_to_k_pos = nn.Linear(c, c, bias=False)
        self._ca_proj_to_v_enc = nn.Linear(c, c, bias=False)
        self.norm1 = nn.LayerNorm(2 * c)
        self.norm2 = nn.LayerNorm(2 * c)
        self.dropout1 = nn.Dropout(0.3)

    def forward(self, obj_selected, enc_output, obj_coords, obj_pos_embed, obj_sin_embed, enc_pos_embed,
                enc_key_mask):
        """Reference signature (decoder_block.py:157-166): obj_selected (B,Q,512); enc_output, enc_pos_embed
        (B,N,256); obj_coords (B,Q,4) cxcyhw; obj_pos_embed, obj_sin_embed (B,Q,256); enc_key_mask (B,N) bool."""
        B, Q, _ = obj_selected.shape
        N = enc_output.shape[1]
        t = lambda a: a.reshape(-1, a.shape[-1]).to(BF16)
        p = {"blk." + k: v for k, v in self.named_parameters()}
        kpm = enc_key_mask.contiguous()
        bits = ops.pack_key_mask(kpm, B, N, device=obj_selected.device)
        pairs = ops.pair_indices(obj_coords.detach().float().contiguous())
        qk_pos = Fn.linear(t(obj_pos_embed), torch.cat([p["blk._sa_proj_to_q_pos.weight"], p["blk._sa_proj_to_k_pos.weight"]]))
        kv = Fn.linear(t(enc_output), torch.cat([p["blk._ca_proj_to_k_enc.weight"], p["blk._ca_proj_to_v_enc.weight"]]))
        k_pos = Fn.linear(t(enc_pos_embed), p["blk._ca_proj_to_k_pos.weight"])
        with drop_scope(self, obj_selected.device):
            y = Fn.decoder_block_core(t(obj_selected), t(obj_sin_embed), pairs, qk_pos, kv[:, :256], k_pos, kv[:, 256:],
                                      bits, kpm, p, "blk.", B, Q, N, self._lambda)
        return y.view(B, Q, 512).to(obj_selected.dtype)


class Decoder(nn.Module):
    """reference: decoder_block.py:12-67.  `_pos_scale` and `norm` are shared by all layers."""

    def __init__(self, decoder_block: nn.Module, num_decoder_blocks: int):
        super().__init__()
        self._decoder = nn.ModuleList([copy.deepcopy(decoder_block) for _ in range(num_decoder_blocks)])
        self._num_dec = num_decoder_blocks
        self._pos_scale = nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 256))
        self.norm = nn.LayerNorm(512)

    def forward_tokens(self, x, enc_out, bits, kpm, fine_pos, pos_embed, centers, bbox_embed: nn.Module, B, Q, N):
        """Token-major fast path (bf16 [rows, C] activations, fp32 centers [B*Q,2])."""
        with drop_scope(self, x.device):
            return Fn.decoder_tokens(x, enc_out, bits, kpm, fine_pos, pos_embed, centers, _params(self),
                                     _params(bbox_embed), self._num_dec, B, Q, N)

    def forward(self, selected_objects, encoder_output, mask, fine_pos, selected_objects_pos_embed,
                selected_centers, bbox_embed: nn.Module):
        """Reference signature: selected_objects (B,Q,512); encoder_output, fine_pos (B,N,256);
        mask (B,N) bool; selected_objects_pos_embed (B,Q,256); selected_centers (B,Q,2) -> (B,Q,512)."""
        B, Q, _ = selected_objects.shape
        N = encoder_output.shape[1]
        t = lambda a: a.reshape(-1, a.shape[-1]).to(BF16)
        kpm = mask.contiguous()
        bits = ops.pack_key_mask(kpm, B, N, device=selected_objects.device)
        y = self.forward_tokens(t(selected_objects), t(encoder_output), bits, kpm, t(fine_pos),
                                t(selected_objects_pos_embed), selected_centers.reshape(B * Q, 2).float().contiguous(),
                                bbox_embed, B, Q, N)
        return y.view(B, Q, 512).to(selected_objects.dtype)


def build_decoder(args):
    """reference: decoder_block.py:263-274."""
    return Decoder(decoder_block=DecoderBlock(hidden_dim=args.hidden_dim), num_decoder_blocks=args.num_decoder_blocks)
