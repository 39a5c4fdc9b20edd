"""Tensor-level wrappers over the C-ABI kernels (raw pointers + current CUDA stream).

Each function allocates its outputs with torch (caching allocator), passes `data_ptr()`s and the
current stream to libdestr_b200.so, and returns torch tensors.  No function here has a CPU or
pure-torch fallback.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

Tensor = torch.Tensor
BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: Tensor, dtype, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (destr_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t


def copy_many(dsts, srcs) -> None:
    """dst[k].copy_(src[k]) for up to 16 contiguous CUDA tensor pairs of equal byte size in ONE launch (the batch
    hand-over into a captured step's static inputs, engine.py)."""
    import ctypes as C
    n = len(dsts)
    if n != len(srcs) or not 0 < n <= 16:
        raise ValueError("copy_many: 1..16 (dst, src) pairs")
    sizes = []
    for d, s_ in zip(dsts, srcs):
        if not (d.is_cuda and s_.is_cuda):
            raise RuntimeError("copy_many: expected CUDA tensors (destr_b200 has no CPU path)")
        nb = d.numel() * d.element_size()
        if nb != s_.numel() * s_.element_size() or d.dtype != s_.dtype or not (d.is_contiguous() and s_.is_contiguous()):
            raise ValueError("copy_many: pairs must be contiguous, of the same dtype and size")
        sizes.append(nb)
    src = (C.c_void_p * n)(*[t.data_ptr() for t in srcs])
    dst = (C.c_void_p * n)(*[t.data_ptr() for t in dsts])
    nbytes = (C.c_int64 * n)(*sizes)
    _lib.call("destr_copy_many", C.cast(src, C.c_void_p), C.cast(dst, C.c_void_p), C.cast(nbytes, C.c_void_p), n, _stream())


def _dargs(drop):
    """drop = None | (seed int32/uint32 device tensor [1], thr16, site[, site2]) -> C arguments (seed ptr, thr16, site...)."""
    if drop is None or drop[1] == 0:
        return None, 0, 0
    return drop[0].data_ptr(), int(drop[1]), int(drop[2])


def drop_thr16(p: float) -> int:
    """probability -> 16-bit threshold of the kernels' keep test (kept values are scaled by 65536/(65536-thr16))."""
    return int(round(float(p) * 65536.0))


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------------------------
# masks / positional embeddings
# ----------------------------------------------------------------------------------------------
def mask_words(n_keys: int) -> int:
    # >= 4 words per 128-key tile (decoder / backward kernels) and >= 3 words per 96-key tile (encoder forward)
    return (max(((n_keys + 127) // 128) * 4, ((n_keys + 95) // 96) * 3) + 3) // 4 * 4


def pack_key_mask(kpm: Optional[Tensor], B: int, n_keys: int, device=None) -> Tensor:
    """kpm: bool/uint8 [B, n_keys] (True = padded) or None -> uint32-bit words [B, mask_words]."""
    wpr = mask_words(n_keys)
    if kpm is not None:
        kpm = _chk(kpm.contiguous().view(torch.uint8) if kpm.dtype == torch.bool else kpm.contiguous(),
                   torch.uint8, "kpm")
        device = kpm.device
    bits = torch.empty(B, wpr, dtype=torch.int32, device=device)
    _lib.call("destr_pack_key_mask", _ptr(kpm), bits.data_ptr(), B, n_keys, wpr, _stream())
    return bits


def sine_pos2d(mask: Tensor, want_f32: bool = True, want_bf16: bool = True):
    """mask bool [B,H,W] -> token-major pos [B, H*W, 256] (fp32, bf16)."""
    B, H, W = mask.shape
    m = _chk(mask.contiguous().view(torch.uint8), torch.uint8, "mask")
    pf = torch.empty(B, H * W, 256, dtype=torch.float32, device=m.device) if want_f32 else None
    pb = torch.empty(B, H * W, 256, dtype=BF16, device=m.device) if want_bf16 else None
    _lib.call("destr_sine_pos2d", m.data_ptr(), _ptr(pf), _ptr(pb), B, H, W, _stream())
    return pf, pb


def query_sine_embed(centers: Tensor, want_f32: bool = True, want_bf16: bool = False):
    c = _chk(centers.contiguous(), torch.float32, "centers")
    M = c.numel() // 2
    shape = c.shape[:-1] + (256,)
    of = torch.empty(shape, dtype=torch.float32, device=c.device) if want_f32 else None
    ob = torch.empty(shape, dtype=BF16, device=c.device) if want_bf16 else None
    _lib.call("destr_query_sine_embed", c.data_ptr(), _ptr(of), _ptr(ob), M, _stream())
    return of, ob


# ----------------------------------------------------------------------------------------------
# elementwise / layernorm (raw, non-autograd; autograd wrappers live in functional.py)
# ----------------------------------------------------------------------------------------------
def pos_mul_add(x: Tensor, pos: Tensor, s: Tensor) -> Tensor:
    x, pos, s = (_chk(t.contiguous(), BF16, n) for t, n in ((x, "x"), (pos, "pos"), (s, "s")))
    y = torch.empty_like(x)
    _lib.call("destr_pos_mul_add_fwd", x.data_ptr(), pos.data_ptr(), s.data_ptr(), y.data_ptr(), x.numel(), _stream())
    return y


def pos_mul_add_bwd(dy: Tensor, pos: Tensor) -> Tensor:
    dy, pos = _chk(dy.contiguous(), BF16, "dy"), _chk(pos.contiguous(), BF16, "pos")
    ds = torch.empty_like(dy)
    _lib.call("destr_pos_mul_add_bwd", dy.data_ptr(), pos.data_ptr(), ds.data_ptr(), dy.numel(), _stream())
    return ds


def mul(a: Tensor, b: Tensor) -> Tensor:
    a, b = _chk(a.contiguous(), BF16, "a"), _chk(b.contiguous(), BF16, "b")
    y = torch.empty_like(a)
    _lib.call("destr_mul_fwd", a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), _stream())
    return y


def _rows(t: Tensor, dtype, name: str, D: int) -> Tensor:
    """2-D [M, D] view with unit column stride (row pitch free)."""
    _chk(t, dtype, name)
    if t.dim() != 2:
        t = t.reshape(-1, t.shape[-1])
    if t.stride(1) != 1 or t.shape[1] != D or t.stride(0) % 8 != 0:
        t = t.contiguous()
    return t


def add_layernorm(a: Tensor, b: Optional[Tensor], gamma: Tensor, beta: Tensor, save_stats: bool = False,
                  out: Optional[Tensor] = None, drop=None):
    """y = LN(a + dropout(b)).  a, b (and `out`, if given) may be strided [M, D] views; drop: see _dargs."""
    D = a.shape[-1]
    a = _rows(a, BF16, "a", D)
    M = a.shape[0]
    if b is not None:
        b = _rows(b, BF16, "b", D)
    g, be = _chk(gamma.contiguous(), torch.float32, "gamma"), _chk(beta.contiguous(), torch.float32, "beta")
    y = torch.empty(M, D, dtype=BF16, device=a.device) if out is None else out
    mean = torch.empty(M, dtype=torch.float32, device=a.device) if save_stats else None
    rstd = torch.empty(M, dtype=torch.float32, device=a.device) if save_stats else None
    _lib.call("destr_add_layernorm_fwd", a.data_ptr(), a.stride(0), _ptr(b), 0 if b is None else b.stride(0),
              g.data_ptr(), be.data_ptr(), y.data_ptr(), y.stride(0), _ptr(mean), _ptr(rstd), M, D, *_dargs(drop), _stream())
    return (y, mean, rstd) if save_stats else y


def add_layernorm2(a: Tensor, b: Tensor, gamma1: Tensor, beta1: Tensor, c: Tensor, gamma2: Tensor, beta2: Tensor,
                   drop=None):
    """y1 = LN1(a + dropout(b)), y2 = LN2(c + y1) in one row pass (D = 256).  -> (y1, mean1, rstd1, y2, mean2, rstd2)."""
    a, b, c = _rows(a, BF16, "a", 256), _rows(b, BF16, "b", 256), _rows(c, BF16, "c", 256)
    M = a.shape[0]
    dev = a.device
    y1, y2 = torch.empty(M, 256, dtype=BF16, device=dev), torch.empty(M, 256, dtype=BF16, device=dev)
    st = torch.empty(4, M, dtype=torch.float32, device=dev)
    _lib.call("destr_add_layernorm2_fwd", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
              _chk(gamma1, torch.float32, "gamma1").data_ptr(), _chk(beta1, torch.float32, "beta1").data_ptr(), y1.data_ptr(),
              256, st[0].data_ptr(), st[1].data_ptr(), c.data_ptr(), c.stride(0),
              _chk(gamma2, torch.float32, "gamma2").data_ptr(), _chk(beta2, torch.float32, "beta2").data_ptr(), y2.data_ptr(),
              256, st[2].data_ptr(), st[3].data_ptr(), M, 256, *_dargs(drop), _stream())
    return y1, st[0], st[1], y2, st[2], st[3]


def add_layernorm2_bwd(dy: Tensor, c: Tensor, y1: Tensor, gamma2: Tensor, mean2: Tensor, rstd2: Tensor, a: Tensor,
                       b: Tensor, gamma1: Tensor, mean1: Tensor, rstd1: Tensor, dgamma2: Tensor, dbeta2: Tensor,
                       dgamma1: Tensor, dbeta1: Tensor, dbias: Optional[Tensor] = None, drop=None, want_sum: bool = False):
    """Backward of add_layernorm2 (dense [M,256] bf16 operands).  -> (d3, dxb, dsum or None); the five parameter
    gradients (fp32 [256]) are accumulated into."""
    ts = [_chk(t.contiguous(), BF16, n) for t, n in ((dy, "dy"), (c, "c"), (y1, "y1"), (a, "a"), (b, "b"))]
    dy, c, y1, a, b = ts
    M = a.shape[0]
    dev = a.device
    d3, dxb = torch.empty(M, 256, dtype=BF16, device=dev), torch.empty(M, 256, dtype=BF16, device=dev)
    dsum = torch.empty(M, 256, dtype=BF16, device=dev) if want_sum else None
    _lib.call("destr_add_layernorm2_bwd", dy.data_ptr(), c.data_ptr(), y1.data_ptr(), gamma2.data_ptr(), mean2.data_ptr(),
              rstd2.data_ptr(), a.data_ptr(), b.data_ptr(), gamma1.data_ptr(), mean1.data_ptr(), rstd1.data_ptr(),
              d3.data_ptr(), dxb.data_ptr(), _ptr(dsum), dgamma2.data_ptr(), dbeta2.data_ptr(), dgamma1.data_ptr(),
              dbeta1.data_ptr(), _ptr(dbias), M, 256, *_dargs(drop), _stream())
    return d3, dxb, dsum


def add_layernorm_bwd(dy: Tensor, a: Tensor, b: Optional[Tensor], gamma: Tensor, mean: Tensor, rstd: Tensor,
                      dgamma: Optional[Tensor] = None, dbeta: Optional[Tensor] = None, dbias: Optional[Tensor] = None,
                      res_in: Optional[Tensor] = None, dx_out: Optional[Tensor] = None,
                      res_out: Optional[Tensor] = None, drop=None, want_sum: bool = False):
    """dx = gradient w.r.t. b (through b's dropout mask, if any).  dgamma/dbeta/dbias (fp32 [D]) are accumulated into
    when given (else fresh zeros for dgamma/dbeta).  res_in (or want_sum): returns additionally
    res_out = [res_in +] d(a+b), the un-masked gradient of the residual stream."""
    D = a.shape[-1]
    dy = _rows(dy, BF16, "dy", D)
    a = _rows(a, BF16, "a", D)
    if b is not None:
        b = _rows(b, BF16, "b", D)
    M = a.shape[0]
    dx = torch.empty(M, D, dtype=BF16, device=a.device) if dx_out is None else dx_out
    dg = torch.zeros(D, dtype=torch.float32, device=a.device) if dgamma is None else dgamma
    db = torch.zeros(D, dtype=torch.float32, device=a.device) if dbeta is None else dbeta
    ro = None
    if res_in is not None:
        res_in = _rows(res_in, BF16, "res_in", D)
    if res_in is not None or want_sum:
        ro = torch.empty(M, D, dtype=BF16, device=a.device) if res_out is None else res_out
    _lib.call("destr_add_layernorm_bwd", dy.data_ptr(), dy.stride(0), a.data_ptr(), a.stride(0), _ptr(b),
              0 if b is None else b.stride(0), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
              dx.stride(0), dg.data_ptr(), db.data_ptr(), _ptr(dbias), _ptr(res_in),
              0 if res_in is None else res_in.stride(0), _ptr(ro), 0 if ro is None else ro.stride(0), M, D,
              *_dargs(drop), _stream())
    if ro is not None:
        return dx, dg, db, ro
    return dx, dg, db


def pos_mul_add_bwd_acc(dy: Tensor, pos: Tensor, dx_in: Tensor):
    """-> (ds = dy*pos, dx_out = dx_in + dy)."""
    ds, dx_out = torch.empty_like(dy), torch.empty_like(dy)
    _lib.call("destr_pos_mul_add_bwd_acc", dy.data_ptr(), pos.data_ptr(), dx_in.data_ptr(), ds.data_ptr(),
              dx_out.data_ptr(), dy.numel(), _stream())
    return ds, dx_out


def relu_bwd_colsum(dy: Tensor, h: Optional[Tensor], dbias: Tensor, scale: float = 1.0) -> Optional[Tensor]:
    """dpre = scale*dy*(h>0) (returned), dbias += colsum(dpre).  h=None: dbias += colsum(dy), returns None.
    scale = 1/(1-p) when h is a dropped post-ReLU activation (its zeros carry the dropout mask).
    dy, h: bf16 [M,C] views with unit column stride."""
    M, C = dy.shape
    dpre = torch.empty(M, C, dtype=BF16, device=dy.device) if h is not None else None
    _lib.call("destr_relu_bwd_colsum", dy.data_ptr(), dy.stride(0), _ptr(h), 0 if h is None else h.stride(0),
              _ptr(dpre), C, dbias.data_ptr(), M, C, float(scale), _stream())
    return dpre


def dropout_inplace(x: Tensor, drop) -> Tensor:
    """x <- dropout(x) in place (bf16 [M,C], unit column stride); drop: see _dargs."""
    ptr, thr, site = _dargs(drop)
    if thr:
        _chk(x, BF16, "x")
        _lib.call("destr_dropout_inplace", x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], ptr, thr, site, _stream())
    return x


def linear_bias_relu_dropout(x: Tensor, w: Tensor, bias: Optional[Tensor], drop=None, relu: bool = True) -> Tensor:
    """dropout(relu(x w^T + bias)) as one tcgen05 GEMM (csrc/gemm_bias_relu.cu): x bf16 [M,K] (unit column stride),
    w bf16 [N,K] contiguous, bias fp32 [N]; K == 256, N % 256 == 0.  drop: see _dargs (None = no dropout)."""
    _chk(x, BF16, "x"), _chk(w, BF16, "w")
    M, K = x.shape
    N = w.shape[0]
    if x.stride(1) != 1 or not w.is_contiguous() or w.shape[1] != K or K != 256 or N % 256:
        raise ValueError(f"linear_bias_relu_dropout: unsupported operands x {tuple(x.shape)} w {tuple(w.shape)}")
    out = torch.empty(M, N, dtype=BF16, device=x.device)
    ptr, thr, site = _dargs(drop)
    _lib.call("destr_linear_bias_relu_dropout", x.data_ptr(), x.stride(0), w.data_ptr(),
              _ptr(None if bias is None else _chk(bias, torch.float32, "bias")), out.data_ptr(), N, M, N, K, int(relu),
              ptr, thr, site, _stream())
    return out


# ----------------------------------------------------------------------------------------------
# tcgen05 GEMM family with fused epilogues (csrc/gemm_tc.cu)
# ----------------------------------------------------------------------------------------------
def _mat(t: Optional[Tensor], name: str, rows: int, cols: int) -> Optional[Tensor]:
    """bf16 [rows, cols] view with unit column stride and a row pitch that is a multiple of 8 elements."""
    if t is None:
        return None
    _chk(t, BF16, name)
    if t.dim() != 2 or t.shape[0] != rows or t.shape[1] != cols:
        raise ValueError(f"{name}: expected [{rows}, {cols}], got {tuple(t.shape)}")
    if t.stride(1) != 1 or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0:
        raise ValueError(f"{name}: needs unit column stride, a row pitch multiple of 8 and 16-byte alignment")
    return t


def gemm(a: Tensor, b: Tensor, *, b_kn: bool = False, bias: Optional[Tensor] = None, relu: bool = False, drop=None,
         mul: Optional[Tensor] = None, add: Optional[Tensor] = None, out: Optional[Tensor] = None,
         add2: Optional[Tensor] = None, out2=None):
    """out = add + mul * dropout(act(a . op(b) + bias)) on tcgen05 tensor cores (destr_gemm_bf16).
    a bf16 [M,K]; b bf16 [N,K] (b_kn=False: nn.Linear weight, forward) or [K,N] (b_kn=True: dX = dY W);
    bias fp32 [N]; mul/add/add2 bf16 [M,N] views; out2=True (or a tensor) additionally returns add2 + x.
    Returns out, or (out, out2)."""
    M, K = a.shape
    N = b.shape[1] if b_kn else b.shape[0]
    if a.stride(1) != 1 or a.stride(0) % 8 != 0:  # e.g. a transposed view: TMA needs rows of contiguous elements
        a = a.contiguous()
    a = _mat(a, "a", M, K)
    b = _mat(b, "b", K if b_kn else N, N if b_kn else K)
    if out is None:
        out = torch.empty(M, N, dtype=BF16, device=a.device)
    if out2 is True:
        out2 = torch.empty(M, N, dtype=BF16, device=a.device)
    for t, n in ((mul, "mul"), (add, "add"), (add2, "add2"), (out, "out"), (out2, "out2")):
        _mat(t, n, M, N)
    ld = lambda t: 0 if t is None else t.stride(0)
    _lib.call("destr_gemm_bf16", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), int(b_kn), M, N, K,
              _ptr(None if bias is None else _chk(bias, torch.float32, "bias")), int(relu), *_dargs(drop), _ptr(mul), ld(mul),
              _ptr(add), ld(add), out.data_ptr(), out.stride(0), _ptr(add2), ld(add2), _ptr(out2), ld(out2), _stream())
    return out if out2 is None else (out, out2)


def gemm_relu_bwd(dy: Tensor, w: Tensor, h: Tensor, scale: float, dbias: Optional[Tensor]) -> Tensor:
    """dpre = scale * (dy @ w) * (h > 0), dbias += colsum(dpre): dy bf16 [M,K], w bf16 [K,N] (fc2.weight), h bf16 [M,N]."""
    M, K = dy.shape
    N = w.shape[1]
    dy, w, h = _mat(dy, "dy", M, K), _mat(w, "w", K, N), _mat(h, "h", M, N)
    dpre = torch.empty(M, N, dtype=BF16, device=dy.device)
    _lib.call("destr_gemm_relu_bwd", dy.data_ptr(), dy.stride(0), w.data_ptr(), w.stride(0), M, N, K, h.data_ptr(),
              h.stride(0), float(scale), dpre.data_ptr(), N, _ptr(dbias), _stream())
    return dpre


def gemm_res_ln(a: Tensor, w: Tensor, bias: Optional[Tensor], res: Tensor, gamma: Tensor, beta: Tensor, drop=None,
                want_z: bool = True, res2: Optional[Tensor] = None, gamma2: Optional[Tensor] = None,
                beta2: Optional[Tensor] = None, y2_out: Optional[Tensor] = None):
    """z = res + dropout(a w^T + bias); y = LN(z); with res2: y2 = LN(res2 + y).  w bf16 [256,K].
    Returns (y, z, mean, rstd) or (y, z, mean, rstd, y2, mean2, rstd2); z is None when want_z is False."""
    M, K = a.shape
    a, w, res = _mat(a, "a", M, K), _mat(w, "w", 256, K), _mat(res, "res", M, 256)
    dev = a.device
    y = torch.empty(M, 256, dtype=BF16, device=dev)
    z = torch.empty(M, 256, dtype=BF16, device=dev) if want_z else None
    mean, rstd = torch.empty(M, dtype=torch.float32, device=dev), torch.empty(M, dtype=torch.float32, device=dev)
    y2 = mean2 = rstd2 = None
    if res2 is not None:
        res2 = _mat(res2, "res2", M, 256)
        y2 = torch.empty(M, 256, dtype=BF16, device=dev) if y2_out is None else _mat(y2_out, "y2", M, 256)
        mean2, rstd2 = torch.empty(M, dtype=torch.float32, device=dev), torch.empty(M, dtype=torch.float32, device=dev)
    _lib.call("destr_gemm_res_ln", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, K,
              _ptr(None if bias is None else _chk(bias, torch.float32, "bias")), *_dargs(drop), res.data_ptr(), res.stride(0),
              _chk(gamma, torch.float32, "gamma").data_ptr(), _chk(beta, torch.float32, "beta").data_ptr(), _ptr(z), 256,
              y.data_ptr(), 256, mean.data_ptr(), rstd.data_ptr(), _ptr(res2), 0 if res2 is None else res2.stride(0),
              _ptr(gamma2), _ptr(beta2), _ptr(y2), 0 if y2 is None else y2.stride(0), _ptr(mean2), _ptr(rstd2), _stream())
    if res2 is None:
        return y, z, mean, rstd
    return y, z, mean, rstd, y2, mean2, rstd2


def gemm_dw(dy: Tensor, x: Tensor, dw: Tensor) -> Tensor:
    """dw (fp32 [Nout,Kin] view, ACCUMULATED into) += dy^T x;  dy bf16 [M,Nout], x bf16 [M,Kin]."""
    M, Nout = dy.shape
    Kin = x.shape[1]
    dy, x = _mat(dy, "dy", M, Nout), _mat(x, "x", M, Kin)
    _chk(dw, torch.float32, "dw")
    if tuple(dw.shape) != (Nout, Kin) or dw.stride(1) != 1:
        raise ValueError(f"dw: expected fp32 [{Nout}, {Kin}] with unit column stride, got {tuple(dw.shape)}")
    _lib.call("destr_gemm_dw", dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), M, Nout, Kin, dw.data_ptr(),
              dw.stride(0), _stream())
    return dw


# ----------------------------------------------------------------------------------------------
# encoder attention
# ----------------------------------------------------------------------------------------------
def attn_dropout_bits(drop, BH: int, N: int, device, want_row: bool = True, want_col: bool = True,
                      n_sites: int = 1, site_stride: int = 0, out=None):
    """Bit matrices (1 = dropped) of the encoder attention-probability dropout mask for `drop` = (seed, thr16, site):
    (rowbits, colbits), int32 [n_sites, BH, mask_words(N), 128*ceil(N/128)] each (the leading axis is dropped when
    n_sites == 1) -- the forward kernel reads rowbits, the backward colbits (keep colbits with the saved activations
    to skip regenerating it).  Site i of the batch is drop's site + i*site_stride."""
    seed, thr, site = drop
    words, Np = mask_words(N), (N + 127) // 128 * 128
    shape = (BH, words, Np) if n_sites == 1 else (n_sites, BH, words, Np)
    if out is not None:  # caller-provided (rowbits, colbits) of that shape
        rb, cb = out
    else:
        rb = torch.empty(shape, dtype=torch.int32, device=device) if want_row else None
        cb = torch.empty(shape, dtype=torch.int32, device=device) if want_col else None
    _lib.call("destr_attn_dropout_bits", seed.data_ptr(), int(thr), int(site), int(site_stride), n_sites, BH, N, words,
              _ptr(rb), _ptr(cb), _stream())
    return rb, cb


def _bits_args(drop, bits, BH, N, device, row: bool):
    if drop is None or not drop[1]:
        return None, 0, 0
    if bits is None:
        bits = attn_dropout_bits(drop, BH, N, device, want_row=row, want_col=not row)[0 if row else 1]
    return bits.data_ptr(), bits.shape[1], int(drop[1])


def enc_attn_fwd(q: Tensor, k: Tensor, v: Tensor, mask_bits: Tensor, B: int, N: int, heads: int,
                 scale: float, need_lse: bool = True, drop=None, rowbits: Optional[Tensor] = None):
    """q,k,v: bf16 2-D views [B*N, heads*32] (any row pitch, unit column stride).
    Returns (out bf16 [B*N, heads*32], lse fp32 [B,heads,N]).  drop = (seed, thr16, site) applies the attention
    dropout; `rowbits` from attn_dropout_bits (generated here when omitted)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, BF16, n)
        if t.dim() != 2 or t.stride(1) != 1 or t.shape != (B * N, heads * 32):
            raise ValueError(f"{n}: expected a [B*N, heads*32] view with unit column stride, got {tuple(t.shape)} "
                             f"strides {t.stride()}")
    out = torch.empty(B * N, heads * 32, dtype=BF16, device=q.device)
    lse = torch.empty(B, heads, N, dtype=torch.float32, device=q.device) if need_lse else None
    _lib.call("destr_enc_attn_fwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), k.stride(0), v.stride(0),
              mask_bits.data_ptr(), mask_bits.shape[1], out.data_ptr(), _ptr(lse), B, N, heads, float(scale),
              *_bits_args(drop, rowbits, B * heads, N, q.device, True), _stream())
    return out, lse


def enc_attn_bwd(q: Tensor, k: Tensor, v: Tensor, mask_bits: Tensor, out: Tensor, dout: Tensor, lse: Tensor,
                 B: int, N: int, heads: int, scale: float, drop=None, colbits: Optional[Tensor] = None):
    """Returns (dqk bf16 [B*N, 2*heads*32] = [dq | dk], dv bf16 [B*N, heads*32])."""
    C = heads * 32
    dout = _chk(dout.contiguous(), BF16, "dout")
    dqk = torch.empty(B * N, 2 * C, dtype=BF16, device=q.device)
    dv = torch.empty(B * N, C, dtype=BF16, device=q.device)
    delta = torch.empty(_lib.lib.destr_enc_attn_bwd_stats_floats(B, N, heads), dtype=torch.float32, device=q.device)  # stats workspace
    dq_acc = torch.empty(B * N, C, dtype=torch.float32, device=q.device)
    dq, dk = dqk[:, :C], dqk[:, C:]
    _lib.call("destr_enc_attn_bwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), k.stride(0), v.stride(0),
              mask_bits.data_ptr(), mask_bits.shape[1], out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
              delta.data_ptr(), dq_acc.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), 2 * C, 2 * C, C,
              B, N, heads, float(scale), *_bits_args(drop, colbits, B * heads, N, q.device, False), _stream())
    return dqk, dv


# ----------------------------------------------------------------------------------------------
# decoder small kernels
# ----------------------------------------------------------------------------------------------
def pair_indices(coords: Tensor) -> Tensor:
    """coords fp32 [B,Q,4] cxcyhw -> int32 [B,Q,2]  (_get_pairs)."""
    c = _chk(coords.contiguous(), torch.float32, "coords")
    B, Q, _ = c.shape
    pairs = torch.empty(B, Q, 2, dtype=torch.int32, device=c.device)
    _lib.call("destr_pair_indices", c.data_ptr(), pairs.data_ptr(), B, Q, _stream())
    return pairs


def box_refine(delta: Tensor, centers: Tensor) -> Tensor:
    d = _chk(delta.contiguous(), torch.float32, "delta")
    c = _chk(centers.contiguous(), torch.float32, "centers")
    out = torch.empty_like(d)
    _lib.call("destr_box_refine", d.data_ptr(), c.data_ptr(), out.data_ptr(), d.numel() // 4, _stream())
    return out


def box_head_refine(hidden: Tensor, W2: Tensor, b2: Tensor, centers: Tensor) -> Tensor:
    """boxes = sigmoid(hidden W2^T + b2 + [logit(centers), 0, 0]): hidden bf16 [M,256] view (post-ReLU output of
    bbox_embed[0]), W2 fp32 [4,256], b2 fp32 [4], centers fp32 [M,2] -> fp32 [M,4]  (decoder_block.py:51-54)."""
    _chk(hidden, BF16, "hidden")
    if hidden.dim() != 2 or hidden.shape[1] != 256 or hidden.stride(1) != 1:
        raise ValueError("hidden: expected a bf16 [M,256] view with unit column stride")
    W2 = _chk(W2.contiguous(), torch.float32, "W2")
    b2 = _chk(b2.contiguous(), torch.float32, "b2")
    c = _chk(centers.contiguous(), torch.float32, "centers")
    M = hidden.shape[0]
    out = torch.empty(M, 4, dtype=torch.float32, device=hidden.device)
    _lib.call("destr_box_head_refine", hidden.data_ptr(), hidden.stride(0), W2.data_ptr(), b2.data_ptr(), c.data_ptr(),
              out.data_ptr(), M, _stream())
    return out


def dec_qkv_prep(qkv_obj: Tensor, qk_pos: Tensor, pairs: Tensor, B: int, Q: int):
    """-> head-major (qkv bf16 [3, B, 8, Q, 64], cat bf16 [3, B, 8, Q, 128])."""
    qkv_obj = _chk(qkv_obj.contiguous(), BF16, "qkv_obj")
    qk_pos = _rows(qk_pos, BF16, "qk_pos", 512)
    pairs = _chk(pairs.contiguous(), torch.int32, "pairs")
    qkv = torch.empty(3, B, 8, Q, 64, dtype=BF16, device=qkv_obj.device)
    cat = torch.empty(3, B, 8, Q, 128, dtype=BF16, device=qkv_obj.device)
    _lib.call("destr_dec_qkv_prep", qkv_obj.data_ptr(), qk_pos.data_ptr(), qk_pos.stride(0), pairs.data_ptr(), qkv.data_ptr(),
              cat.data_ptr(), B, Q, _stream())
    return qkv, cat


def dec_self_pair_attn_fwd(qkv: Tensor, cat: Tensor, B: int, Q: int, need_lse: bool = True, drop=None):
    """-> (o1 bf16 [B*Q,512], o2 bf16 [B*Q,1024], lse1, lse2 fp32 [B,8,Q])."""
    o1 = torch.empty(B * Q, 512, dtype=BF16, device=qkv.device)
    o2 = torch.empty(B * Q, 1024, dtype=BF16, device=qkv.device)
    lse1 = torch.empty(B, 8, Q, dtype=torch.float32, device=qkv.device) if need_lse else None
    lse2 = torch.empty(B, 8, Q, dtype=torch.float32, device=qkv.device) if need_lse else None
    _lib.call("destr_dec_self_pair_attn_fwd", _chk(qkv, BF16, "qkv").data_ptr(), _chk(cat, BF16, "cat").data_ptr(),
              o1.data_ptr(), o2.data_ptr(), _ptr(lse1), _ptr(lse2), B, Q, *_dargs(drop), _stream())
    return o1, o2, lse1, lse2


def dual_ln_mix(x: Tensor, o1: Tensor, o2: Tensor, pairs: Tensor, g1, b1, g2, b2, lam: float, Q: int,
                save_stats: bool = True, drop=None):
    x, o1, o2 = (_chk(t.contiguous(), BF16, n) for t, n in ((x, "x"), (o1, "o1"), (o2, "o2")))
    M = x.shape[0]
    out = torch.empty_like(x)
    stats = torch.empty(M, 4, dtype=torch.float32, device=x.device) if save_stats else None
    _lib.call("destr_dual_ln_mix_fwd", x.data_ptr(), o1.data_ptr(), o2.data_ptr(), pairs.data_ptr(), g1.data_ptr(),
              b1.data_ptr(), g2.data_ptr(), b2.data_ptr(), float(lam), out.data_ptr(), _ptr(stats), M, Q,
              *_dargs(drop), 0 if drop is None else int(drop[3]), _stream())
    return out, stats


def dual_ln_mix_bwd(dout: Tensor, x: Tensor, o1: Tensor, o2: Tensor, pairs: Tensor, g1, g2, stats, lam: float, Q: int,
                    pg: Optional[Sequence[Tensor]] = None, head_major: bool = False, drop=None):
    """pg: optional (dg1, db1, dg2, db2) fp32 [512] buffers to ACCUMULATE the parameter gradients into.
    head_major: do1 -> [B,8,Q,64], do2 -> [B,8,Q,128] and additionally returns (delta1, delta2) fp32 [B,8,Q]."""
    dout = _chk(dout.contiguous(), BF16, "dout")
    M = x.shape[0]
    B = M // Q
    dx = torch.empty_like(x)
    if head_major:
        do1 = torch.empty(B, 8, Q, 64, dtype=BF16, device=x.device)
        do2 = torch.empty(B, 8, Q, 128, dtype=BF16, device=x.device)
        delta = torch.empty(2, B, 8, Q, dtype=torch.float32, device=x.device)
    else:
        do1, do2, delta = torch.empty_like(o1), torch.empty_like(o2), None
    if pg is None:
        pg = torch.zeros(4, 512, dtype=torch.float32, device=x.device)
    _lib.call("destr_dual_ln_mix_bwd", dout.data_ptr(), x.data_ptr(), o1.data_ptr(), o2.data_ptr(), pairs.data_ptr(),
              g1.data_ptr(), g2.data_ptr(), stats.data_ptr(), float(lam), dx.data_ptr(), do1.data_ptr(),
              do2.data_ptr(), pg[0].data_ptr(), pg[1].data_ptr(), pg[2].data_ptr(), pg[3].data_ptr(), M, Q,
              int(head_major), None if delta is None else delta[0].data_ptr(),
              None if delta is None else delta[1].data_ptr(), *_dargs(drop), 0 if drop is None else int(drop[3]),
              _stream())
    if head_major:
        return dx, do1, do2, delta[0], delta[1]
    return dx, do1, do2, pg[0], pg[1], pg[2], pg[3]


def split_cross_attn_fwd(q_obj: Tensor, q_pos: Tensor, k_enc: Tensor, k_pos: Tensor, v: Tensor, mask_bits: Tensor,
                         B: int, Q: int, N: int, need_lse: bool = True, drop=None):
    """q_obj bf16 [B*Q,512], q_pos bf16 [B*Q,256], k_enc/k_pos/v bf16 [B*N,256] views (unit column stride).
    -> (out bf16 [B*Q,512] = [cls|reg], lse fp32 [B,2,Q])."""
    q_obj, q_pos = _chk(q_obj.contiguous(), BF16, "q_obj"), _chk(q_pos.contiguous(), BF16, "q_pos")
    for t, n in ((k_enc, "k_enc"), (k_pos, "k_pos"), (v, "v")):
        _chk(t, BF16, n)
        if t.dim() != 2 or t.stride(1) != 1 or t.shape != (B * N, 256):
            raise ValueError(f"{n}: expected a [B*N,256] view with unit column stride")
    out = torch.empty(B * Q, 512, dtype=BF16, device=q_obj.device)
    lse = torch.empty(B, 2, Q, dtype=torch.float32, device=q_obj.device) if need_lse else None
    ws = torch.empty(_lib.lib.destr_split_cross_attn_ws_floats(B, Q, N), dtype=torch.float32, device=q_obj.device)
    _lib.call("destr_split_cross_attn_fwd", q_obj.data_ptr(), q_pos.data_ptr(), k_enc.data_ptr(), k_pos.data_ptr(),
              v.data_ptr(), k_enc.stride(0), k_pos.stride(0), v.stride(0), mask_bits.data_ptr(), mask_bits.shape[1],
              out.data_ptr(), _ptr(lse), ws.data_ptr(), B, Q, N, 1.0 / math.sqrt(512.0), *_dargs(drop), _stream())
    return out, lse


import os as _os
_FUSED_DEC_BWD = _os.environ.get("DESTR_FUSED_DEC_BWD", "1") == "1"  # "0": two-stage path (ds kernel + batched GEMMs)


def dec_self_pair_attn_bwd(qkv: Tensor, cat: Tensor, do1: Tensor, do2: Tensor, lse1: Tensor, lse2: Tensor,
                           delta1: Tensor, delta2: Tensor, B: int, Q: int, drop=None):
    """Backward of dec_self_pair_attn_fwd.  All operands head-major: qkv [3,B,8,Q,64], cat [3,B,8,Q,128],
    do1 [B,8,Q,64], do2 [B,8,Q,128]; lse/delta fp32 [B,8,Q].  Q <= 128: ONE fused tcgen05 kernel (S, dP, softmax
    backward, dQ / dK / dV).  Q > 128: the tcgen05 kernel recomputes S, dP and does the softmax backward, six batched
    library GEMMs finish.  -> head-major (d_qkv [3,B,8,Q,64], d_cat [3,B,8,Q,128])."""
    dev = qkv.device
    Qp = ((Q + 127) // 128) * 128
    BH = B * 8
    if Q <= 128 and _FUSED_DEC_BWD:  # one kernel: P / dS never leave the chip (csrc/dec_attn_bwd.cu, fused kernel)
        d_qkv = torch.empty(3, B, 8, Q, 64, dtype=BF16, device=dev)
        d_cat = torch.empty(3, B, 8, Q, 128, dtype=BF16, device=dev)
        _lib.call("destr_dec_self_pair_attn_bwd", qkv.data_ptr(), cat.data_ptr(), do1.data_ptr(), do2.data_ptr(),
                  lse1.data_ptr(), lse2.data_ptr(), delta1.data_ptr(), delta2.data_ptr(), d_qkv.data_ptr(),
                  d_cat.data_ptr(), B, Q, *_dargs(drop), _stream())
        return d_qkv, d_cat
    PD = torch.empty(4, BH, Q, Qp, dtype=BF16, device=dev)  # P1, dS1, P2, dS2
    _lib.call("destr_dec_self_pair_attn_bwd_ds", qkv.data_ptr(), cat.data_ptr(), do1.data_ptr(), do2.data_ptr(),
              lse1.data_ptr(), lse2.data_ptr(), delta1.data_ptr(), delta2.data_ptr(), PD[0].data_ptr(),
              PD[1].data_ptr(), PD[2].data_ptr(), PD[3].data_ptr(), B, Q, *_dargs(drop), _stream())
    d_qkv = torch.empty(3, BH, Q, 64, dtype=BF16, device=dev)
    d_cat = torch.empty(3, BH, Q, 128, dtype=BF16, device=dev)
    for x, dx, P, dS, dO in ((qkv.view(3, BH, Q, 64), d_qkv, PD[0][:, :, :Q], PD[1][:, :, :Q], do1.view(BH, Q, 64)),
                             (cat.view(3, BH, Q, 128), d_cat, PD[2][:, :, :Q], PD[3][:, :, :Q], do2.view(BH, Q, 128))):
        torch.bmm(dS, x[1], out=dx[0])                   # dQ = dS K
        torch.bmm(dS.transpose(1, 2), x[0], out=dx[1])   # dK = dS^T Q
        torch.bmm(P.transpose(1, 2), dO, out=dx[2])      # dV = P^T dO
    return d_qkv.view(3, B, 8, Q, 64), d_cat.view(3, B, 8, Q, 128)


def dec_qkv_prep_bwd(d_qkv: Tensor, d_cat: Tensor, pairs: Tensor, B: int, Q: int, d_pos_out: Optional[Tensor] = None):
    """-> (d_qkv_obj bf16 [B*Q,1536], d_qk_pos bf16 [B*Q,512] (written into `d_pos_out` if given, any row pitch))."""
    d_obj = torch.empty(B * Q, 1536, dtype=BF16, device=d_qkv.device)
    d_pos = torch.empty(B * Q, 512, dtype=BF16, device=d_qkv.device) if d_pos_out is None else d_pos_out
    _lib.call("destr_dec_qkv_prep_bwd", d_qkv.data_ptr(), d_cat.data_ptr(), pairs.data_ptr(), d_obj.data_ptr(),
              d_pos.data_ptr(), d_pos.stride(0), B, Q, _stream())
    return d_obj, d_pos


def _bmm_aligned_k(a: Tensor, b: Tensor) -> Tensor:
    """bmm(a [B,M,K], b [B,K,N]) whose contraction length K (= the token count, e.g. 1050) is not a multiple of 8:
    cuBLAS would fall back to its slow 4-byte-aligned kernels for the whole product, so the bulk goes through the
    16-byte-aligned kernels and the last K % 8 keys are a rank-(K % 8) update."""
    K = a.shape[2]
    r = K % 8
    if r == 0 or K < 64:
        return torch.bmm(a, b)
    out = torch.bmm(a[:, :, :K - r], b[:, :K - r])
    return out.baddbmm_(a[:, :, K - r:], b[:, K - r:])


def split_cross_attn_bwd(q_obj: Tensor, q_pos: Tensor, k_enc: Tensor, k_pos: Tensor, v: Tensor, mask_bits: Tensor,
                         out: Tensor, dout: Tensor, lse: Tensor, B: int, Q: int, N: int,
                         dke_out: Optional[Tensor] = None, dkp_out: Optional[Tensor] = None,
                         dv_out: Optional[Tensor] = None, drop=None):
    """Backward of split_cross_attn_fwd.  Q <= 128: one fused tcgen05 kernel (S, dP, softmax backward, dV / dK_enc /
    dK_pos; P never leaves the chip) + two batched tcgen05 GEMMs for dq_obj / dq_pos.  Q > 128: the tcgen05 kernel
    recomputes S, dP and does the softmax backward (P, dS, dS_cls+dS_reg in bf16), five batched library GEMMs finish.
    -> (dq_obj [B*Q,512], dq_pos [B*Q,256], dk_enc, dk_pos, dv [B*N,256]) bf16."""
    Np = ((N + 127) // 128) * 128
    dev = q_obj.device
    dout = _chk(dout.contiguous(), BF16, "dout")
    if Q <= 128 and _FUSED_DEC_BWD:
        # one fused kernel (P stays on chip, key-side gradients finished in it) + two batched tcgen05 GEMMs for the
        # query side, which contracts over all keys of an image
        dS_all = torch.empty(B, 2 * Q, Np, dtype=BF16, device=dev)
        dS_sum = torch.empty(B, Q, Np, dtype=BF16, device=dev)
        delta = torch.empty(B, 2, Q, dtype=torch.float32, device=dev)
        dke = torch.empty(B * N, 256, dtype=BF16, device=dev) if dke_out is None else dke_out
        dkp = torch.empty(B * N, 256, dtype=BF16, device=dev) if dkp_out is None else dkp_out
        dv = torch.empty(B * N, 256, dtype=BF16, device=dev) if dv_out is None else dv_out
        _lib.call("destr_split_cross_attn_bwd_fused", q_obj.data_ptr(), q_pos.data_ptr(), k_enc.data_ptr(),
                  k_pos.data_ptr(), v.data_ptr(), k_enc.stride(0), k_pos.stride(0), v.stride(0), mask_bits.data_ptr(),
                  mask_bits.shape[1], out.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                  dS_all.data_ptr(), dS_sum.data_ptr(), dke.data_ptr(), dke.stride(0), dkp.data_ptr(), dkp.stride(0),
                  dv.data_ptr(), dv.stride(0), B, Q, N, 1.0 / math.sqrt(512.0), *_dargs(drop), _stream())
        dqo = torch.empty(B * 2 * Q, 256, dtype=BF16, device=dev)
        dqp = torch.empty(B * Q, 256, dtype=BF16, device=dev)
        # dq_obj = dS k_enc and dq_pos = (dS_cls + dS_reg) k_pos per image: two batched products, one launch
        _lib.call("destr_gemm_bf16_batched2", dS_all.data_ptr(), Np, 2 * Q, k_enc.data_ptr(), k_enc.stride(0), N, 2 * Q,
                  dqo.data_ptr(), 256, dS_sum.data_ptr(), Np, Q, k_pos.data_ptr(), k_pos.stride(0), N, Q, dqp.data_ptr(),
                  256, 1, B, 256, Np, _stream())
        return dqo.view(B * Q, 512), dqp, dke, dkp, dv
    P_all = torch.empty(B, 2 * Q, Np, dtype=BF16, device=dev)
    dS_all = torch.empty(B, 2 * Q, Np, dtype=BF16, device=dev)
    dS_sum = torch.empty(B, Q, Np, dtype=BF16, device=dev)
    delta = torch.empty(B, 2, Q, dtype=torch.float32, device=dev)
    _lib.call("destr_split_cross_attn_bwd_ds", q_obj.data_ptr(), q_pos.data_ptr(), k_enc.data_ptr(), k_pos.data_ptr(),
              v.data_ptr(), k_enc.stride(0), k_pos.stride(0), v.stride(0), mask_bits.data_ptr(), mask_bits.shape[1],
              out.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(), P_all.data_ptr(), dS_all.data_ptr(),
              dS_sum.data_ptr(), B, Q, N, 1.0 / math.sqrt(512.0), *_dargs(drop), _stream())
    Pv, dSv, dSs = P_all[:, :, :N], dS_all[:, :, :N], dS_sum[:, :, :N]
    v3 = lambda t: t.as_strided((B, N, 256), (N * t.stride(0), t.stride(0), 1))  # [B*N,256] view -> [B,N,256]
    do_v, qo_v, qp_v = dout.view(B, 2 * Q, 256), q_obj.view(B, 2 * Q, 256), q_pos.view(B, Q, 256)
    # the key-side gradients may land directly in column slices of a wider buffer ([B*N,256] views with any row
    # pitch): cuBLAS writes them with ldc = pitch, no copy
    def _into(a, b, dst):
        if dst is None:
            return torch.bmm(a, b).view(B * N, 256)
        torch.bmm(a, b, out=v3(dst))
        return dst
    dv = _into(Pv.transpose(1, 2), do_v, dv_out)
    dke = _into(dSv.transpose(1, 2), qo_v, dke_out)
    dkp = _into(dSs.transpose(1, 2), qp_v, dkp_out)
    dqo = _bmm_aligned_k(dSv, v3(k_enc)).view(B * Q, 512)
    dqp = _bmm_aligned_k(dSs, v3(k_pos)).view(B * Q, 256)
    return dqo, dqp, dke, dkp, dv


# ----------------------------------------------------------------------------------------------
# matcher cost
# ----------------------------------------------------------------------------------------------
def match_cost_blockdiag(logits: Tensor, boxes: Tensor, tgt_ids: Tensor, tgt_boxes: Tensor, tgt_offsets: Tensor,
                         total_t: int, w_class: float, w_bbox: float, w_ciou: float, with_l1: bool,
                         out: Optional[Tensor] = None) -> Tensor:
    """Returns the flat fp32 cost buffer [Q * sum(T_i)]; image b's (Q,T_b) block starts at Q*offsets[b]."""
    lg = _chk(logits.contiguous(), torch.float32, "logits")
    bx = _chk(boxes.contiguous(), torch.float32, "boxes")
    B, Q, Cn = lg.shape
    if out is None:
        out = torch.empty(max(Q * total_t, 1), dtype=torch.float32, device=lg.device)
    if total_t > 0:
        _lib.call("destr_match_cost_blockdiag", lg.data_ptr(), bx.data_ptr(), _chk(tgt_ids, torch.int32, "tgt_ids").data_ptr(),
                  _chk(tgt_boxes, torch.float32, "tgt_boxes").data_ptr(), _chk(tgt_offsets, torch.int32, "tgt_offsets").data_ptr(),
                  out.data_ptr(), B, Q, Cn, float(w_class), float(w_bbox), float(w_ciou), int(with_l1), _stream())
    return out


def lsap_blockdiag(cost: Tensor, tgt_offsets: Tensor, B: int, Q: int, max_targets: int, n_slots: Optional[int] = None,
                   out=None):
    """Hungarian matching of every image on the device (bit-identical to scipy.optimize.linear_sum_assignment).
    cost: flat fp32 buffer of match_cost_blockdiag; tgt_offsets int32 [B+1].
    Returns (pred_idx int64 [B,n], tgt_idx int64 [B,n], valid bool [B,n], status int32 [B]); padded slots are
    (Q, 0, False); status != 0 marks an image whose block holds NaN/-inf (scipy would raise)."""
    n = n_slots if n_slots is not None else min(Q, max_targets)
    dev = cost.device
    if out is None:
        out = (torch.empty(B, n, dtype=torch.int64, device=dev), torch.empty(B, n, dtype=torch.int64, device=dev),
               torch.empty(B, n, dtype=torch.bool, device=dev), torch.empty(B, dtype=torch.int32, device=dev))
    pi, ti, valid, status = out
    _lib.call("destr_lsap_blockdiag", _chk(cost, torch.float32, "cost").data_ptr(),
              _chk(tgt_offsets, torch.int32, "tgt_offsets").data_ptr(), B, Q, max_targets, n, pi.data_ptr(), ti.data_ptr(),
              valid.data_ptr(), status.data_ptr(), _stream())
    return pi, ti, valid, status


# ----------------------------------------------------------------------------------------------
# fused set-prediction loss (forward + backward in one launch)
# ----------------------------------------------------------------------------------------------
class _HeadsFn(torch.autograd.Function):
    """Class + box heads (csrc/heads.cu): one forward launch, two backward launches."""

    @staticmethod
    def forward(ctx, dec, centers, Wc, bc, W1, b1, W2, b2, grad_out=None, off_path=None):
        ctx.grad_out = grad_out
        ctx.off_path = off_path
        dec = _chk(dec.contiguous(), BF16, "dec")
        M, C = dec.shape[0], Wc.shape[0]
        if dec.shape[1] != 512 or tuple(W1.shape) != (256, 256) or tuple(W2.shape) != (4, 256) or Wc.shape[1] != 256:
            raise ValueError("heads: dec [M,512], Wc [C,256], W1 [256,256], W2 [4,256] expected")
        ws = [_chk(t.detach().contiguous(), torch.float32, n) for t, n in
              ((Wc, "Wc"), (bc, "bc"), (W1, "W1"), (b1, "b1"), (W2, "W2"), (b2, "b2"))]
        cen = _chk(centers.contiguous(), torch.float32, "centers")
        logits = torch.empty(M, C, dtype=torch.float32, device=dec.device)
        boxes = torch.empty(M, 4, dtype=torch.float32, device=dec.device)
        hidden = torch.empty(M, 256, dtype=torch.float32, device=dec.device)
        _lib.call("destr_heads_fwd", dec.data_ptr(), cen.data_ptr(), ws[0].data_ptr(), ws[1].data_ptr(), C,
                  ws[2].data_ptr(), ws[3].data_ptr(), ws[4].data_ptr(), ws[5].data_ptr(), logits.data_ptr(),
                  boxes.data_ptr(), hidden.data_ptr(), M, _stream())
        ctx.save_for_backward(dec, hidden, boxes, ws[0], ws[2], ws[4])
        return logits, boxes

    @staticmethod
    def backward(ctx, dlogits, dboxes):
        dec, hidden, boxes, Wc, W1, W2 = ctx.saved_tensors
        M, C, dev = dec.shape[0], Wc.shape[0], dec.device
        f32 = dict(dtype=torch.float32, device=dev)
        dlogits = torch.zeros(M, C, **f32) if dlogits is None else _chk(dlogits.contiguous(), torch.float32, "dlogits")
        dboxes = torch.zeros(M, 4, **f32) if dboxes is None else _chk(dboxes.contiguous(), torch.float32, "dboxes")
        d_dec = torch.empty(M, 512, dtype=BF16, device=dev)
        dh, dz = torch.empty(M, 256, **f32), torch.empty(M, 4, **f32)
        if ctx.grad_out is not None:  # caller-owned gradient tensors (overwritten), e.g. views of a flat buffer
            dWc, dbc, dW1, db1, dW2, db2 = (_chk(t, torch.float32, "grad_out") for t in ctx.grad_out)
            if not all(t.is_contiguous() for t in ctx.grad_out):
                raise ValueError("heads: grad_out tensors must be contiguous")
        else:
            dWc, dbc = torch.empty(C, 256, **f32), torch.empty(C, **f32)
            dW1, db1 = torch.empty(256, 256, **f32), torch.empty(256, **f32)
            dW2, db2 = torch.empty(4, 256, **f32), torch.empty(4, **f32)
        def launch(which):
            _lib.call("destr_heads_bwd", dec.data_ptr(), hidden.data_ptr(), boxes.data_ptr(), dlogits.data_ptr(),
                      dboxes.data_ptr(), Wc.data_ptr(), W1.data_ptr(), W2.data_ptr(), C, d_dec.data_ptr(), dh.data_ptr(),
                      dz.data_ptr(), dWc.data_ptr(), dbc.data_ptr(), dW1.data_ptr(), db1.data_ptr(), dW2.data_ptr(),
                      db2.data_ptr(), M, which, _stream())
        if ctx.grad_out is not None and ctx.off_path is not None:
            # the parameter-gradient kernel (57 us running alone) feeds nothing downstream of d_dec: it goes to the
            # caller's side stream (a parallel branch of the step graph), joined where the caller joins that stream
            launch(1)
            ctx.off_path(lambda: launch(2), dec, hidden, dlogits, dh, dz)
        else:
            launch(3)
        if ctx.grad_out is not None:
            return (d_dec,) + (None,) * 9
        return d_dec, None, dWc, dbc, dW1, db1, dW2, db2, None, None


def heads(dec: Tensor, centers: Tensor, Wc: Tensor, bc: Tensor, W1: Tensor, b1: Tensor, W2: Tensor, b2: Tensor,
          grad_out=None, off_path=None):
    """dec bf16 [M,512] (class stream | box stream), centers fp32 [M,2] -> (logits fp32 [M,C], boxes fp32 [M,4]):
    logits = Linear(Wc, bc)(dec[:, :256]); boxes = sigmoid(Linear(W2,b2)(relu(Linear(W1,b1)(dec[:, 256:]))) +
    [inverse_sigmoid(centers), 0, 0])  (model.py:120-131).  Differentiable w.r.t. dec and the six parameters.
    grad_out = six caller-owned tensors (dWc, dbc, dW1, db1, dW2, db2): backward overwrites them instead of returning
    parameter gradients to autograd.  off_path (with grad_out): callable (fn, *tensors_to_keep_alive) that runs fn on a side
    stream the caller joins later (runtime.FlatParams.off_path) -- the parameter-gradient kernel then leaves the
    critical path."""
    return _HeadsFn.apply(dec, centers, Wc, bc, W1, b1, W2, b2, grad_out, off_path)


class _SetLossFn(torch.autograd.Function):
    """total = w_class*class + w_bbox*bbox + w_ciou*ciou with hand-written gradients (csrc/set_loss.cu)."""

    @staticmethod
    def forward(ctx, logits, boxes, tl, tb, pi, ti, valid, weights, ws):
        B, Q, C = logits.shape
        lg = _chk(logits.contiguous(), torch.float32, "logits")
        bx = _chk(boxes.contiguous(), torch.float32, "boxes")
        losses = torch.empty(4, dtype=torch.float32, device=lg.device)
        dlg, dbx = torch.empty_like(lg), torch.empty_like(bx)
        v8 = valid.contiguous().view(torch.uint8) if valid.dtype == torch.bool else valid.contiguous()
        _lib.call("destr_set_loss_fwd_bwd", lg.data_ptr(), bx.data_ptr(), tl.data_ptr(), tb.data_ptr(), pi.data_ptr(),
                  ti.data_ptr(), v8.data_ptr(), B, Q, C, tl.shape[1], pi.shape[1], float(weights[0]), float(weights[1]),
                  float(weights[2]), losses.data_ptr(), dlg.data_ptr(), dbx.data_ptr(), ws.data_ptr(), _stream())
        ctx.save_for_backward(dlg, dbx)
        ctx.mark_non_differentiable(losses)
        return losses[3].clone(), losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        dlg, dbx = ctx.saved_tensors
        return dlg * g_total, dbx * g_total, None, None, None, None, None, None, None


def set_loss_workspace(B: int, device) -> Tensor:
    """Zero-initialised scratch of destr_set_loss_fwd_bwd (allocate once, reuse every step)."""
    return torch.zeros(3 * B + 1, dtype=torch.float32, device=device)


def set_loss(logits: Tensor, boxes: Tensor, tl: Tensor, tb: Tensor, pi: Tensor, ti: Tensor, valid: Tensor,
             weights=(1.0, 1.0, 1.0), workspace: Optional[Tensor] = None):
    """logits (B,Q,C) fp32, boxes (B,Q,4) cxcyhw fp32; tl (B,Tm) int64 / tb (B,Tm,4) padded targets (xyxy);
    pi, ti (B,n) int64 matched query / target index, valid (B,n) bool.
    Returns (total, losses[4] = class, bbox, ciou, total); `total` is differentiable w.r.t. logits and boxes."""
    for t, n_ in ((tl, "tl"), (pi, "pi"), (ti, "ti")):
        _chk(t, torch.int64, n_)
    tb = _chk(tb.contiguous(), torch.float32, "tb")
    ws = workspace if workspace is not None else set_loss_workspace(logits.shape[0], logits.device)
    return _SetLossFn.apply(logits, boxes, tl.contiguous(), tb, pi.contiguous(), ti.contiguous(), valid, weights, ws)
