"""ctypes binding of libdestr_b200.so (the C-ABI declared in include/destr_b200.h).

There is no CPU fallback: importing this module without the built library, or calling an op on a
non-CUDA tensor, raises.  Build with `python -c "import __graft_entry__ as g; g.build()"`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdestr_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the CUDA library has not been built. "
        "Run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_p, _i, _f, _i64, _u = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_uint32

# name -> argtypes; mirrors include/destr_b200.h one to one (tests/test_abi.py checks the header)
SIGNATURES = {
    "destr_version": [],
    "destr_debug_knob": [_i, _i],
    "destr_copy_many": [_p, _p, _p, _i, _p],
    "destr_pack_key_mask": [_p, _p, _i, _i, _i, _p],
    "destr_sine_pos2d": [_p, _p, _p, _i, _i, _i, _p],
    "destr_query_sine_embed": [_p, _p, _p, _i, _p],
    "destr_pos_mul_add_fwd": [_p, _p, _p, _p, _i64, _p],
    "destr_pos_mul_add_bwd": [_p, _p, _p, _i64, _p],
    "destr_mul_fwd": [_p, _p, _p, _i64, _p],
    "destr_add_layernorm_fwd": [_p, _i, _p, _i, _p, _p, _p, _i, _p, _p, _i, _i, _p, _u, _u, _p],
    "destr_add_layernorm2_fwd": [_p, _i, _p, _i, _p, _p, _p, _i, _p, _p, _p, _i, _p, _p, _p, _i, _p, _p, _i, _i, _p, _u, _u, _p],
    "destr_add_layernorm2_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _u, _u,
                                 _p],
    "destr_add_layernorm_bwd": [_p, _i, _p, _i, _p, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _p, _i, _i, _i, _p, _u, _u,
                                _p],
    "destr_dropout_inplace": [_p, _i, _i, _i, _p, _u, _u, _p],
    "destr_pos_mul_add_bwd_acc": [_p, _p, _p, _p, _p, _i64, _p],
    "destr_relu_bwd_colsum": [_p, _i, _p, _i, _p, _i, _p, _i, _i, _f, _p],
    "destr_enc_attn_fwd": [_p, _p, _p, _i, _i, _i, _p, _i, _p, _p, _i, _i, _i, _f, _p, _i, _u, _p],
    "destr_attn_dropout_bits": [_p, _u, _u, _u, _i, _i, _i, _i, _p, _p, _p],
    "destr_enc_attn_bwd": [_p, _p, _p, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i,
                           _f, _p, _i, _u, _p],
    "destr_dual_ln_mix_fwd": [_p, _p, _p, _p, _p, _p, _p, _p, _f, _p, _p, _i, _i, _p, _u, _u, _u, _p],
    "destr_dual_ln_mix_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _u, _u,
                              _u, _p],
    "destr_dec_qkv_prep": [_p, _p, _i, _p, _p, _p, _i, _i, _p],
    "destr_dec_self_pair_attn_fwd": [_p, _p, _p, _p, _p, _p, _i, _i, _p, _u, _u, _p],
    "destr_split_cross_attn_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _p, _i, _p, _p, _p, _i, _i, _i, _f, _p, _u, _u, _p],
    "destr_split_cross_attn_bwd_ds": [_p, _p, _p, _p, _p, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i,
                                      _f, _p, _u, _u, _p],
    "destr_split_cross_attn_bwd_fused": [_p, _p, _p, _p, _p, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _i,
                                         _p, _i, _i, _i, _i, _f, _p, _u, _u, _p],
    "destr_gemm_bf16_batched2": [_p, _i, _i, _p, _i, _i, _i, _p, _i, _p, _i, _i, _p, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p],
    "destr_gemm_bf16_batched": [_p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p],
    "destr_dec_self_pair_attn_bwd_ds": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _u, _u, _p],
    "destr_dec_self_pair_attn_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _u, _u, _p],
    "destr_dec_qkv_prep_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "destr_pair_indices": [_p, _p, _i, _i, _p],
    "destr_box_refine": [_p, _p, _p, _i, _p],
    "destr_box_head_refine": [_p, _i, _p, _p, _p, _p, _i, _p],
    "destr_match_cost_blockdiag": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _f, _i, _p],
    "destr_lsap_blockdiag": [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p],
    "destr_linear_bias_relu_dropout": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _p, _u, _u, _p],
    "destr_gemm_bf16": [_p, _i, _p, _i, _i, _i, _i, _i, _p, _i, _p, _u, _u, _p, _i, _p, _i, _p, _i, _p, _i, _p, _i, _p],
    "destr_gemm_relu_bwd": [_p, _i, _p, _i, _i, _i, _i, _p, _i, _f, _p, _i, _p, _p],
    "destr_gemm_res_ln": [_p, _i, _p, _i, _i, _i, _p, _p, _u, _u, _p, _i, _p, _p, _p, _i, _p, _i, _p, _p, _p, _i, _p, _p,
                          _p, _i, _p, _p, _p],
    "destr_gemm_dw": [_p, _i, _p, _i, _i, _i, _i, _p, _i, _p],
    "destr_select_queries": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "destr_heads_fwd": [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p],
    "destr_heads_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "destr_flat_adamw": [_p, _p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _p, _p, _i64, _i64, _f, _p],
    "destr_set_loss_fwd_bwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _f, _p, _p, _p, _p, _p],
}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch: fail loudly
    _fn.argtypes = _args
    _fn.restype = _i
lib.destr_split_cross_attn_ws_floats.argtypes = [_i, _i, _i]
lib.destr_split_cross_attn_ws_floats.restype = _i64
lib.destr_enc_attn_bwd_stats_floats.argtypes = [_i, _i, _i]
lib.destr_enc_attn_bwd_stats_floats.restype = _i
lib.destr_last_error.argtypes = []
lib.destr_last_error.restype = C.c_char_p

# kernels launched per C-ABI call (bench.py reports the sum over a step as gpu_launches)
KERNELS_PER_CALL = {"destr_enc_attn_bwd": 3, "destr_select_queries": 2, "destr_split_cross_attn_fwd": 2, "destr_split_cross_attn_bwd_ds": 2,
                    "destr_split_cross_attn_bwd_fused": 2}
launch_count = 0
# bench.py: {name: []} -> (start, end) CUDA-event pairs are appended around every call of `name`
KERNEL_TIMERS = None


def call(name: str, *args) -> None:
    global launch_count
    timers = KERNEL_TIMERS
    if timers is not None and name in timers:
        import torch
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        rc = getattr(lib, name)(*args)
        en.record()
        timers[name].append((st, en))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (rc={rc}): {lib.destr_last_error().decode()}")
    launch_count += KERNELS_PER_CALL.get(name, 1)
