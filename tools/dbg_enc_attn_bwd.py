"""Knock-out timings of the encoder attention backward main kernel (destr_debug_knob(18, bits), see enc_attn_bwd.cu:
1 no dV, 2 no dK, 4 no dQ, 8 no softmax-backward math, 16 no S/dP, 32 no dQ reduction, 64 no dK/dV stores, 128 no Q/dO
loads; results are meaningless with a bit set, only the time is read).  profiles/r02_enc_attn_bwd_knockout.txt."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from object_detection_destr_b200 import ops, _lib
B, N = 8, 1050
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
qk = torch.randn(B * N, 512, generator=g).bfloat16().to(dev)
v = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
do = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
bits = ops.pack_key_mask(None, B, N, device=dev)
scale = 1.0 / math.sqrt(32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, iters=20):
    for _ in range(10): fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record(); fn(); en.record(); en.synchronize()
        tot += st.elapsed_time(en)
    return tot / iters * 1e3
for p in (0.3, 0.0):
    drop = (torch.ones(1, dtype=torch.int32, device=dev), ops.drop_thr16(p), 0) if p > 0 else None
    rb, cb = ops.attn_dropout_bits(drop, B * 8, N, dev) if drop else (None, None)
    out, lse = ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, scale, drop=drop, rowbits=rb)
    bw = lambda: ops.enc_attn_bwd(qk[:, :256], qk[:, 256:], v, bits, out, do, lse, B, N, 8, scale, drop=drop, colbits=cb)
    _lib.lib.destr_debug_knob(14, 1)
    for dbg in (0, 8, 23, 127, 0):
        _lib.lib.destr_debug_knob(18, dbg)
        print(f"p={p} dbg={dbg:2d}: bwd main {t(bw):6.1f} us", flush=True)
    _lib.lib.destr_debug_knob(18, 0)
    _lib.lib.destr_debug_knob(14, 0)
