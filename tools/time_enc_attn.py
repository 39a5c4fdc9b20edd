"""Times the encoder attention kernels alone (CUDA events, L2 flushed, mean of 20 launches) at config 2 (or B N from
argv) with and without dropout: forward, backward main kernel, backward op (prep + main + convert), bit generator."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from object_detection_destr_b200 import ops, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1050
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
qk = torch.randn(B * N, 512, generator=g).bfloat16().to(dev)
v = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
do = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
bits = ops.pack_key_mask(None, B, N, device=dev)
scale = 1.0 / math.sqrt(32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, iters=20):
    for _ in range(10):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record(); fn(); en.record(); en.synchronize()
        tot += st.elapsed_time(en)
    return tot / iters * 1e3


fl = 4.0 * N * N * 256 * B
for p in (0.3, 0.0):
    drop = (torch.ones(1, dtype=torch.int32, device=dev), ops.drop_thr16(p), 0) if p > 0 else None
    rb, cb = ops.attn_dropout_bits(drop, B * 8, N, dev) if drop else (None, None)
    out, lse = ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, scale, drop=drop, rowbits=rb)
    tf = t(lambda: ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, scale, drop=drop, rowbits=rb))
    bw = lambda: ops.enc_attn_bwd(qk[:, :256], qk[:, 256:], v, bits, out, do, lse, B, N, 8, scale, drop=drop, colbits=cb)
    to = t(bw)
    _lib.lib.destr_debug_knob(14, 1)
    tb = t(bw)
    _lib.lib.destr_debug_knob(14, 0)
    tg = t(lambda: ops.attn_dropout_bits(drop, B * 8, N, dev)) if drop else 0.0
    print(f"B={B} N={N} p={p}: fwd {tf:6.1f} us ({fl / tf / 1e6 / 1660.7:.3f})  bwd main {tb:6.1f} us "
          f"({2 * fl / tb / 1e6 / 1660.7:.3f})  bwd op {to:6.1f} us  bits {tg:5.1f} us", flush=True)
