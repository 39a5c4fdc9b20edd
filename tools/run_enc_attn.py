"""Runs the encoder attention forward + backward (with the attention dropout of the default bench) twice at the
config-2 shape (B=8, N=1050, 8 heads x 32) -- or B N from argv -- for ncu captures."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from object_detection_destr_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1050
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
qk = torch.randn(B * N, 512, generator=g).bfloat16().to(dev)
v = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
do = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
bits = ops.pack_key_mask(None, B, N, device=dev)
scale = 1.0 / math.sqrt(32)
drop = (torch.ones(1, dtype=torch.int32, device=dev), ops.drop_thr16(p), 0) if p > 0 else None
for _ in range(2):
    rb, cb = ops.attn_dropout_bits(drop, B * 8, N, dev) if drop else (None, None)
    out, lse = ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, scale, drop=drop, rowbits=rb)
    ops.enc_attn_bwd(qk[:, :256], qk[:, 256:], v, bits, out, do, lse, B, N, 8, scale, drop=drop, colbits=cb)
torch.cuda.synchronize()
