"""Tiny driver for `ncu --set full`: encoder attention fwd+bwd at the config-2 shape (B=8, N=1050, 8 heads)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from object_detection_destr_b200 import ops, _lib  # noqa: E402

for kv in filter(None, os.environ.get("KNOBS", "").split(",")):  # e.g. KNOBS=18=127,14=1 (debug knobs)
    _lib.lib.destr_debug_knob(int(kv.split("=")[0]), int(kv.split("=")[1]))

B, N = int(os.environ.get("PB", 8)), int(os.environ.get("PN", 1050))
g = torch.Generator().manual_seed(0)
qk = torch.randn(B * N, 512, generator=g).bfloat16().cuda()
v = torch.randn(B * N, 256, generator=g).bfloat16().cuda()
do = torch.randn(B * N, 256, generator=g).bfloat16().cuda()
bits = ops.pack_key_mask(None, B, N, device=qk.device)
sc = 1 / math.sqrt(32)
pdrop = float(os.environ.get("PDROP", 0.0))  # PDROP=0.3: the dropout variants + the mask bit-matrix generator
drop = (torch.ones(1, dtype=torch.int32, device="cuda"), ops.drop_thr16(pdrop), 0) if pdrop > 0 else None
for _ in range(3):
    rb, cb = ops.attn_dropout_bits(drop, B * 8, N, qk.device) if drop else (None, None)
    out, lse = ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, sc, drop=drop, rowbits=rb)
    ops.enc_attn_bwd(qk[:, :256], qk[:, 256:], v, bits, out, do, lse, B, N, 8, sc, drop=drop, colbits=cb)
torch.cuda.synchronize()
print("ok")
