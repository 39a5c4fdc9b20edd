"""Tiny driver for `ncu --set full`: encoder attention fwd+bwd at the config-2 shape (B=8, N=1050, 8 heads)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from object_detection_destr_b200 import ops  # noqa: E402

B, N = int(os.environ.get("PB", 8)), int(os.environ.get("PN", 1050))
g = torch.Generator().manual_seed(0)
qk = torch.randn(B * N, 512, generator=g).bfloat16().cuda()
v = torch.randn(B * N, 256, generator=g).bfloat16().cuda()
do = torch.randn(B * N, 256, generator=g).bfloat16().cuda()
bits = ops.pack_key_mask(None, B, N, device=qk.device)
sc = 1 / math.sqrt(32)
for _ in range(3):
    out, lse = ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, sc)
    ops.enc_attn_bwd(qk[:, :256], qk[:, 256:], v, bits, out, do, lse, B, N, 8, sc)
torch.cuda.synchronize()
print("ok")
