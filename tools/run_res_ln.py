"""Runs the fused GEMM + residual + LayerNorm kernels once each at the encoder shapes (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from object_detection_destr_b200 import ops
BF = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
M = 8400
rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
x, x2, f1 = rn(M, 256).to(BF), rn(M, 256).to(BF), torch.relu(rn(M, 2048)).to(BF)
w256, wfc2 = (rn(256, 256) / 16).to(BF), (rn(256, 2048) / 45).to(BF)
b256, gam, bet = rn(256), torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
seed = torch.tensor([3], dtype=torch.int32, device="cuda")
drop = (seed, ops.drop_thr16(0.3), 1)
for _ in range(3):
    ops.gemm_res_ln(x, w256, b256, x2, gam, bet, drop=drop)
    ops.gemm_res_ln(f1, wfc2, b256, x2, gam, bet, drop=drop, res2=x, gamma2=gam, beta2=bet)
    ops.gemm(x, w256, bias=b256)
torch.cuda.synchronize()
