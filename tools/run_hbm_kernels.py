"""Runs the bandwidth-bound kernels of the hot path once each at the config-2 shapes (B=8, N=1050 -> 8400 token rows,
Q=100 -> 800 query rows; matcher also at config 5: B=64, Q=300), for `ncu --set full` captures digested by
tools/hbm_digest.py into profiles/r02_hbm_kernels.txt."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from object_detection_destr_b200 import ops, _lib
from oracle import destr_oracle as O
BF = torch.bfloat16
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device=dev)
M, Mq = 8400, 800
a, b, dy, res = (rn(M, 256).to(BF) for _ in range(4))
gam, bet = torch.ones(256, device=dev), torch.zeros(256, device=dev)
seed = torch.tensor([5], dtype=torch.int32, device=dev)
drop = (seed, ops.drop_thr16(0.3), 3)
x5, o1, o2 = rn(Mq, 512).to(BF), rn(Mq, 512).to(BF), rn(Mq, 1024).to(BF)
g5, b5 = torch.ones(512, device=dev), torch.zeros(512, device=dev)
pairs = torch.stack([torch.arange(100, device=dev, dtype=torch.int32).repeat(8), torch.randint(0, 100, (800,), generator=g, device=dev, dtype=torch.int32)], -1).view(8, 100, 2).contiguous()
big = rn(M, 512).to(BF)
dbias = torch.zeros(512, device=dev)
f = torch.relu(rn(Mq, 1024)).to(BF)
n = 25_000_000 // 4 * 4
pm, pg, pe, pv = (torch.zeros(n, device=dev) for _ in range(4))
ps = torch.zeros(n, dtype=BF, device=dev)
step = torch.ones(1, device=dev)


def matcher(B, Q):
    gl = torch.Generator().manual_seed(B)
    logits = torch.randn(B, Q, 91, generator=gl).to(dev)
    boxes = torch.cat([0.1 + 0.8 * torch.rand(B, Q, 2, generator=gl), 0.03 + 0.4 * torch.rand(B, Q, 2, generator=gl)], -1).to(dev)
    labels, tboxes = O.make_targets(B, seed=1, max_t=40, num_cls=91)
    sizes = [int(l.numel()) for l in labels]
    ids = torch.cat(labels).to(device=dev, dtype=torch.int32)
    tb = torch.cat(tboxes).to(dev)
    offs = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int32, device=dev)
    return lambda: ops.match_cost_blockdiag(logits, boxes, ids, tb, offs, sum(sizes), 0.5, 0.0, 0.5, False)


m8, m64 = matcher(8, 100), matcher(64, 300)
for it in range(2):
    y, mean, rstd = ops.add_layernorm(a, b, gam, bet, save_stats=True, drop=drop)
    ops.add_layernorm_bwd(dy, a, b, gam, mean, rstd, dbias=torch.zeros(256, device=dev), res_in=res, drop=drop)
    out, st = ops.dual_ln_mix(x5, o1, o2, pairs, g5, b5, g5, b5, 0.5, 100)
    ops.dual_ln_mix_bwd(x5, x5, o1, o2, pairs, g5, g5, st, 0.5, 100, head_major=True)
    ops.relu_bwd_colsum(big, None, dbias)
    ops.dropout_inplace(f, drop)
    ops.pos_mul_add(a, b, dy)
    m8()
    m64()
    _lib.call("destr_flat_adamw", pm.data_ptr(), pg.data_ptr(), pe.data_ptr(), pv.data_ptr(), ps.data_ptr(), n, 1e-4, 0.9,
              0.999, 1e-8, 0.01, step.data_ptr(), None, 0, 0, 1.0, ops._stream())
torch.cuda.synchronize()
