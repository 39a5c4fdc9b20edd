"""Micro-timings of the small HBM/latency-bound kernels at config-2 sizes (CUDA events, 50 iterations each)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from object_detection_destr_b200 import ops
dev = "cuda"
def timeit(fn, iters=20):
    """20 back-to-back launches captured in a CUDA graph (no Python / launch overhead in the number)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(iters): fn()
    gr.replay(); torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(5): gr.replay()
    en.record(); torch.cuda.synchronize()
    return st.elapsed_time(en) / (5 * iters) * 1e3
g = torch.Generator().manual_seed(0)
for M, D in ((8400, 256), (800, 256), (800, 512)):
    a, b, dy = (torch.randn(M, D, generator=g).bfloat16().to(dev) for _ in range(3))
    gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    y, mean, rstd = ops.add_layernorm(a, b, gam, bet, save_stats=True)
    dg, db, dbias = (torch.zeros(D, device=dev) for _ in range(3))
    print(f"add_ln_fwd M={M} D={D}: {timeit(lambda: ops.add_layernorm(a, b, gam, bet, save_stats=True)):.1f} us")
    print(f"add_ln_bwd M={M} D={D}: {timeit(lambda: ops.add_layernorm_bwd(dy, a, b, gam, mean, rstd, dgamma=dg, dbeta=db, dbias=dbias)):.1f} us")
for M, C in ((8400, 2048), (8400, 512), (8400, 256), (800, 1024)):
    dy, h = (torch.randn(M, C, generator=g).bfloat16().to(dev) for _ in range(2))
    dbias = torch.zeros(C, device=dev)
    print(f"relu_bwd_colsum M={M} C={C}: {timeit(lambda: ops.relu_bwd_colsum(dy, h, dbias)):.1f} us; colsum only {timeit(lambda: ops.relu_bwd_colsum(dy, None, dbias)):.1f} us")
