import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from object_detection_destr_b200 import ops, _lib
import time_gemm as T
BF = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
for M in (8400, 67200, 1050):
    x = torch.randn(M, 256, generator=g, device="cuda").to(BF)
    f1 = torch.randn(M, 2048, generator=g, device="cuda").to(BF)
    for (N, K, a) in ((256, 256, x), (512, 256, x), (2048, 256, x), (256, 2048, f1), (3072, 256, x)):
        w = (torch.randn(N, K, generator=g, device="cuda") / 16).to(BF)
        b = torch.randn(N, device="cuda"); bh = b.to(BF)
        r = []
        for knob in (1, 2):
            _lib.lib.destr_debug_knob(17, knob)
            r.append(T.t(lambda: ops.gemm(a, w, bias=b)))
        _lib.lib.destr_debug_knob(17, 0)
        lib = T.t(lambda: torch.addmm(bh, a, w.t()))
        print(f"M={M:6d} N={N:5d} K={K:5d}: BN=128 {r[0]:7.1f} us  BN=256 {r[1]:7.1f} us  library {lib:7.1f} us", flush=True)
