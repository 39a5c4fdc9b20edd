"""Times destr_linear_bias_relu_dropout against the library path (cuBLASLt bias+ReLU GEMM, then destr_dropout_inplace)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from object_detection_destr_b200 import ops  # noqa: E402


def timeit(fn, iters=30):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1000


for M, N, K in ((8400, 2048, 256), (800, 1024, 256)):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / 16).bfloat16().cuda()
    b = torch.randn(N, generator=g).cuda()
    bb = b.bfloat16()
    drop = (torch.ones(1, dtype=torch.int32, device="cuda"), ops.drop_thr16(0.3), 3)
    lib = lambda: ops.dropout_inplace(torch._addmm_activation(bb, x, w.t(), use_gelu=False), drop)
    ours = lambda: ops.linear_bias_relu_dropout(x, w, b, drop)
    ours0 = lambda: ops.linear_bias_relu_dropout(x, w, b, None)
    print(f"M={M} N={N} K={K}: cuBLASLt+dropout pass {timeit(lib):.1f} us | fused tcgen05 {timeit(ours):.1f} us "
          f"(no dropout {timeit(ours0):.1f} us; output write alone = {2 * M * N / 6.5e6:.1f} us at 6.5 TB/s)")
