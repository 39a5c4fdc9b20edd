"""Times the tcgen05 GEMM family (csrc/gemm_tc.cu) against the library path it replaces, at the encoder shapes of
config 2 (M = 8400 tokens).  CUDA events, L2 flushed between launches, mean of 20.  Prints one line per case."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from object_detection_destr_b200 import ops  # noqa: E402

BF = torch.bfloat16
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, iters=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        fn()
        en.record()
        en.synchronize()
        tot += st.elapsed_time(en)
    return tot / iters * 1e3  # us


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 8400
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g, device=dev)
    x = rn(M, 256).to(BF)
    x2 = rn(M, 256).to(BF)
    pos = rn(M, 256).to(BF)
    f1 = torch.relu(rn(M, 2048)).to(BF)
    W = lambda n, k: (rn(n, k) / k ** 0.5).to(BF)
    w256, w512, w2048, wfc2 = W(256, 256), W(512, 256), W(2048, 256), W(256, 2048)
    b256, b512, b2048 = rn(256), rn(512), rn(2048)
    gam, bet = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    seed = torch.tensor([3], dtype=torch.int32, device=dev)
    drop = (seed, ops.drop_thr16(0.3), 1)
    rows = []

    def case(name, ours, lib):
        a, b = t(ours), t(lib)
        rows.append((name, a, b))
        print(f"{name:58s} ours {a:7.1f} us   library path {b:7.1f} us   x{b / a:.2f}", flush=True)

    b256h, b512h, b2048h = b256.to(BF), b512.to(BF), b2048.to(BF)
    case("linear 256->256 + bias", lambda: ops.gemm(x, w256, bias=b256), lambda: torch.addmm(b256h, x, w256.t()))
    case("linear 256->512 + bias", lambda: ops.gemm(x, w512, bias=b512), lambda: torch.addmm(b512h, x, w512.t()))
    case("linear 256->2048 + bias + relu", lambda: ops.gemm(x, w2048, bias=b2048, relu=True),
         lambda: torch._addmm_activation(b2048h, x, w2048.t(), use_gelu=False))
    case("x + pos*(h W^T + b)   [GEMM + pos_mul_add]", lambda: ops.gemm(x, w256, bias=b256, mul=pos, add=x2),
         lambda: ops.pos_mul_add(x2, pos, torch.addmm(b256h, x, w256.t())))
    case("LN(x + drop(a Wo^T + b))   [GEMM + add_layernorm]",
         lambda: ops.gemm_res_ln(x, w256, b256, x2, gam, bet, drop=drop),
         lambda: ops.add_layernorm(x2, torch.addmm(b256h, x, w256.t()), gam, bet, save_stats=True, drop=drop))
    case("LN(x + LN2(x1 + drop(f1 W2^T + b)))   [GEMM + 2 add_layernorm]",
         lambda: ops.gemm_res_ln(f1, wfc2, b256, x2, gam, bet, drop=drop, res2=x, gamma2=gam, beta2=bet),
         lambda: ops.add_layernorm(x, ops.add_layernorm(x2, torch.addmm(b256h, f1, wfc2.t()), gam, bet, save_stats=True, drop=drop)[0],
                                   gam, bet, save_stats=True))
    dbias = torch.zeros(2048, device=dev)
    case("dpre = relu_bwd(d2 fc2_w) + colsum   [GEMM + relu_bwd_colsum]",
         lambda: ops.gemm_relu_bwd(x, wfc2, f1, 1 / 0.7, dbias),
         lambda: ops.relu_bwd_colsum(torch.mm(x, wfc2), f1, dbias, scale=1 / 0.7))
    case("dx1 = d2s + dpre fc1_w   (K=2048, [K,N] weight)", lambda: ops.gemm(f1, w2048, b_kn=True, add=x2),
         lambda: torch.addmm(x2, f1, w2048))
    case("da = d1 out_w   (K=256, [K,N] weight)", lambda: ops.gemm(x, w256, b_kn=True), lambda: torch.mm(x, w256))
    dqk = rn(M, 512).to(BF)
    case("ds = (dqk Wqk)*pos, dx += dqk Wqk   [GEMM + pos_mul_add_bwd_acc]",
         lambda: ops.gemm(dqk, w512, b_kn=True, mul=pos, add2=x2, out2=True),
         lambda: ops.pos_mul_add_bwd_acc(torch.mm(dqk, w512), pos, x2))
    from object_detection_destr_b200 import _lib
    for (no, ki, dy_, x_) in ((2048, 256, f1, x), (256, 2048, x, f1), (512, 256, dqk, x), (256, 256, x, x2)):
        dw = torch.zeros(no, ki, device=dev)
        outb = torch.empty(no, ki, dtype=BF, device=dev)
        case(f"dW {no}x{ki} = dY^T X (split-K, fp32 red)", lambda: ops.gemm_dw(dy_, x_, dw),
             lambda: torch.mm(dy_.t(), x_, out=outb))
        for S in (1, 2, 4, 8, 16, 32):
            _lib.lib.destr_debug_knob(15, S)
            print(f"      split {S:3d}: {t(lambda: ops.gemm_dw(dy_, x_, dw)):7.1f} us", flush=True)
        _lib.lib.destr_debug_knob(15, 0)


if __name__ == "__main__":
    main()
