"""BASELINE.json configs 4 and 5 (parity-test shapes, measured for the record; bench.py stays on config 2).

  config 4: encoder-only attention stress -- stride-16 features, N = 50x84 = 4200 tokens, batch 16, d = 256:
            attention kernels alone (fwd / bwd) and the whole 6-layer encoder forward and forward+backward.
  config 5: inference-only sweep -- 300 queries, batch 1..64, forward of the transformer half with the matcher
            (cost kernel + device assignment) in the loop; one CUDA graph per batch size; images/s.
"""
import math, os, sys, json
from argparse import Namespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from object_detection_destr_b200 import ops
from object_detection_destr_b200.encoder import disable_dropout
from object_detection_destr_b200.hotpath import TransformerHalf

dev = torch.device("cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(iters):
        fn()
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / iters


def config4():
    B, H, W = 16, 50, 84
    N = H * W
    g = torch.Generator().manual_seed(0)
    qk = torch.randn(B * N, 512, generator=g).bfloat16().to(dev)
    v = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
    do = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
    bits = ops.pack_key_mask(None, B, N, device=dev)
    sc = 1 / math.sqrt(32)
    out, lse = ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, sc)
    t_f = timeit(lambda: ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, sc))
    t_b = timeit(lambda: ops.enc_attn_bwd(qk[:, :256], qk[:, 256:], v, bits, out, do, lse, B, N, 8, sc))
    fl = 4.0 * N * N * 256 * B
    res = {"attn_fwd_ms": t_f, "attn_fwd_tflops": fl / t_f / 1e9, "attn_bwd_ms": t_b, "attn_bwd_tflops": 2 * fl / t_b / 1e9}
    del qk, v, do, out, lse
    torch.manual_seed(0)
    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=6, num_decoder_blocks=1, num_cls=91))
    disable_dropout(model).to(dev).train()
    rt = model.runtime()
    rt.dc = rt._drop_config()  # (what HotPathRuntime.forward() sets up: dropout is disabled here, p = 0 everywhere)
    rt._attn_bits = None
    x = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
    mask = torch.zeros(B, H, W, dtype=torch.bool, device=dev)
    _, pos = ops.sine_pos2d(mask, want_f32=False, want_bf16=True)
    pos = pos.view(B * N, 256)

    def enc_fwd():
        y, saved = x, []
        for l in range(6):
            y, sv = rt._enc_fwd(l, y, pos, bits, B, N)
            saved.append(sv)
        return y, saved

    def enc_fwd_bwd():
        y, saved = enc_fwd()
        rt.P.begin_backward()
        d = y
        for l in reversed(range(6)):
            d = rt._enc_bwd(l, d, saved[l], pos, bits, B, N)
        rt.P.end_backward()

    res["encoder_fwd_ms"] = timeit(lambda: enc_fwd(), iters=5)
    res["encoder_fwd_bwd_ms"] = timeit(enc_fwd_bwd, iters=5)
    res["encoder_fwd_images_per_s"] = B / res["encoder_fwd_ms"] * 1e3
    return res


def config5():
    from oracle import destr_oracle as O  # synthetic targets only
    cfg = dict(bench.CFG, Q=300)
    torch.manual_seed(0)
    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=6, num_decoder_blocks=6, num_cls=91))
    disable_dropout(model).to(dev).eval()
    out = {}
    for B in (1, 2, 4, 8, 16, 32, 64):
        feats, mask, sel, centers, labels, boxes = bench.make_batch(0, 0, B, cfg)
        feats, mask, sel, centers = (t.to(dev) for t in (feats, mask, sel, centers))
        sizes = [int(l.numel()) for l in labels]
        offs = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int32, device=dev)
        ids = torch.cat(labels).int().to(dev)
        tb = torch.cat(boxes).to(dev)
        cost = torch.empty(300 * max(sum(sizes), 1), device=dev)
        lsa_out = (torch.empty(B, 40, dtype=torch.int64, device=dev), torch.empty(B, 40, dtype=torch.int64, device=dev),
                   torch.empty(B, 40, dtype=torch.bool, device=dev), torch.empty(B, dtype=torch.int32, device=dev))

        def step(with_matcher):
            with torch.no_grad():
                o, _ = model(feats, mask, sel, centers)
                if with_matcher:
                    ops.match_cost_blockdiag(o["pred_class"], o["pred_boxes"], ids, tb, offs, sum(sizes), 0.5, 0.0, 0.5, False,
                                             out=cost)
                    ops.lsap_blockdiag(cost, offs, B, 300, 40, 40, out=lsa_out)
            return o

        row = {}
        for with_matcher in (False, True):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    step(with_matcher)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                step(with_matcher)
            ms = timeit(gr.replay, iters=10)
            row["with_matcher" if with_matcher else "forward_only"] = {"ms": ms, "images_per_s": B / ms * 1e3}
            del gr
        out[str(B)] = row
        model._rt.saved = None
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    res = {}
    if which in ("all", "4"):
        res["config4"] = config4()
    if which in ("all", "5"):
        res["config5"] = config5()
    print(json.dumps(res, indent=1))
