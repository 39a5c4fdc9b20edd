#!/usr/bin/env python
"""bench.py -- DESTR transformer-half hot path, fwd+bwd, images/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W             # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...   # reference algorithm on the host CPU cores

One "step" = one training pass of the hot path over one synthetic batch:
  encoder (6 layers, N = 25x42 = 1050 tokens of an 800x1333 image at stride 32) -> fine_pos ->
  decoder (6 layers, Q = 100 queries: self + pair + split cross attention) -> class/box heads ->
  matching cost matrix -> per-image linear sum assignment (device kernel, bit-identical to scipy) -> set loss ->
  backward -> (N>1: NCCL gradient all-reduce) -> AdamW step.  One CUDA graph replay per step.
Inputs are what the out-of-scope stages hand to the path: `reduce_dim(backbone(img))` features
(B,256,25,42), the padding mask, and the mini-detector's selected queries/centres (synthetic, seeded).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(workload="config2: DESTR transformer half fwd+bwd, 6 enc/6 dec, d=256, 8 heads, 800x1333 -> N=1050 tokens, "
                    "Q=100 queries, 91 classes, batch 8 per GPU",
           B=8, H=25, W=42, Q=100, C=91, L=6)
METRIC = "images/sec fwd+bwd at 800x1333 (transformer-half hot path)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def make_batch(rank: int, step: int, B: int, cfg=CFG, padded: bool = False):
    """Seeded synthetic inputs (CPU tensors)."""
    from oracle import destr_oracle as O  # generator of the synthetic targets only (data, not compute)
    g = torch.Generator().manual_seed(1234 + 1000 * rank + step)
    feats = torch.randn(B, 256, cfg["H"], cfg["W"], generator=g)
    mask = torch.zeros(B, cfg["H"], cfg["W"], dtype=torch.bool)
    if padded:
        for b in range(1, B):
            mask[b, :, cfg["W"] - 2 * b:] = True
    sel = torch.randn(B, cfg["Q"], 512, generator=g)
    centers = 0.05 + 0.9 * torch.rand(B, cfg["Q"], 2, generator=g)
    labels, boxes = O.make_targets(B, seed=100 * rank + step, max_t=40, num_cls=cfg["C"])
    return feats, mask, sel, centers, labels, boxes


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (oracle port, fp32 eager torch) on the host CPU cores
# ----------------------------------------------------------------------------------------------------
class _TorchDropper:
    """The reference's own dropout (nn.Dropout / F.dropout, p = 0.3 at every site) for the CPU arm's oracle hooks."""

    def __init__(self, p: float):
        self.p = p

    def rows(self, prefix, name, x):
        return torch.nn.functional.dropout(x, self.p, training=True)

    def attn(self, prefix, name, p):
        return torch.nn.functional.dropout(p, self.p, training=True)


class CpuReference:
    def __init__(self, cfg=CFG, dropout: float = 0.0):
        from oracle import destr_oracle as O
        self.O, self.cfg = O, cfg
        self.dropper = _TorchDropper(dropout) if dropout > 0 else None
        req = lambda sd: {k: v.requires_grad_() for k, v in sd.items()}
        self.enc, self.dec = req(O.make_encoder_weights(cfg["L"], 0)), req(O.make_decoder_weights(cfg["L"], 1))
        c, b = O.make_head_weights(cfg["C"], 2)
        self.cls, self.bbox = req(c), req(b)

    def step(self, batch):
        with self.O.dropout(self.dropper):
            return self._step(batch)

    def _step(self, batch):
        O, L = self.O, self.cfg["L"]
        feats, mask, sel, centers, labels, boxes = batch
        pos = O.sine_pos2d(mask)
        enc = O.encoder_forward(feats, mask, pos, self.enc, L)
        fine = O.fine_pos_tokens(enc, pos, self.enc)
        dec = O.decoder_forward(sel, enc.flatten(2).transpose(1, 2), mask.flatten(1), fine,
                                O.query_sine_embed(centers, 256), centers, self.dec, self.bbox, L)
        out = O.heads_forward(dec, centers, self.cls, self.bbox)
        with torch.no_grad():
            idx = O.hungarian_match(O.match_cost_blocks(out["pred_class"], out["pred_boxes"], labels, boxes,
                                                        0.5, 0.0, 0.5, with_l1=False))
        losses = O.set_criterion(out["pred_class"], out["pred_boxes"], labels, boxes, idx, self.cfg["C"])
        loss = 0.5 * losses["class"] + 0.0 * losses["bbox"] + 0.5 * losses["ciou"]
        for sd in (self.enc, self.dec, self.cls, self.bbox):
            for v in sd.values():
                v.grad = None
        loss.sum().backward()
        return float(loss.sum())


def _ref_harness():
    """oracle/ref_harness.py when the staged reference (oracle/_ref, see oracle/stage_ref.py) travelled with the repo."""
    from oracle import ref_harness as RH
    return RH if RH.available() else None


def time_cpu_reference(sample_b: int, steps: int, warmup: int, dropout: float = 0.0, budget_s: float = 200.0):
    """The reference's transformer-half training step on the host cores.  kind "reference": the UNMODIFIED reference
    modules (oracle/_ref: build_encoder / build_decoder / HungarianMatcherWoL1 / SetCriterion + AdamW) through
    oracle/ref_harness.py; kind "port" (only if the staged copy is missing): the oracle restatement.
    Each step is a bounded sample of the workload: batch `sample_b` of the config-2 images, halved until
    (steps + warmup) steps fit `budget_s` (estimated from the first warm-up step).
    -> (images/s, s/step, threads, kind, batch used, warm-up steps run)"""
    torch.set_num_threads(os.cpu_count() or 1)
    RH = _ref_harness()
    if RH is not None:
        tr = RH.RefTrainer(CFG, "cpu", dropout=dropout > 0)
        step, kind = (lambda b: float(tr.step(b).sum())), "reference"
    else:
        ref = CpuReference(dropout=dropout)
        step, kind = ref.step, "port"
    warmup = max(warmup, 1)
    t0 = time.perf_counter()
    step(make_batch(0, 0, sample_b))                      # warm-up step 1, also the cost estimate
    t1 = time.perf_counter() - t0
    b = sample_b
    while b > 1 and (steps + warmup - 1) * t1 * b / sample_b > budget_s:
        b //= 2
    for s in range(1, warmup):
        step(make_batch(0, s, b))
    t0 = time.perf_counter()
    for s in range(steps):
        step(make_batch(0, 100 + s, b))
    dt = time.perf_counter() - t0
    return b * steps / dt, dt / steps, torch.get_num_threads(), kind, b, warmup


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, spstep, cores, kind, b, warm = time_cpu_reference(CFG["B"], args.steps, args.warmup, 0.3 if args.dropout else 0.0)
    what = ("the reference's own modules (oracle/_ref: build_encoder, build_decoder, HungarianMatcherWoL1, SetCriterion, AdamW)"
            if kind == "reference" else "reference algorithm via oracle/destr_oracle.py")
    sample = (f"{args.steps} steps of batch {b} of the config-2 images per step ({'the full batch' if b == CFG['B'] else 'bounded sample of the batch of ' + str(CFG['B'])}); "
              f"{what}, fp32 eager torch, all host threads")
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CFG["workload"], "dropout": 0.3 if args.dropout else 0.0, "sample_batch": b},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def time_reference_on_gpu(dev, dropout: bool, steps: int = 10, warmup: int = 3):
    """SURVEY 8(d) "practical bar": the reference's own modules run by stock eager torch ON THIS GPU, same batch and
    step (fwd + matcher + loss + bwd + AdamW), inputs resident on the device, in fp32 and under autocast(bf16).
    None when the staged reference is absent."""
    RH = _ref_harness()
    if RH is None:
        return None
    out = {"what": "unmodified reference modules (oracle/_ref) through oracle/ref_harness.py, stock eager torch on this GPU, "
                   "batch 8, device-resident inputs, fwd + matcher (C.cpu() + scipy) + loss + bwd + AdamW",
           "steps": steps, "warmup": warmup, "dropout": 0.3 if dropout else 0.0}
    batches = [make_batch(0, s, CFG["B"]) for s in range(4)]
    for key, autocast in (("fp32", False), ("bf16", True)):
        tr = RH.RefTrainer(CFG, dev, dropout=dropout, autocast=autocast)
        ips, sps = RH.time_trainer(tr, batches, steps, warmup, on_device=True)
        out[key] = {"images_per_s": ips, "ms_per_step": sps * 1e3}
        tr2 = RH.RefTrainer(CFG, dev, dropout=dropout, autocast=autocast, optimizer=False)
        ips2, sps2 = RH.time_trainer(tr2, batches, steps, warmup, on_device=True)
        out[key]["fwd_bwd_no_optimizer"] = {"images_per_s": ips2, "ms_per_step": sps2 * 1e3}
        del tr, tr2
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def time_dominant_kernels(cfg, B, dev, iters: int = 20, dropout: float = 0.0):
    """CUDA-event duration of the encoder attention kernels at the workload shape, each launch alone with the
    L2 flushed (a 256 MB write) in between.  (Inside the CUDA-graph replay individual kernels cannot be
    bracketed by events; the ncu launch list in profiles/ gives their share of the step.)"""
    import math
    from object_detection_destr_b200 import ops
    N = cfg["H"] * cfg["W"]
    g = torch.Generator(device="cpu").manual_seed(0)
    qk = torch.randn(B * N, 512, generator=g).bfloat16().to(dev)
    v = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
    do = torch.randn(B * N, 256, generator=g).bfloat16().to(dev)
    bits = ops.pack_key_mask(None, B, N, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    scale = 1.0 / math.sqrt(32)
    out, lse = ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, scale)
    res = {}
    # same dropout setting as the timed step: the kernels then hash a keep-mask per probability
    drop = (torch.ones(1, dtype=torch.int32, device=dev), ops.drop_thr16(dropout), 0) if dropout > 0 else None
    rb, cb = ops.attn_dropout_bits(drop, B * 8, N, dev) if drop else (None, None)
    bwd = lambda: ops.enc_attn_bwd(qk[:, :256], qk[:, 256:], v, bits, out, do, lse, B, N, 8, scale, drop=drop, colbits=cb)
    from object_detection_destr_b200 import _lib
    for name, fn in (("destr_attn_dropout_bits", (lambda: ops.attn_dropout_bits(drop, B * 8, N, dev)) if drop else (lambda: None)),
                     ("destr_enc_attn_fwd", lambda: ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, scale, drop=drop, rowbits=rb)),
                     ("destr_enc_attn_bwd_op", bwd),     # all three launches of the op: prep + tcgen05 kernel + dQ convert
                     ("destr_enc_attn_bwd", bwd)):       # the tcgen05 kernel alone (debug knob 14 skips the two helpers)
        _lib.lib.destr_debug_knob(14, 1 if name == "destr_enc_attn_bwd" else 0)
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
        res[name] = tot / iters
    _lib.lib.destr_debug_knob(14, 0)
    return res



def run_ours(args):
    import torch.distributed as dist
    from object_detection_destr_b200 import _lib, ops
    from object_detection_destr_b200.encoder import disable_dropout
    from object_detection_destr_b200.hotpath import TransformerHalf
    from object_detection_destr_b200.engine import GraphedTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg, B = CFG, CFG["B"]

    torch.manual_seed(0)
    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=cfg["L"], num_decoder_blocks=cfg["L"],
                                      num_cls=cfg["C"]))
    if args.dropout:
        model.to(dev).train()      # the reference's defaults: p = 0.3 at every dropout site, applied inside the kernels
    else:
        disable_dropout(model).to(dev).train()
    opt = model.make_optimizer(lr=1e-5)  # AdamW: flat kernel over the runtime's parameter buffer + heads
    model.runtime().seed.fill_(7919 * rank + 1)  # dropout streams differ across ranks, like independent workers' RNGs
    weights = {"class": 0.5, "bbox": 0.0, "ciou": 0.5}  # arg_parser.py:41-61 defaults

    n_batches = 4
    host = [make_batch(rank, s, B) for s in range(n_batches)]
    pin = lambda t: t.pin_memory()
    host_pinned = [(pin(f), pin(m), pin(s), pin(c), l, b) for f, m, s, c, l, b in host]
    resident = [tuple(t.to(dev) for t in bt[:4]) + (bt[4], bt[5]) for bt in host_pinned]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_pinned[0][:4]) + \
        sum(l.numel() * 8 + b.numel() * 4 for l, b in zip(host_pinned[0][4], host_pinned[0][5]))

    eng = GraphedTrainStep(model, opt, B=B, H=cfg["H"], W=cfg["W"], Q=cfg["Q"], num_classes=cfg["C"], t_max=40,
                           cost_class=0.5, cost_ciou=0.5, loss_weights=weights, world=world, device=dev)
    eng.load_batch(*resident[0])
    if args.eager:
        for _ in range(3):
            eng.eager_step()
    else:
        eng.capture(warmup=3)

    def train_step(batch):
        eng.load_batch(*batch)
        return eng.eager_step() if args.eager else eng.step()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        torch.cuda.profiler.start()  # ncu --profile-from-start off: profile the timed region only
        for s in range(steps):
            fn(s)
        torch.cuda.profiler.stop()
        en.record()
        sync()
        ms = torch.tensor([st.elapsed_time(en)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- warm-up ----
    for s in range(max(args.warmup, 3)):
        train_step(resident[s % n_batches])
    # ---- device-resident timing (value) ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = _lib.launch_count
    ms = timed(lambda s: train_step(resident[s % n_batches]), args.steps)
    launches = eng.launches_per_step if not args.eager else (_lib.launch_count - launches0) // args.steps
    # ---- end-to-end timing: pinned host inputs, H2D every step, loss read back every step ----
    # Every step's inputs come from pinned host memory (H2D inside the timed region) and its loss + assignment status
    # are read back.  The feed is pipelined like a data loader with prefetch: batch s+1 is packed and its H2D copies are
    # enqueued on a copy stream while step s runs on the GPU.
    # Both feeds are pipelined by one step, like a training loop with a prefetching loader and asynchronous logging:
    # batch s+1 is packed and copied H2D while step s runs, and the loss + assignment status of step s are read back
    # (pinned, non-blocking D2H enqueued right behind the step) while step s+1 already runs.  EVERY step's inputs cross
    # the bus and EVERY step's result is read inside the timed region; timed() waits for the last read.
    rb_pending = []

    def e2e_drain(keep=0):
        loss = None
        while len(rb_pending) > keep:
            loss = eng.finish_readback(rb_pending.pop(0))      # loss (D2H) + status check (scipy would raise on NaN costs)
        return loss

    def e2e_step(s):
        if args.eager:
            loss = train_step(host_pinned[s % n_batches]).item()
            eng.raise_if_invalid()
            return loss
        eng.step_prefetched()
        rb_pending.append(eng.enqueue_readback())              # D2H read of THIS step's result, finished one step later
        eng.prefetch(*host_pinned[(s + 1) % n_batches])
        return e2e_drain(keep=1)
    if args.no_e2e:
        ms_e2e = float("nan")
    else:
        if not args.eager:
            eng.prefetch(*host_pinned[0])
        for s in range(2):
            e2e_step(s)
        e2e_drain()

        def e2e_run(s):
            e2e_step(s)
            if s == args.steps - 1:
                e2e_drain()                                    # the last step's read belongs to the timed region
        ms_e2e = timed(e2e_run, args.steps)
    clocks = sampler.stop() if sampler else None
    # ---- fwd + bwd without the optimizer step (SURVEY 8d reports both): a second graph over the same model ----
    ms_nopt = None
    if not args.eager and not args.no_e2e:
        eng2 = GraphedTrainStep(model, None, B=B, H=cfg["H"], W=cfg["W"], Q=cfg["Q"], num_classes=cfg["C"], t_max=40,
                                cost_class=0.5, cost_ciou=0.5, loss_weights=weights, world=world, device=dev)
        eng2.load_batch(*resident[0])
        eng2.capture(warmup=2)

        def nopt_step(s):
            eng2.load_batch(*resident[s % n_batches])
            return eng2.step()
        for s in range(3):
            nopt_step(s)
        ms_nopt = timed(nopt_step, args.steps)
        eng2.gA = eng2.gB = None
    kernel_ms = time_dominant_kernels(cfg, B, dev, dropout=0.3 if args.dropout else 0.0) if rank == 0 else None
    kernel_ms0 = time_dominant_kernels(cfg, B, dev) if (rank == 0 and args.dropout) else kernel_ms

    if rank == 0:
        pk = peaks()
        N = cfg["H"] * cfg["W"]
        fwd_flops = 4.0 * N * N * 256 * B                       # 4*N^2*d per image per layer
        bwd_flops = 2.0 * fwd_flops                             # algorithmic (dQ,dK,dV,dP), no recompute credit
        t_f, t_b = kernel_ms["destr_enc_attn_fwd"], kernel_ms["destr_enc_attn_bwd"]
        dom = "destr_enc_attn_bwd" if t_b >= t_f else "destr_enc_attn_fwd"
        ach = (bwd_flops / t_b if dom.endswith("bwd") else fwd_flops / t_f) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
                traffic = tj.get(dom + "@dropout", tj.get(dom)) if args.dropout else tj.get(dom)
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                "frac": ach / pk["tf_burst"], "traffic": traffic,
                "peak_source": pk["src"] + " (burst bf16 GEMM; kernel timed alone with CUDA events, L2 flushed between launches)",
                "launch_ms": t_b if dom.endswith("bwd") else t_f,
                "also": {"destr_enc_attn_fwd": {"launch_ms": t_f, "achieved": fwd_flops / t_f / 1e9,
                                                "frac": fwd_flops / t_f / 1e9 / pk["tf_burst"]},
                         "destr_enc_attn_bwd": {"launch_ms": t_b, "achieved": bwd_flops / t_b / 1e9,
                                                "frac": bwd_flops / t_b / 1e9 / pk["tf_burst"],
                                                "op_ms_with_prep_and_convert_launches": kernel_ms["destr_enc_attn_bwd_op"]},
                         "destr_attn_dropout_bits": {"launch_ms": kernel_ms.get("destr_attn_dropout_bits", 0.0),
                                                     "note": "mask bit matrices, once per layer (0 when dropout is off)"},
                         "without_dropout": {k: {"launch_ms": kernel_ms0[k], "frac": f / kernel_ms0[k] / 1e9 / pk["tf_burst"]}
                                             for k, f in (("destr_enc_attn_fwd", fwd_flops), ("destr_enc_attn_bwd", bwd_flops))}}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ips, spstep, cores, kind, b, _ = time_cpu_reference(2, 3, 1, 0.3 if args.dropout else 0.0, budget_s=30.0)
            cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": kind,
                   "sample": f"3 steps of batch {b} of the config-2 images ("
                             + ("the reference's own modules from oracle/_ref" if kind == "reference" else "oracle/destr_oracle.py")
                             + ", fp32 eager torch, all host threads)"}
        ref_gpu = None
        if world == 1 and not args.no_ref_gpu:
            ref_gpu = time_reference_on_gpu(dev, args.dropout)
            if ref_gpu is not None:
                for key in ("fp32", "bf16"):
                    ref_gpu[key]["our_speedup"] = (B * args.steps / (ms / 1e3)) / ref_gpu[key]["images_per_s"]
        imgs = B * world * args.steps
        nodes = getattr(eng, "graph_kernel_nodes", None) if not args.eager else None
        line = {"metric": METRIC, "value": imgs / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": cfg["workload"], "global_batch": B * world, "parallelism": f"dp{world}",
                           "step": "fwd + matcher (cost kernel + device LSAP, bit-identical to scipy) + fused set loss + bwd + grad all-reduce + flat AdamW, one CUDA graph",
                           "dropout": 0.3 if args.dropout else 0.0, "cuda_graphs": not args.eager, "l2": "4 rotating input batches; activations (~0.6 GB/step) exceed the 126 MB L2"},
                "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": int(h2d_bytes),
                        "d2h_bytes_per_step": 4 + 4 * B, "ms_per_step": ms_e2e / args.steps,
                        "feed": "pinned host batch s+1 packed + copied H2D on a copy stream while step s runs; loss + assignment status of step s read back (pinned, non-blocking) while step s+1 runs; every step's copies and reads are inside the timed region"},
                "gpu_launches": int(nodes if nodes else launches) * args.steps,
                "gpu_launches_per_step": int(nodes if nodes else launches),
                "gpu_launches_detail": {"graph_kernel_nodes_per_step": nodes, "own_kernels_per_step": int(launches),
                                        "note": "graph_kernel_nodes = every kernel node of the captured step graph (ours + cuBLAS/NCCL/torch); own_kernels = launched through libdestr_b200.so"},
                "fwd_bwd_no_optimizer": None if ms_nopt is None else {"value": imgs / (ms_nopt / 1e3), "unit": "images/s",
                                                                      "ms_per_step": ms_nopt / args.steps},
                "reference_gpu_eager": ref_gpu,
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL kernels are captured inside the CUDA graphs: release the graphs first, and leave without the
        # NCCL communicator teardown (destroy_process_group can block forever behind captured collectives).
        eng.gA = eng.gB = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs (debug / comparison)")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the end-to-end leg")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-on-this-GPU eager comparator")
    ap.add_argument("--no-dropout", dest="dropout", action="store_false",
                    help="p = 0 at every dropout site (the parity configuration) instead of the reference's training "
                         "default p = 0.3; applies to both arms")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
