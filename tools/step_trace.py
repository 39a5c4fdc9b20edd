"""Kernel timeline of one graph-replayed training step (torch.profiler / CUPTI): per-kernel start, duration, stream.
Prints the busy time per stream, the wall time of the step, and the largest gaps on the main stream."""
import os, sys, json
from argparse import Namespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from torch.profiler import profile, ProfilerActivity
from object_detection_destr_b200.encoder import disable_dropout
from object_detection_destr_b200.engine import GraphedTrainStep
from object_detection_destr_b200.hotpath import TransformerHalf

cfg, B = bench.CFG, bench.CFG["B"]
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:  # under torchrun: the data-parallel step (NCCL kernels show up in the timeline of rank 0)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=cfg["L"], num_decoder_blocks=cfg["L"], num_cls=cfg["C"]))
(model if "--dropout" in sys.argv else disable_dropout(model)).cuda().train()  # --dropout: the bench default (p = 0.3)
opt = model.make_optimizer(lr=1e-5)
eng = GraphedTrainStep(model, opt, B=B, H=cfg["H"], W=cfg["W"], Q=cfg["Q"], num_classes=cfg["C"], t_max=40, world=world)
bt = bench.make_batch(0, 0, B)
res = tuple(t.cuda() for t in bt[:4]) + (bt[4], bt[5])
eng.load_batch(*res)
eng.capture(warmup=3)
for _ in range(5):
    eng.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ks = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_index", 0)) for e in ev), key=lambda t: t[0])
ks = [k for k in ks if "Memcpy" not in k[2] and "Memset" not in k[2]] or ks
t0, t1 = ks[0][0], max(k[1] for k in ks)
print(f"kernels {len(ks)}  wall {(t1 - t0):.1f} us  sum {sum(k[1] - k[0] for k in ks):.1f} us")
# coverage: time when at least one kernel runs / idle
events = sorted([(k[0], 1) for k in ks] + [(k[1], -1) for k in ks])
busy, depth, last = 0.0, 0, t0
conc = {}
for t, d in events:
    if depth > 0:
        busy += t - last
    conc[depth] = conc.get(depth, 0.0) + (t - last)
    depth += d
    last = t
print(f"busy (>=1 kernel) {busy:.1f} us, idle {(t1 - t0) - busy:.1f} us; time at concurrency: " +
      ", ".join(f"{k}:{v:.0f}" for k, v in sorted(conc.items())))
# per-name totals of the time during which the kernel was the ONLY one running (critical-path proxy; partial overlaps
# count for their exclusive part)
import collections
solo = collections.defaultdict(float)
cnt = collections.defaultdict(int)
tot = collections.defaultdict(float)
bounds = sorted(set([k[0] for k in ks] + [k[1] for k in ks]))
import bisect
active = collections.defaultdict(list)
for idx, k in enumerate(ks):
    i0, i1 = bisect.bisect_left(bounds, k[0]), bisect.bisect_left(bounds, k[1])
    for j in range(i0, i1):
        active[j].append(idx)
    cnt[k[2][:60]] += 1
    tot[k[2][:60]] += k[1] - k[0]
for j, lst in active.items():
    if len(lst) == 1:
        solo[ks[lst[0]][2][:60]] += bounds[j + 1] - bounds[j]
print("exclusive time (only kernel running) / total time, by kernel:")
for n, v in sorted(solo.items(), key=lambda kv: -kv[1])[:60]:
    print(f"  {v:8.1f} / {tot[n]:8.1f} us x{cnt[n]:3d}  {n}")
# phases of the step, by landmark kernels
def first(sub):
    return next((k[0] for k in ks if sub in k[2]), None)
def last(sub):
    return max((k[1] for k in ks if sub in k[2]), default=None)
marks = [("start", t0), ("first enc_attn_fwd", first("enc_attn_fwd")), ("last enc_attn_fwd", last("enc_attn_fwd")),
         ("match cost", first("match_cost")), ("lsap end", last("lsap")), ("set loss end", last("set_loss")),
         ("first enc_attn_bwd", first("enc_attn_bwd")), ("last enc_attn_bwd", last("enc_attn_bwd")), ("end", t1)]
prev = t0
for n, t in marks:
    if t is None:
        continue
    print(f"  {n:22s} at {t - t0:8.1f} us  (+{t - prev:7.1f})")
    prev = t

if world > 1:
    if rank == 0:
        nc = [k for k in ks if "nccl" in k[2].lower()]
        print(f"NCCL kernels: {len(nc)}, total {sum(k[1] - k[0] for k in nc):.1f} us")
        for k in nc:
            print(f"  start {k[0] - t0:8.1f} us  dur {k[1] - k[0]:7.1f} us  {k[2][:70]}")
        last_compute = max(k[1] for k in ks if "nccl" not in k[2].lower() and "adamw" not in k[2].lower())
        print(f"last non-NCCL, non-AdamW kernel ends at {last_compute - t0:.1f} us; step ends at {t1 - t0:.1f} us")
        tail = [k for k in ks if k[0] >= last_compute - 1.0]
        for k in tail:
            print(f"  tail: start {k[0] - t0:8.1f} dur {k[1] - k[0]:7.1f}  {k[2][:70]}")
    sys.stdout.flush()
    os._exit(0)
