"""Where one training step goes: graph A (forward + cost), host assignment, graph B (loss + backward + AdamW)."""
import os, sys, time
from argparse import Namespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from object_detection_destr_b200.encoder import disable_dropout
from object_detection_destr_b200.engine import GraphedTrainStep
from object_detection_destr_b200.hotpath import TransformerHalf

cfg, B = bench.CFG, bench.CFG["B"]
torch.manual_seed(0)
model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=cfg["L"], num_decoder_blocks=cfg["L"], num_cls=cfg["C"]))
disable_dropout(model).cuda().train()
opt = model.make_optimizer(lr=1e-5)
eng = GraphedTrainStep(model, opt, B=B, H=cfg["H"], W=cfg["W"], Q=cfg["Q"], num_classes=cfg["C"], t_max=40,
                       gpu_lsa="--host-lsa" not in sys.argv)
batches = [bench.make_batch(0, s, B) for s in range(4)]
res = [tuple(t.cuda() for t in bt[:4]) + (bt[4], bt[5]) for bt in batches]
eng.load_batch(*res[0])
eng.capture(warmup=3)
for s in range(5):
    eng.load_batch(*res[s % 4]); eng.step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tA = tH = tB = tL = 0.0
n = 30
for s in range(n):
    t0 = time.perf_counter()
    eng.load_batch(*res[s % 4])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ev[0].record(); eng.gA.replay(); ev[1].record()
    torch.cuda.current_stream().synchronize()
    t2 = time.perf_counter()
    if eng.gB is not None:
        eng._assign()
    t3 = time.perf_counter()
    ev[2].record()
    if eng.gB is not None:
        eng.gB.replay()
    ev[3].record()
    torch.cuda.synchronize()
    tL += t1 - t0; tA += ev[0].elapsed_time(ev[1]); tH += (t3 - t2) * 1e3; tB += ev[2].elapsed_time(ev[3])
print(f"load_batch {tL/n*1e3:.3f} ms | graph A {tA/n:.3f} ms | host assign {tH/n:.3f} ms | graph B {tB/n:.3f} ms")
