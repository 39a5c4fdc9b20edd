"""Times the graph-replayed training step at the bench configuration (CUDA events over 100 replays, after 10).
usage: [DESTR_*=..] python tools/time_step.py [--no-dropout]"""
import os, sys
from argparse import Namespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from object_detection_destr_b200.encoder import disable_dropout
from object_detection_destr_b200.engine import GraphedTrainStep
from object_detection_destr_b200.hotpath import TransformerHalf

cfg, B = bench.CFG, bench.CFG["B"]
from object_detection_destr_b200 import _lib
for kv in filter(None, os.environ.get("KNOBS", "").split(",")):  # e.g. KNOBS=17=2 (debug knobs)
    _lib.lib.destr_debug_knob(int(kv.split("=")[0]), int(kv.split("=")[1]))
torch.manual_seed(0)
model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=cfg["L"], num_decoder_blocks=cfg["L"], num_cls=cfg["C"]))
(disable_dropout(model) if "--no-dropout" in sys.argv else model).cuda().train()
opt = model.make_optimizer(lr=1e-5)
eng = GraphedTrainStep(model, opt, B=B, H=cfg["H"], W=cfg["W"], Q=cfg["Q"], num_classes=cfg["C"], t_max=40)
bt = bench.make_batch(0, 0, B)
eng.load_batch(*(tuple(t.cuda() for t in bt[:4]) + (bt[4], bt[5])))
eng.capture(warmup=3)
for _ in range(10):
    eng.step()
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(100):
        eng.step()
    en.record()
    torch.cuda.synchronize()
    best = min(best, st.elapsed_time(en) / 100)
print(f"step {best:.4f} ms  ({B / best * 1e3:.0f} img/s)  kernel nodes {eng.graph_kernel_nodes}", flush=True)
