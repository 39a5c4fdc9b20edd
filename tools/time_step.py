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
from object_detection_destr_b200 import ops
a8, b8 = torch.zeros(4096, device="cuda"), torch.zeros(4096, device="cuda")
hp = torch.zeros(4096, pin_memory=True)
res = tuple(t.cuda() for t in bt[:4]) + (bt[4], bt[5])
variants = {"replay only": lambda: eng.step(),
            "tiny kernel + replay": lambda: (ops.copy_many([a8], [b8]), eng.step()),
            "tiny H2D + replay": lambda: (a8.copy_(hp, non_blocking=True), eng.step()),
            "load_batch + replay": lambda: (eng.load_batch(*res), eng.step())}
for name, fn in variants.items():
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(100):
            fn()
        en.record()
        torch.cuda.synchronize()
        best = min(best, st.elapsed_time(en) / 100)
    print(f"{name:24s} step {best:.4f} ms  ({B / best * 1e3:.0f} img/s)  kernel nodes {eng.graph_kernel_nodes}", flush=True)
