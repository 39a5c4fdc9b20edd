"""Diagnostic for tests/test_gpu_engine.py::test_overlapped_optimizer_step_equals_plain_step: parameter differences
between repeated runs of the overlapped (T) and plain (F) optimizer paths after 3 steps (EPS env: Adam eps)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch, bench
from object_detection_destr_b200.encoder import disable_dropout
from object_detection_destr_b200.engine import GraphedTrainStep
from object_detection_destr_b200.hotpath import TransformerHalf
cfg = dict(bench.CFG, B=2, L=2, H=10, W=14, Q=60)
batches = [bench.make_batch(0, s, 2, cfg, padded=True) for s in range(3)]
def run(overlap, sync_each=False):
    torch.manual_seed(0)
    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=2, num_decoder_blocks=2, num_cls=cfg["C"]))
    disable_dropout(model).cuda().train()
    opt = model.make_optimizer(lr=1e-3, eps=float(__import__('os').environ.get('EPS', '1e-8')))
    eng = GraphedTrainStep(model, opt, B=2, H=10, W=14, Q=60, num_classes=cfg["C"], t_max=40)
    eng.overlap_opt = overlap
    snaps = []
    for bt in batches:
        eng.load_batch(*bt); eng.eager_step()
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    P = model.runtime().P
    snaps = [(P.m32.clone(), P.g32.clone())] * 3
    return snaps, model.runtime().P
order = [("T1", True, False), ("F1", False, False), ("F2", False, False), ("F3", False, False), ("T2", True, False), ("F4", False, False)]
res = {k: run(o, s) for k, o, s in order}
P = res["T1"][1]
names = sorted(P.off.items(), key=lambda kv: kv[1][0])
def where(i):
    last = None
    for n, (o, *_r) in names:
        if o <= i: last = n
        else: break
    return last
for k, _, sy in order:
    dm = (res[k][0][2][0] - res["T1"][0][2][0]).abs()
    print(k, f"dparam {float(dm.max()):.2e} at {where(int(dm.argmax()))} n>2e-5: {int((dm > 2e-5).sum())}", flush=True)
