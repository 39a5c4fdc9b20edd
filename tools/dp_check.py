"""torchrun --nproc-per-node 2 tools/dp_check.py : the overlapped bf16 gradient exchange of the runtime equals the mean of
the per-rank gradients (each rank steps on different data; compared against gradients gathered BEFORE the exchange)."""
import os, sys
from argparse import Namespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from object_detection_destr_b200.encoder import disable_dropout
from object_detection_destr_b200.engine import GraphedTrainStep
from object_detection_destr_b200.hotpath import TransformerHalf

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = dict(bench.CFG, B=2, L=2, H=10, W=14, Q=60)


def grads(dp: bool):
    torch.manual_seed(0)
    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=2, num_decoder_blocks=2, num_cls=cfg["C"]))
    disable_dropout(model).cuda().train()
    opt = model.make_optimizer(lr=0.0)
    eng = GraphedTrainStep(model, opt, B=2, H=10, W=14, Q=60, num_classes=cfg["C"], t_max=40, world=world if dp else 1)
    eng.load_batch(*bench.make_batch(rank, 3, 2, cfg, padded=True))
    eng.eager_step()
    P = model.runtime().P
    return P.g32.clone(), [p.grad.clone() for p in (model._cls_embed.weight, model._bbox_embed[2].bias)]


local_flat, local_heads = grads(False)
dp_flat, dp_heads = grads(True)
mean_flat = local_flat.clone()
dist.all_reduce(mean_flat)
mean_flat /= world
err = float((dp_flat - mean_flat).abs().max()) / (float(mean_flat.abs().max()) + 1e-12)
ok = err < 1e-2  # bf16 transport of the weight gradients
for a, b in zip(dp_heads, local_heads):
    m = b.clone(); dist.all_reduce(m); m /= world
    ok &= torch.allclose(a, m, rtol=1e-5, atol=1e-7)
same = [torch.zeros_like(dp_flat) for _ in range(world)]
dist.all_gather(same, dp_flat)
ok &= all(torch.equal(same[0], t) for t in same)
print(f"rank {rank}: rel err of exchanged flat gradient {err:.2e}; identical on all ranks; {'OK' if ok else 'FAIL'}", flush=True)
dist.barrier()
os._exit(0 if ok else 1)
