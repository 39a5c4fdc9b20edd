"""Times the device LSAP at the bench shape (B = 8, Q = 100, T_i ~ U{1..40}) on costs of two kinds: random (distinct
optimum quickly found) and near-tied (what a freshly initialised model produces: long augmenting paths)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from object_detection_destr_b200 import ops
B, Q, TM = 8, 100, 40
g = torch.Generator().manual_seed(0)
sizes = [int(torch.randint(1, TM + 1, (1,), generator=g)) for _ in range(B)]
sizes[0] = TM
offs = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int32).cuda()
for kind in ("random", "near-tied"):
    blocks = []
    for t in sizes:
        c = torch.rand(Q, t, generator=g)
        if kind == "near-tied":
            c = 0.5 + 1e-3 * c + 0.2 * torch.rand(1, t, generator=g)
        blocks.append(c.reshape(-1))
    flat = torch.cat(blocks).cuda()
    out = ops.lsap_blockdiag(flat, offs, B, Q, TM)
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(20):
        ops.lsap_blockdiag(flat, offs, B, Q, TM, out=out)
    en.record(); torch.cuda.synchronize()
    print(f"lsap {kind}: {st.elapsed_time(en) / 20 * 1e3:.1f} us  (sizes {sizes})")
