"""Generates tests/golden/query_select_golden.npz by running the REFERENCE's MiniDetector.get_topk_index and the
gathers of MiniDetector.forward (src/model/blocks/mini_detector.py:70-104, 142-170) on seeded inputs.
Run in the build container only (needs /root/reference):  python tests/golden/make_query_select_golden.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from src.model.blocks.mini_detector import MiniDetector  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(seed, B, N, C, k, pad, tie_free=True):
    """tie_free: re-seed until the reference's sort key (fp32 sigmoid of sigmoid, max over classes) has no duplicates
    among the valid positions, so torch.topk's order is fully specified and indices can be compared exactly.  With
    many classes the double sigmoid squeezes the maxima into ~1e5 distinct fp32 values and ties are unavoidable
    (tie_free=False): such a case pins the selected KEY VALUES, not the order inside a tie."""
    for _ in range(200):
        g = torch.Generator().manual_seed(seed)
        logits = torch.randn(B, N, C, generator=g) * 2.0
        mask = torch.zeros(B, N, dtype=torch.bool)
        for b in range(1, B):
            mask[b, N - pad * b:] = True
        scores = logits.sigmoid().masked_fill(mask.unsqueeze(-1), 0.0)  # what forward passes (:146-148)
        key = scores.sigmoid().max(-1).values                          # the reference's sort key (:79-80)
        ok = all(len(torch.unique(key[b][~mask[b]])) == int((~mask[b]).sum()) for b in range(B))
        if ok or not tie_free:
            break
        seed += 1000
    else:
        raise RuntimeError("no tie-free seed found")
    cls_feat = torch.randn(B, N, 16, generator=g)   # (feature width 16 instead of 256: the fixture stays small)
    reg_feat = torch.randn(B, N, 16, generator=g)
    coords = torch.rand(B, N, 4, generator=g)
    valid0 = int((~mask[0]).sum())
    avail_k = min(k, N, valid0)                                         # :153-154
    bi, idx = MiniDetector.get_topk_index(None, scores, k=avail_k, padding_mask=mask)
    feats = torch.concat([cls_feat, reg_feat], dim=-1)
    sel = feats[(bi.long(), idx)].reshape(B, avail_k, -1)               # :162-164
    cen = coords[..., :2][(bi.long(), idx)].reshape(B, avail_k, -1)     # :165-170
    out = dict(scores=scores.numpy(), mask=mask.numpy(), k=np.int64(avail_k), idx=idx.reshape(B, avail_k).numpy(),
               key=key.numpy(), tie_free=np.bool_(ok))
    if N <= 200:  # the gathers too (kept out of the large case: the fixture stays small)
        out.update(cls_feat=cls_feat.numpy(), reg_feat=reg_feat.numpy(), coords=coords.numpy(), sel=sel.numpy(),
                   cen=cen.numpy())
    return out


def main():
    out = {}
    #            seed  B   N    C   k   pad
    for i, a in enumerate([(1, 2, 150, 7, 100, 0), (2, 3, 150, 7, 100, 30), (3, 4, 96, 5, 60, 25),
                           (4, 2, 1050, 91, 100, 400, False), (5, 3, 40, 3, 100, 12)]):
        for name, v in case(*a).items():
            out[f"c{i}_{name}"] = v
    np.savez_compressed(os.path.join(HERE, "query_select_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
