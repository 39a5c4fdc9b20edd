"""Query selection of the mini-detector on the B200 kernels (reference src/model/blocks/mini_detector.py).

`get_topk_index` keeps the reference method's signature and return value; `select_queries` is the tail of
MiniDetector.forward (:142-170) -- top-k + padding fix-up + gathers -- as two small kernel launches (keys, rank + gather).  The only host-side
quantity is k = min(top_k, H*W, valid positions of image 0) (:153-154): it fixes the output SHAPE, so the reference
reads it back from the device; pass `valid0` when the padding mask was built on the host (as data loaders do) and no
synchronisation happens at all.
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Tuple

import torch
from torch import Tensor

from . import _lib, ops


class QuerySelection(NamedTuple):
    """(selected_objects (B,k,2D), selected_centers (B,k,2), topk_idx (B,k) int64, status (B,) int32 device tensor:
    non-zero = that image has no valid position -- the reference raises there; see check_status)."""
    selected_objects: Tensor
    selected_centers: Tensor
    topk_idx: Tensor
    status: Tensor


def _avail_k(top_k: int, N: int, mask: Optional[Tensor], valid0: Optional[int]) -> int:
    if valid0 is None:
        valid0 = N if mask is None else int((~mask[0].bool()).sum())  # device read-back, as in the reference
    return min(int(top_k), N, int(valid0))


def select_queries(scores: Tensor, mask: Optional[Tensor], cls_features: Tensor, reg_features: Tensor,
                   coords: Tensor, top_k: int, valid0: Optional[int] = None, want_bf16: bool = False) -> QuerySelection:
    """scores (B,N,C) fp32 = det_output_class as passed to get_topk_index (:156); mask (B,N) bool (True = padded);
    cls_features / reg_features (B,N,D) fp32; coords (B,N,4) fp32 (masked det_output_coord).
    -> QuerySelection(selected_objects (B,k,2D), selected_centers (B,k,2), topk_idx (B,k) int64, status (B,));
    selected_objects is bf16 when want_bf16 (the dtype the decoder kernels consume) else fp32.  Outputs are detached,
    as in the reference.  Nothing is read back from the device: `check_status(result.status)` does that on demand."""
    if not scores.is_cuda:
        raise RuntimeError("select_queries needs CUDA tensors (there is no CPU fallback)")
    B, N, C = scores.shape
    D = cls_features.shape[-1]
    k = _avail_k(top_k, N, mask, valid0)
    f = lambda t, n: ops._chk(t.detach().contiguous(), torch.float32, n)
    sc, cf, rf, co = f(scores, "scores"), f(cls_features, "cls_features"), f(reg_features, "reg_features"), f(coords, "coords")
    mk = None
    if mask is not None:
        mk = mask.reshape(B, N).contiguous()
        mk = mk.view(torch.uint8) if mk.dtype == torch.bool else mk.to(torch.uint8)
    dev = scores.device
    idx = torch.empty(B, k, dtype=torch.int64, device=dev)
    sel = torch.empty(B, k, 2 * D, dtype=torch.bfloat16 if want_bf16 else torch.float32, device=dev)
    cen = torch.empty(B, k, 2, dtype=torch.float32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    ws = torch.empty(B * N + B, dtype=torch.int32, device=dev)
    _lib.call("destr_select_queries", sc.data_ptr(), ops._ptr(mk), cf.data_ptr(), rf.data_ptr(), co.data_ptr(), B, N, C,
              D, k, idx.data_ptr(), None if want_bf16 else sel.data_ptr(), sel.data_ptr() if want_bf16 else None,
              cen.data_ptr(), status.data_ptr(), ws.data_ptr(), ops._stream())
    return QuerySelection(sel, cen, idx, status)


def check_status(status: Tensor) -> None:
    """Raise what the reference raises for an image without valid positions (ZeroDivisionError at :93).
    Reads `status` back from the device (a synchronisation)."""
    if bool((status != 0).any()):
        raise ZeroDivisionError("select_queries: an image has no valid (un-padded) position")


def get_topk_index(scores: Tensor, k: int, padding_mask: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """MiniDetector.get_topk_index (:70-104): -> (batch_idx int32 [B*k], idx int64 [B*k])."""
    B, N, _ = scores.shape
    z = torch.zeros(B, N, 4, dtype=torch.float32, device=scores.device)
    r = select_queries(scores, padding_mask, z, z, z, k, valid0=N)
    check_status(r.status)  # the reference method raises here (:93), and it synchronises anyway
    idx = r.topk_idx
    batch_idx = torch.arange(B, device=scores.device, dtype=torch.int32).repeat_interleave(k)
    return batch_idx, idx.flatten()
