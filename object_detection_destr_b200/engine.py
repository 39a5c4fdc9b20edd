"""CUDA-graph training step for the hot path.

The eager step issues ~1000 small launches and is CPU-bound; the engine captures the WHOLE step into one CUDA
graph:

    forward (encoder, decoder, heads) -> block-diagonal cost kernel -> per-image assignment ON THE DEVICE
    (ops.lsap_blockdiag, bit-identical to scipy) -> fused set loss (+ its gradients) -> backward ->
    [NCCL gradient all-reduce] -> flat AdamW

so a step is one graph replay with no host round trip.  `gpu_lsa=False` keeps the reference's own arrangement
(`C.cpu()` + scipy linear_sum_assignment, matcher.py:107-112) as two graphs around a host assignment; the two
modes are compared in tests/test_gpu_engine.py.

All shapes are static: targets are padded to `t_max` per image (the cost kernel, the assignment and the loss
read the true counts from device tensors), so one capture serves every batch.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from . import ops
from .matcher import _ciou_cost_batched, _cxcyhw_to_xyxy


def set_loss_static(logits, boxes, tl, tb, pi, ti, valid, num_classes: int):
    """SetCriterion.forward (criterion.py:29-79) on static, padded shapes (no host branches).
    logits (B,Q,C), boxes (B,Q,4) cxcyhw; tl (B,Tm) int64, tb (B,Tm,4) padded targets;
    pi (B,n) matched query index (Q in padded slots), ti (B,n) matched target index, valid (B,n) bool."""
    B, Q, _ = logits.shape
    n = pi.shape[1]
    logits, boxes = logits.float(), boxes.float()
    tcls = torch.ones(B, Q + 1, dtype=torch.int64, device=logits.device)
    tcls.scatter_(1, pi, tl.gather(1, ti))
    onehot = F.one_hot(tcls[:, :Q], num_classes).to(logits.dtype)
    prob = logits.sigmoid()
    ce = F.binary_cross_entropy_with_logits(logits, onehot, reduction="none")
    pt = prob * onehot + (1 - prob) * (1 - onehot)
    focal = (0.25 * onehot + 0.75 * (1 - onehot)) * ce * (1 - pt) ** 2
    loss_cls = (focal.mean(2).sum(1) / Q).mean()
    pic = pi.clamp(max=Q - 1)
    pb = _cxcyhw_to_xyxy(boxes).gather(1, pic[..., None].expand(B, n, 4))
    gb = tb.gather(1, ti[..., None].expand(B, n, 4))
    vf = valid.to(logits.dtype)
    cnt = vf.sum(1).clamp(min=1)
    has = (vf.sum(1) > 0).to(logits.dtype)
    l1 = ((pb - gb).abs().sum(-1) * vf).sum(1) / (4 * cnt)
    neutral = 0.25 + 0.25 * (torch.arange(4, device=logits.device) >= 2).to(logits.dtype)  # no H2D (graph capture)
    ci = _ciou_cost_batched(torch.where(valid[..., None], pb, neutral), torch.where(valid[..., None], gb, neutral + 0.1))
    ciou = (ci * (vf[:, :, None] * vf[:, None, :])).sum((1, 2)) / (cnt * cnt)
    denom = has.sum().clamp(min=1)
    return {"class": loss_cls, "bbox": (l1 * has).sum() / denom, "ciou": (ciou * has).sum() / denom}


class _TargetBlock:
    """The five target arrays of a batch in ONE byte buffer (pinned host or device), 16-byte aligned segments:
    [ids int32 B*tm | offs int32 B+1 | pad | boxes fp32 B*tm*4 | padded labels int64 B*tm | padded boxes fp32 B*tm*4].
    One H2D copy moves them all; the typed views are what the packer writes / the hand-over kernel reads."""

    def __init__(self, B: int, tm: int, device=None, pin: bool = True):
        al = lambda n: (n + 15) // 16 * 16
        o_i, n_i = 0, 4 * (B * tm + B + 1)
        o_f, n_f = al(n_i), 16 * B * tm
        o_l, n_l = o_f + al(n_f), 8 * B * tm
        o_b, n_b = o_l + al(n_l), 16 * B * tm
        total = o_b + al(n_b)
        self.buf = torch.empty(total, dtype=torch.uint8, pin_memory=pin) if device is None else \
            torch.empty(total, dtype=torch.uint8, device=device)
        v = lambda o, n, dt: self.buf[o:o + n].view(dt)
        self.ints = v(o_i, n_i, torch.int32)                       # ids followed by offsets (the packer's layout)
        self.ids, self.offs = self.ints[:B * tm], self.ints[B * tm:]
        self.flt = v(o_f, n_f, torch.float32)
        self.tboxes = self.flt.view(-1, 4)
        self.tl = v(o_l, n_l, torch.int64).view(B, tm)
        self.tb = v(o_b, n_b, torch.float32).view(B, tm, 4)

    def views(self):
        """In the order of GraphedTrainStep._statics()[4:]."""
        return [self.ids, self.offs, self.tboxes, self.tl, self.tb]


_CAPTURE_STREAMS = {}


def _capture_stream(device):
    """ONE high-priority capture stream per device, shared by every capture of the process (as torch.cuda.graph shares
    its default capture stream): autograd's gradient accumulators stay bound to the stream they were first used on, and
    a second engine capturing on a fresh stream would make the capture depend on that stream's uncaptured work."""
    if os.environ.get("DESTR_PRIO", "1") == "0":
        return None
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _CAPTURE_STREAMS:
        _CAPTURE_STREAMS[key] = torch.cuda.Stream(device=key, priority=-1)
    return _CAPTURE_STREAMS[key]


def _count_kernel_nodes(graphs) -> Optional[int]:
    """Kernel nodes of captured CUDA graphs (cudaGraphGetNodes on the raw cudaGraph_t): the number of GPU kernels one
    replay launches, library kernels included.  None when the raw graph is not reachable."""
    try:
        from cuda.bindings import runtime as rt
        total = 0
        for g in graphs:
            raw = rt.cudaGraph_t(int(g.raw_cuda_graph()))
            err, _, n = rt.cudaGraphGetNodes(raw, 0)
            if int(err) != 0:
                return None
            err, nodes, n = rt.cudaGraphGetNodes(raw, n)
            if int(err) != 0:
                return None
            for nd in nodes[:n]:
                err, ty = rt.cudaGraphNodeGetType(nd)
                if int(err) == 0 and ty == rt.cudaGraphNodeType.cudaGraphNodeTypeKernel:
                    total += 1
        return total
    except Exception:
        return None


class GraphedTrainStep:
    """One training step of the hot path as two CUDA graphs (see module docstring)."""

    def __init__(self, model, optimizer, *, B: int, H: int, W: int, Q: int, num_classes: int, t_max: int = 40,
                 cost_class: float = 0.5, cost_ciou: float = 0.5, loss_weights: Optional[Dict[str, float]] = None,
                 world: int = 1, device=None, fused_loss: bool = True, gpu_lsa: bool = True):
        """optimizer=None: the step ends after backward (+ gradient exchange) -- the "fwd+bwd" figure of SURVEY 8(d)."""
        self.model, self.opt = model, optimizer
        self.B, self.Q, self.C, self.t_max, self.world = B, Q, num_classes, t_max, world
        self.wc, self.wi = cost_class, cost_ciou
        self.lw = loss_weights or {"class": 0.5, "bbox": 0.0, "ciou": 0.5}
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        n = min(Q, t_max)
        self.n = n
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
        self.s_feats, self.s_mask = z(B, 256, H, W), z(B, H, W, dt=torch.bool)
        self.s_sel, self.s_centers = z(B, Q, 512), z(B, Q, 2) + 0.5
        self.s_ids, self.s_tboxes = z(B * t_max, dt=torch.int32), z(B * t_max, 4)
        self.s_offs = z(B + 1, dt=torch.int32)
        self.s_cost = z(Q * B * t_max)
        self.h_cost = torch.empty(Q * B * t_max, dtype=torch.float32, pin_memory=True)
        # loss-side statics: padded per-image targets and the matching
        self.s_tl, self.s_tb = z(B, t_max, dt=torch.int64) + 1, z(B, t_max, 4)
        self.s_pi, self.s_ti, self.s_valid = z(B, n, dt=torch.int64) + Q, z(B, n, dt=torch.int64), z(B, n, dt=torch.bool)
        self.h_idx = torch.empty(3, B, n, dtype=torch.int64, pin_memory=True)
        self.h_blk = _TargetBlock(B, t_max)               # pinned: packed by the host
        self.d_blk = _TargetBlock(B, t_max, device=dev)   # its device copy (one H2D), scattered by the hand-over kernel
        self.h_tgt_i, self.h_tgt_f, self.h_tl, self.h_tb = self.h_blk.ints, self.h_blk.flt, self.h_blk.tl, self.h_blk.tb
        self.params = [p for p in model.parameters()]
        if world > 1 and getattr(model, "use_runtime", False):
            model.runtime().P.enable_data_parallel(world)  # gradient exchange overlapped with the encoder backward
        self.fused_loss = fused_loss
        self.gpu_lsa = gpu_lsa
        self.overlap_opt = True  # FlatAdamW: decoder share of the step under the encoder backward
        self.s_status = z(B, dt=torch.int32)
        self.loss_ws = ops.set_loss_workspace(B, dev)
        self.losses = None
        self.gA: Optional[torch.cuda.CUDAGraph] = None
        self.gB: Optional[torch.cuda.CUDAGraph] = None
        self.loss = None
        self.out = None
        self.sizes: List[int] = []
        self.launches_per_step = 0
        self.graph_kernel_nodes = None  # kernel nodes of the captured graph(s): every GPU kernel of a step, ours and libraries'
        self._slots = None
        self._loaded = None  # event: the H2D copies of the last load_batch() have left the pinned staging arrays

    # ---- pieces (also run eagerly during warm-up) ----
    def _forward(self):
        rt = getattr(self.model, "_rt", None)
        if rt is not None and any(rt._drop_config().values()):
            rt.seed.add_(1)  # fresh dropout masks every step (device-side counter: also under graph replay)
        out, _ = self.model(self.s_feats, self.s_mask, self.s_sel, self.s_centers)
        with torch.no_grad():
            ops.match_cost_blockdiag(out["pred_class"].detach(), out["pred_boxes"].detach(), self.s_ids, self.s_tboxes,
                                     self.s_offs, self.B * self.t_max, self.wc, 0.0, self.wi, False, out=self.s_cost)
            if self.gpu_lsa:
                ops.lsap_blockdiag(self.s_cost, self.s_offs, self.B, self.Q, self.t_max, self.n,
                                   out=(self.s_pi, self.s_ti, self.s_valid, self.s_status))
            else:
                self.h_cost.copy_(self.s_cost, non_blocking=True)
        return out

    def _backward(self, out):
        if self.fused_loss:  # one kernel: the three losses, their weighted total and its gradients
            loss, self.losses = ops.set_loss(out["pred_class"], out["pred_boxes"], self.s_tl, self.s_tb, self.s_pi,
                                             self.s_ti, self.s_valid, (self.lw["class"], self.lw["bbox"], self.lw["ciou"]),
                                             self.loss_ws)
        else:  # batched torch restatement (autograd): the cross-check of the fused kernel
            losses = set_loss_static(out["pred_class"], out["pred_boxes"], self.s_tl, self.s_tb, self.s_pi, self.s_ti,
                                     self.s_valid, self.C)
            loss = sum(self.lw[k] * losses[k] for k in self.lw)
        # the decoder's share of the optimizer step overlaps the encoder backward (runtime.FlatAdamW.enable_overlap);
        # set per call: another engine over the same model (e.g. the fwd+bwd-only graph) must not step
        rt = getattr(self.model, "_rt", None)
        if rt is not None:
            rt.P.early_opt = self.opt._early if (self.overlap_opt and hasattr(self.opt, "_early")) else None
        loss.backward()
        if self.world > 1:  # the one exchange of the step: mean all-reduce of the gradients (NCCL over NVLink)
            from .dataparallel import allreduce_mean_, grads_of
            flat, loose = grads_of(self.model)
            allreduce_mean_(flat, loose, world=self.world)
        if self.opt is None:
            return loss
        self.opt.step()
        if hasattr(self.model, "after_optimizer_step") and not getattr(self.opt, "refreshes_shadows", False):
            self.model.after_optimizer_step()
        return loss

    def _assign(self):
        """host: block-diagonal costs -> scipy LSA -> padded index tensors (pinned) -> device."""
        c = self.h_cost.numpy()
        idx = self.h_idx.numpy()
        idx[0].fill(self.Q)
        idx[1].fill(0)
        idx[2].fill(0)
        off = 0
        for b, t in enumerate(self.sizes):
            if t:
                i, j = linear_sum_assignment(c[off:off + self.Q * t].reshape(self.Q, t))
                k = len(i)
                idx[0, b, :k], idx[1, b, :k], idx[2, b, :k] = i, j, 1
            off += self.Q * t
        self.s_pi.copy_(self.h_idx[0], non_blocking=True)
        self.s_ti.copy_(self.h_idx[1], non_blocking=True)
        self.s_valid.copy_(self.h_idx[2], non_blocking=True)

    def _pack_targets(self, labels, boxes, h_tgt_i, h_tgt_f, h_tl, h_tb):
        """Per-image host targets -> the pinned, padded arrays the kernels read (concatenated ids/boxes + offsets for
        the cost kernel and the assignment; [B, t_max] padded labels/boxes for the loss)."""
        B, tm = self.B, self.t_max
        sizes = [int(l.numel()) for l in labels]
        assert max(sizes) <= tm, "raise t_max"
        for l in labels:  # the reference raises (IndexError / one_hot) on a label outside [0, C): so do we, on the host
            if l.numel() and (int(l.min()) < 0 or int(l.max()) >= self.C):
                raise IndexError(f"target label out of range [0, {self.C}): min {int(l.min())}, max {int(l.max())}")
        hi, hf = h_tgt_i.numpy(), h_tgt_f.numpy().reshape(-1, 4)
        tl, tb = h_tl.numpy(), h_tb.numpy()
        tl.fill(1)
        tb.fill(0)
        off = 0
        for b, (l, bx) in enumerate(zip(labels, boxes)):
            t = sizes[b]
            hi[B * tm + b] = off
            if t:
                hi[off:off + t] = l.numpy()
                hf[off:off + t] = bx.numpy()
                tl[b, :t] = l.numpy()
                tb[b, :t] = bx.numpy()
            off += t
        hi[B * tm + B] = off
        return sizes

    def _statics(self):
        return [self.s_feats, self.s_mask, self.s_sel, self.s_centers, self.s_ids, self.s_offs, self.s_tboxes, self.s_tl,
                self.s_tb]

    def load_batch(self, feats, mask, sel, centers, labels: Sequence[torch.Tensor], boxes: Sequence[torch.Tensor]):
        """Copy one batch into the static buffers.  feats/mask/sel/centers may be host (pinned) or device
        tensors; labels/boxes are per-image HOST tensors (int64 [T_i], fp32 [T_i,4] xyxy), T_i <= t_max."""
        B, tm = self.B, self.t_max
        if self._loaded is not None:
            # the previous call's non-blocking copies read the SAME pinned arrays: they must have executed before the
            # host rewrites them (a caller looping load_batch()+step() without reading the loss runs ahead of the GPU)
            self._loaded.synchronize()
        self.sizes = self._pack_targets(labels, boxes, self.h_tgt_i, self.h_tgt_f, self.h_tl, self.h_tb)
        self.d_blk.buf.copy_(self.h_blk.buf, non_blocking=True)  # all five target arrays in one H2D copy
        dsts = self._statics()
        srcs = [feats, mask, sel, centers] + self.d_blk.views()
        on_dev = [k for k, src in enumerate(srcs) if src.is_cuda and src.is_contiguous() and src.dtype == dsts[k].dtype
                  and src.numel() == dsts[k].numel()]
        ops.copy_many([dsts[k] for k in on_dev], [srcs[k] for k in on_dev])  # one launch: targets (+ resident inputs)
        for k, (dst, src) in enumerate(zip(dsts, srcs)):
            if k not in on_dev:  # host (pinned) inputs, or inputs that need a cast / broadcast
                dst.copy_(src, non_blocking=True)
        if self._loaded is None:
            self._loaded = torch.cuda.Event()
        self._loaded.record()

    # ---- input pipelining: stage batch s+1 (host packing + H2D on a copy stream) while step s runs ----
    def prefetch(self, feats, mask, sel, centers, labels: Sequence[torch.Tensor], boxes: Sequence[torch.Tensor]):
        """Stage the NEXT batch: pack its targets into this slot's pinned arrays and enqueue all H2D copies into
        device staging buffers on a copy stream.  Returns immediately; `step_prefetched()` consumes it."""
        if self._slots is None:
            B, tm = self.B, self.t_max
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._slots = []
            for _ in range(2):
                hb, db = _TargetBlock(B, tm), _TargetBlock(B, tm, device=self.dev)
                self._slots.append(dict(dev=[torch.empty_like(t) for t in self._statics()[:4]] + db.views(), hb=hb, db=db,
                                        ready=torch.cuda.Event(), consumed=torch.cuda.Event()))
            self._slot = 0
        self._slot ^= 1
        st = self._slots[self._slot]
        st["ready"].synchronize()  # this slot's previous H2D copies are done: its pinned arrays may be rewritten
        B, tm = self.B, self.t_max
        hb = st["hb"]
        st["sizes"] = self._pack_targets(labels, boxes, hb.ints, hb.flt, hb.tl, hb.tb)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(st["consumed"])  # the step that used this slot has copied it out
            for dst, src in zip(st["dev"][:4], (feats, mask, sel, centers)):
                dst.copy_(src, non_blocking=True)
            st["db"].buf.copy_(hb.buf, non_blocking=True)  # all five target arrays in one H2D copy
            st["ready"].record(self._copy_stream)
        self._pending = self._slot

    def step_prefetched(self):
        """One training step on the batch staged by the last prefetch(): device-to-device hand-over into the graph's
        static inputs, then the graph replay.  Returns the (device) loss tensor without synchronising."""
        st = self._slots[self._pending]
        main = torch.cuda.current_stream()
        main.wait_event(st["ready"])
        ops.copy_many(self._statics(), st["dev"])  # one launch instead of nine memcpy nodes in front of the replay
        st["consumed"].record(main)
        self.sizes = st["sizes"]
        return self.step()

    def eager_step(self):
        out = self._forward()
        if not self.gpu_lsa:
            torch.cuda.current_stream().synchronize()
            self._assign()
        if self.opt is not None:
            self.opt.zero_grad(set_to_none=False)  # keep the (possibly graph-static) .grad tensors in place
        return self._backward(out)

    def capture(self, warmup: int = 3):
        """Warm up eagerly on a side stream (PyTorch's capture protocol), then capture the step."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.eager_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self.opt is not None:
            self.opt.zero_grad(set_to_none=True)  # backward inside the capture allocates .grad from the graph pool
        from . import _lib
        n0 = _lib.launch_count
        self.gA = torch.cuda.CUDAGraph(keep_graph=True)
        # The step's critical chain is captured on a HIGH-priority stream; the side branches (weight-gradient GEMMs,
        # dropout bit matrices, optimizer share) keep the default priority, so where both have CTAs to place the
        # block scheduler serves the chain first.  (Stream priorities become kernel-node priorities in the graph.)
        cap = _capture_stream(self.dev)
        if self.gpu_lsa:  # one graph: forward, cost, assignment, loss, backward, optimizer
            with torch.cuda.graph(self.gA, stream=cap):
                self.out = self._forward()
                self.loss = self._backward(self.out)
            self.gB = None
        else:
            with torch.cuda.graph(self.gA, stream=cap):
                self.out = self._forward()
            torch.cuda.synchronize()
            self._assign()
            self.gB = torch.cuda.CUDAGraph(keep_graph=True)
            with torch.cuda.graph(self.gB, pool=self.gA.pool(), stream=cap):
                self.loss = self._backward(self.out)
        torch.cuda.synchronize()
        self.launches_per_step = _lib.launch_count - n0  # kernels launched by our C-ABI calls during the capture
        self.graph_kernel_nodes = _count_kernel_nodes([g for g in (self.gA, self.gB) if g is not None])
        for g in (self.gA, self.gB):
            if g is not None:
                g.instantiate()

    def step(self):
        """Replay one training step.  Returns the (device) loss tensor."""
        self.gA.replay()
        if self.gB is not None:
            torch.cuda.current_stream().synchronize()
            self._assign()
            self.gB.replay()
        return self.loss

    # ---- result pipelining: read step s's loss / assignment status while step s+1 already runs ----
    def enqueue_readback(self):
        """Enqueue the device-to-host read of the LAST step's loss and assignment status (pinned slot, non-blocking)
        behind it on the current stream and return a token for finish_readback().  A training loop that launches step
        s+1 before finishing the read of step s keeps the GPU busy across the host's launch latency; the price is that
        an invalid cost matrix (scipy would raise on NaN / -inf) surfaces one step late."""
        if getattr(self, "_rb", None) is None:
            self._rb = [dict(loss=torch.empty(1, dtype=torch.float32, pin_memory=True),
                             status=torch.empty(self.B, dtype=torch.int32, pin_memory=True), ev=torch.cuda.Event())
                        for _ in range(4)]
            self._rb_next = 0
        k = self._rb_next
        self._rb_next = (k + 1) % len(self._rb)
        slot = self._rb[k]
        slot["loss"].copy_(self.loss.detach().reshape(1).float(), non_blocking=True)
        slot["status"].copy_(self.s_status, non_blocking=True)
        slot["ev"].record()
        return k

    def finish_readback(self, token: int) -> float:
        """Wait for the read enqueued by enqueue_readback() and return the loss; raises like raise_if_invalid()."""
        slot = self._rb[token]
        slot["ev"].synchronize()
        if self.gpu_lsa and int(slot["status"].max()) != 0:
            raise ValueError("matrix contains invalid numeric entries")
        return float(slot["loss"][0])

    def raise_if_invalid(self):
        """scipy raises on NaN / -inf costs (identical boxes make the CIoU cost NaN); the device
        assignment records it per image instead.  Synchronises."""
        if self.gpu_lsa and int(self.s_status.max()) != 0:
            raise ValueError("matrix contains invalid numeric entries")
