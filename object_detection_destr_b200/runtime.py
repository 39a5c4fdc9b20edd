"""Hand-scheduled executor of the DESTR transformer half: explicit forward AND backward over the
C-ABI kernels and cuBLAS GEMMs, with flat parameter / gradient buffers.

Why: run through torch autograd op by op, a training step is ~2700 launches of which our kernels are
12 % of the GPU time -- the rest is glue (per-parameter bf16 casts, slice/cat copies, zero fills,
bias-gradient reductions, gradient accumulation adds; see profiles/r01_launches_graph_autograd.txt).
Here the schedule is written out by hand instead:
  * fp32 master parameters, their bf16 shadows and their gradients live in three flat buffers with one
    layout in which every fused GEMM weight (in_proj, decoder q|k|v, the per-layer key/value/position
    projections hoisted over all layers) is one contiguous block -> no cat, no cast, no slice kernels;
    one kernel refreshes the shadows after the optimizer step, one converts the bf16 weight gradients.
  * bias gradients come out of our LayerNorm-backward / ReLU-backward kernels as fused column sums;
    residual-gradient adds are folded into kernel epilogues or GEMM accumulations (addmm, beta = 1).
  * plain GEMMs are cuBLAS (torch.mm/addmm on bf16, fp32 accumulate); everything else is a destr_* kernel.
The module-level autograd path (functional.py) computes the same thing and stays as the cross-check.

Reference semantics: src/model/blocks/encoder_block.py:24-44,88-112; src/model/model.py:89-92;
src/model/blocks/decoder_block.py:28-67,157-220,238-260.
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import ops

Tensor = torch.Tensor
BF16 = torch.bfloat16
_ALIGN = 64
import os as _os
_FUSED_FFN1 = _os.environ.get("DESTR_FUSED_FFN1", "1") == "1"
# encoder projections / FFN / dX / dW on the tcgen05 GEMM family (csrc/gemm_tc.cu) instead of cuBLAS; "0" = library path
_TC = _os.environ.get("DESTR_TC_GEMM", "1") == "1"
# the decoder's projections / FFNs / dX products on the same family (M = B*Q = 800 rows)
# "1": only where an epilogue fusion removes a kernel (ps2 * sine, fc1 + dropout, ReLU-backward + bias gradient);
# "2": every decoder GEMM (parity-green, but measured slower than the library for the plain 800-row products:
#      4.69 vs 4.61 ms/step mid-round, 4.09-4.15 vs 3.99 ms at the end of round 2)
_TC_DEC_LEVEL = int(_os.environ.get("DESTR_TC_DEC", "1")) if _TC else 0
_TC_DEC = _TC_DEC_LEVEL >= 1     # fused members
_TC_DEC_ALL = _TC_DEC_LEVEL >= 2  # plain members too
# out-proj / fc2 with dropout + residual + LayerNorm(s) in the GEMM epilogue (needs _TC).  Parity-green, but OFF by
# default: a CTA must own whole 256-wide rows, so at M = 8400 only 66 CTAs run and the 5-pass epilogue costs more than
# the two row-wise kernels it replaces (measured in situ: 4.97 ms/step with it, 4.77 without)
_TC_LN = _TC and _os.environ.get("DESTR_TC_LN", "0") == "1"
_FORK = {k: _os.environ.get("DESTR_FORK_" + k, d) == "1" for k, d in (("ENC_V", "0"), ("DEC_HEAD", "1"), ("DSIN", "1"))}


def _mm_bias(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return torch.addmm(b, x, w.t())


def _drop_scale(thr16: int) -> float:
    return 65536.0 / (65536.0 - thr16)


def _mm_bias_relu(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    # cuBLASLt GEMM with the bias + ReLU epilogue fused
    return torch._addmm_activation(b, x, w.t(), use_gelu=False)


class FlatParams:
    """Flat fp32 master / bf16 shadow / gradient storage for the encoder + decoder parameters.

    Region W (GEMM weights, fused groups contiguous) comes first, region S (biases, LayerNorm affine)
    after it.  `module.parameters()` are re-pointed at views of the master buffer, so state_dict, the
    reference parameter names and any optimizer keep working; the dead `_proj_to_q/k/v` parameters of
    the reference encoder block are left alone (they never receive gradients)."""

    def __init__(self, encoder: nn.Module, decoder: nn.Module, device, heads=()):
        """heads: [(name, parameter)] of the prediction heads -- they join the fp32 part of the buffer (one optimizer
        launch, one gradient exchange for everything); their gradients are WRITTEN by the heads kernel before
        backward() of the runtime starts, so begin_backward() leaves that region alone."""
        e, d = dict(encoder.named_parameters()), dict(decoder.named_parameters())
        Le, Ld = len(encoder._encoder), len(decoder._decoder)
        self.Le, self.Ld = Le, Ld
        W: List[Tuple[str, Tensor]] = []
        S: List[Tuple[str, Tensor]] = []
        # encoder weights: the shared position-scale MLP first, then the layers in order -- [ps | e0] (whose gradients are
        # final last) is then ONE contiguous tail bucket of the data-parallel exchange
        W += [("e.ps0_w", e["_pos_scale.0.weight"]), ("e.ps2_w", e["_pos_scale.2.weight"])]
        for l in range(Le):
            p = f"_encoder.{l}."
            W += [(f"e{l}.in_w", e[p + "self_attn.in_proj_weight"]), (f"e{l}.out_w", e[p + "self_attn.out_proj.weight"]),
                  (f"e{l}.fc1_w", e[p + "fc1.weight"]), (f"e{l}.fc2_w", e[p + "fc2.weight"])]
            S += [(f"e{l}.in_b", e[p + "self_attn.in_proj_bias"]), (f"e{l}.out_b", e[p + "self_attn.out_proj.bias"]),
                  (f"e{l}.fc1_b", e[p + "fc1.bias"]), (f"e{l}.fc2_b", e[p + "fc2.bias"]),
                  (f"e{l}.n1_w", e[p + "norm1.weight"]), (f"e{l}.n1_b", e[p + "norm1.bias"]),
                  (f"e{l}.n2_w", e[p + "norm2.weight"]), (f"e{l}.n2_b", e[p + "norm2.bias"])]
        S += [("e.ps0_b", e["_pos_scale.0.bias"]), ("e.ps2_b", e["_pos_scale.2.bias"]),
              ("e.n_w", e["norm.weight"]), ("e.n_b", e["norm.bias"])]
        for l in range(Ld):
            p = f"_decoder.{l}."
            W += [(f"d{l}.q_w", d[p + "_sa_proj_to_q_obj.weight"]), (f"d{l}.k_w", d[p + "_sa_proj_to_k_obj.weight"]),
                  (f"d{l}.v_w", d[p + "_sa_proj_to_v_obj.weight"]), (f"d{l}.cq_w", d[p + "_ca_proj_to_q_obj.weight"]),
                  (f"d{l}.cqp_w", d[p + "_ca_proj_to_q_pos.weight"])]
            S += [(f"d{l}.n1_w", d[p + "norm1.weight"]), (f"d{l}.n1_b", d[p + "norm1.bias"]),
                  (f"d{l}.n2_w", d[p + "norm2.weight"]), (f"d{l}.n2_b", d[p + "norm2.bias"])]
            for i, br in enumerate(("_cls_branch.", "_reg_branch.")):
                W += [(f"d{l}.b{i}.fc1_w", d[p + br + "fc1.weight"]), (f"d{l}.b{i}.fc2_w", d[p + br + "fc2.weight"])]
                S += [(f"d{l}.b{i}.fc1_b", d[p + br + "fc1.bias"]), (f"d{l}.b{i}.fc2_b", d[p + br + "fc2.bias"]),
                      (f"d{l}.b{i}.n1_w", d[p + br + "norm1.weight"]), (f"d{l}.b{i}.n1_b", d[p + br + "norm1.bias"]),
                      (f"d{l}.b{i}.n2_w", d[p + br + "norm2.weight"]), (f"d{l}.b{i}.n2_b", d[p + br + "norm2.bias"])]
        for l in range(Ld):  # hoisted groups: contiguous over layers
            p = f"_decoder.{l}."
            W += [(f"d{l}.ke_w", d[p + "_ca_proj_to_k_enc.weight"]), (f"d{l}.ve_w", d[p + "_ca_proj_to_v_enc.weight"])]
        for l in range(Ld):
            W += [(f"d{l}.kp_w", d[f"_decoder.{l}._ca_proj_to_k_pos.weight"])]
        for l in range(Ld):
            p = f"_decoder.{l}."
            W += [(f"d{l}.sqp_w", d[p + "_sa_proj_to_q_pos.weight"]), (f"d{l}.skp_w", d[p + "_sa_proj_to_k_pos.weight"])]
        W += [("d.ps0_w", d["_pos_scale.0.weight"]), ("d.ps2_w", d["_pos_scale.2.weight"])]
        S += [("d.ps0_b", d["_pos_scale.0.bias"]), ("d.ps2_b", d["_pos_scale.2.bias"]),
              ("d.n_w", d["norm.weight"]), ("d.n_b", d["norm.bias"])]

        # fused groups must be adjacent and unpadded: weights are multiples of 64 elements already
        self.off: Dict[str, Tuple[int, torch.Size]] = {}
        o = 0
        for name, t in W:
            assert t.numel() % _ALIGN == 0
            self.off[name] = (o, t.shape)
            o += t.numel()
        self.nW = o
        for name, t in S:
            self.off[name] = (o, t.shape)
            o += (t.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.heads_off = o
        H = [(f"h.{name}", t) for name, t in heads]
        for name, t in H:
            self.off[name] = (o, t.shape)
            o += (t.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.n = o
        self.m32 = torch.zeros(o, dtype=torch.float32, device=device)
        self.s16 = torch.zeros(o, dtype=BF16, device=device)
        self.g32 = torch.zeros(o, dtype=torch.float32, device=device)
        self.g16 = torch.zeros(self.nW, dtype=BF16, device=device)
        self.params: List[nn.Parameter] = []
        self._gviews: List[Tensor] = []
        with torch.no_grad():
            for name, t in W + S + H:
                off, shape = self.off[name]
                view = self.m32[off:off + t.numel()].view(shape)
                view.copy_(t.detach().to(device))
                t.data = view
                t.grad = self.g32[off:off + t.numel()].view(shape)
                self.params.append(t)
                self._gviews.append(t.grad)
        self.refresh()
        self._written = set()
        # weight-gradient GEMMs are off the critical path of backward (nothing downstream reads them before the
        # optimizer): they go to a side stream and overlap the dX chain; inside a CUDA-graph capture this becomes a
        # parallel branch of the graph.  Their operands are kept alive until the streams join (end_backward).
        self.side = torch.cuda.Stream(device=device) if torch.device(device).type == "cuda" else None
        # forks[0]: short parallel chains the main chain joins right away (high priority, like the capture stream of
        # engine.capture); forks[1]: background work (dropout bit matrices) at the default priority
        hi = -1 if _os.environ.get("DESTR_PRIO", "1") != "0" else 0
        self.forks = [torch.cuda.Stream(device=device, priority=hi), torch.cuda.Stream(device=device)] \
            if self.side is not None else []
        self._keep: List[Tensor] = []
        self._forked = set()
        self.defer_grad_cast = False  # set by FlatAdamW: its kernel reads the bf16 weight gradients itself
        self.dec_off = self.off["d0.q_w"][0]
        self.enc_w_end = self.off["d0.q_w"][0]  # W region = [encoder weights | decoder weights]
        self.dec_s_off = self.off["d0.n1_w"][0]  # S region = [encoder biases / LN | decoder biases / LN | heads]
        self.early_opt = None
        self.early_stream = None
        # first parameter whose gradient backward() leaves in bf16 (g16); below it the fp32 buffer is already final
        self.bf16_begin = self.enc_w_end if _TC else 0
        self.g16_pending = False
        self.grad_scale = 1.0

    def _v(self, buf: Tensor, name: str, rows: Optional[int] = None) -> Tensor:
        off, shape = self.off[name]
        if rows is None:
            return buf[off:off + shape.numel()].view(shape)
        return buf[off:off + rows * shape[1]].view(rows, shape[1])  # fused group starting at `name`

    def w(self, name, rows=None):      # bf16 shadow (weights and biases)
        return self._v(self.s16, name, rows)

    def f(self, name):                 # fp32 master (LayerNorm affine)
        return self._v(self.m32, name)

    def g(self, name):                 # fp32 gradient (region S: accumulated into by our kernels)
        return self._v(self.g32, name)

    def gw(self, name, rows=None):     # bf16 weight gradient (region W)
        return self._v(self.g16, name, rows)

    def refresh(self):
        """bf16 shadows <- fp32 masters (one kernel; call after every optimizer step)."""
        self.s16.copy_(self.m32)

    def begin_backward(self):
        self.g32[self.nW:self.heads_off].zero_()
        self._written.clear()
        self._dec_reduced = False
        self.early_stream = None
        self._enc_reduced = self.enc_w_end
        if _TC:
            # the split-K dW kernel accumulates the encoder's weight gradients in fp32 (red.global.add) straight into
            # the flat gradient buffer: zero that region, on the stream the dW kernels run on
            if self.side is not None:
                self.side.wait_stream(torch.cuda.current_stream())
            with (torch.cuda.stream(self.side) if self.side is not None else contextlib.nullcontext()):
                self.g32[:self.enc_w_end].zero_()

    # ---- data parallel: the one exchange of a training step, overlapped with backward ----
    def enable_data_parallel(self, world: int, group=None):
        """Mean all-reduce of the gradients over `world` ranks inside backward(): the weight gradients travel as
        bf16 (they are produced in bf16; half the NVLink bytes), in two pieces -- the decoder's as soon as the
        decoder backward is done (it overlaps the encoder backward on a communication stream), the encoder's and
        the fp32 bias / LayerNorm gradients at the end."""
        self.world, self.group = world, group
        self.comm = torch.cuda.Stream(device=self.g32.device) if world > 1 else None

    def reduce_decoder_grads(self):
        """Called when the decoder's backward is done (its gradients, the heads' included, are final).  Data parallel:
        exchange them now, under the encoder backward.  With an optimizer that registered `early_opt` (FlatAdamW
        .enable_overlap(): the caller guarantees that backward() is followed by step()), the decoder's share of the
        optimizer step -- 60 % of the parameters, an HBM-bound 70 us -- also runs now, on the communication / a side
        stream, instead of after the encoder backward."""
        early = getattr(self, "early_opt", None)
        dp = getattr(self, "world", 1) > 1
        if not dp and early is None:
            return
        if dp:
            stream = self.comm
        else:
            if getattr(self, "opt_stream", None) is None:
                self.opt_stream = torch.cuda.Stream(device=self.g32.device)
            stream = self.opt_stream
        stream.wait_stream(self.side)                      # the decoder's weight-gradient GEMMs, the heads' gradients
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            if dp:
                import torch.distributed as dist
                # the per-layer blocks went during the decoder backward: what is left are the hoisted key / value /
                # position projections and the shared position-scale MLP
                a = self.off["d0.ke_w"][0] if getattr(self, "_dec_reduced", False) else self.dec_off
                dist.all_reduce(self.g16[a:], group=self.group)
                if early is not None:  # the decoder's + heads' fp32 block travels now too
                    dist.all_reduce(self.g32[self.dec_s_off:], group=self.group)
            if early is not None:
                early()
        self.early_stream = stream if early is not None else None

    def reduce_decoder_layer(self, l: int):
        """Data parallel: decoder layer l's own weight gradients (q|k|v, cross-attention query projections, both branch
        FFNs: contiguous in the flat layout) are final when its backward is done -> exchange them under the backward of
        layer l-1.  The decoder's kernels are small (<= 128 CTAs), so a co-running collective costs them little; issuing
        the whole 30 MB at the start of the ENCODER backward instead displaced its single-wave 132-148-CTA kernels."""
        if getattr(self, "world", 1) <= 1:
            return
        import torch.distributed as dist
        a = self.off[f"d{l}.q_w"][0]
        b = self.off[f"d{l + 1}.q_w"][0] if l + 1 < self.Ld else self.off["d0.ke_w"][0]
        self.comm.wait_stream(self.side)
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            dist.all_reduce(self.g16[a:b], group=self.group)
        self._dec_reduced = True

    def reduce_encoder_layer(self, l: int):
        """Data parallel: called when layer l's backward (and its dW kernels on the side stream) is done.  The encoder's
        weight gradients travel in a few LARGE buckets issued under the remaining backward -- layers [L/2, L) after
        layer L/2, layers [1, L/2) after layer 1 -- because every collective costs launch latency and SMs that the
        single-wave compute kernels then miss; [position-scale MLP | layer 0], final only at the very end, is one
        contiguous tail bucket in end_backward()."""
        if getattr(self, "world", 1) <= 1:
            return
        starts = sorted({max(1, self.Le // 2), 1} & set(range(1, self.Le)), reverse=True)
        if l not in starts:
            return
        import torch.distributed as dist
        a = self.off[f"e{l}.in_w"][0]
        b = getattr(self, "_enc_reduced", self.enc_w_end)
        self.comm.wait_stream(self.side)
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            if _TC:
                self.g16[a:b].copy_(self.g32[a:b])  # fp32 accumulators -> the bf16 the exchange carries
            dist.all_reduce(self.g16[a:b], group=self.group)
        self._enc_reduced = a

    def end_backward(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
            self._keep.clear()
        if getattr(self, "world", 1) > 1:
            import torch.distributed as dist
            self.comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm):
                # whatever the per-layer buckets did not cover: the shared position-scale MLP (and, without per-layer
                # calls, the whole encoder) + the fp32 bias / LayerNorm / head block
                # the tail bucket: whatever the in-backward buckets did not cover = [position-scale MLP | layer 0 ...]
                a, b = 0, getattr(self, "_enc_reduced", self.enc_w_end)
                if b > a:
                    if _TC:
                        self.g16[a:b].copy_(self.g32[a:b])
                    dist.all_reduce(self.g16[a:b], group=self.group)
                self._enc_reduced = self.enc_w_end
                # (the decoder's / heads' part went with the decoder exchange when the optimizer overlaps its step)
                s_end = self.dec_s_off if getattr(self, "early_stream", None) is not None else self.n
                dist.all_reduce(self.g32[self.nW:s_end], group=self.group)
            torch.cuda.current_stream().wait_stream(self.comm)
        # the weight gradients are still bf16 (g16) and, data-parallel, everything is a SUM over ranks: a FlatAdamW
        # bound to this buffer folds the widening and the 1/world into its own pass (and leaves the fp32 values in
        # .grad); otherwise do it here
        self.grad_scale = 1.0 / getattr(self, "world", 1)
        self.g16_pending = True
        if not self.defer_grad_cast:
            self.materialize_grads()
        for t, gv in zip(self.params, self._gviews):  # zero_grad(set_to_none=True) may have detached them
            if t.grad is not gv:
                t.grad = gv

    def materialize_grads(self):
        """Finish .grad after backward(): widen the bf16 weight gradients into the fp32 buffer and apply the
        data-parallel 1/world.  (No-op when already done; FlatAdamW.step() does the same inside its kernel.)"""
        if not getattr(self, "g16_pending", False):
            return
        b0 = self.g16_begin()
        if self.grad_scale != 1.0:
            torch.mul(self.g16[b0:], self.grad_scale, out=self.g32[b0:self.nW])
            self.g32[:b0].mul_(self.grad_scale)
            self.g32[self.nW:].mul_(self.grad_scale)
        else:
            self.g32[b0:self.nW].copy_(self.g16[b0:])
        self.g16_pending = False

    def g16_begin(self) -> int:
        """First parameter whose final gradient is in the bf16 buffer after backward(): the decoder's (library dW GEMMs);
        data-parallel, the encoder's too (they travel as bf16)."""
        return 0 if getattr(self, "world", 1) > 1 else self.bf16_begin

    @contextlib.contextmanager
    def fork(self, enable: bool = True, k: int = 0):
        """`with P.fork():` runs the body on side stream k, concurrently with what the caller enqueues next on its
        own stream; the caller joins with P.join(k).  (Graph capture turns this into a parallel branch.)"""
        if not self.forks or not enable:
            yield
            return
        self.forks[k].wait_stream(torch.cuda.current_stream())
        self._forked.add(k)
        with torch.cuda.stream(self.forks[k]):
            yield

    def join(self, k: int = 0):
        if k in self._forked:  # (joining a stream that never forked would pull un-captured work into a capture)
            self._forked.discard(k)
            torch.cuda.current_stream().wait_stream(self.forks[k])

    def off_path(self, fn, *operands: Tensor):
        """Run `fn()` (work nothing downstream waits for, e.g. a bias-gradient column sum) on the side stream."""
        if self.side is None:
            return fn()
        self.side.wait_stream(torch.cuda.current_stream())
        self._keep += list(operands)
        with torch.cuda.stream(self.side):
            return fn()

    def acc_gw(self, name: str, dy: Tensor, x: Tensor, rows: Optional[int] = None, sl: Optional[slice] = None):
        """weight gradient dW (+)= dy^T x into the bf16 gradient buffer (first write overwrites)."""
        if _TC and self.off[name][0] < self.enc_w_end:  # encoder weight: tcgen05 split-K kernel, fp32 accumulation
            gv = self._v(self.g32, name, rows)
            if sl is not None:
                gv = gv[sl]
            return self.off_path(lambda: ops.gemm_dw(dy, x, gv), dy, x)
        gv = self.gw(name, rows)
        key = name
        if sl is not None:
            gv = gv[sl]
            key = (name, sl.start)
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            self._keep += [dy, x]
            ctx = torch.cuda.stream(self.side)
        else:
            ctx = contextlib.nullcontext()
        with ctx:
            if key in self._written:
                gv.addmm_(dy.t(), x)
            else:
                torch.mm(dy.t(), x, out=gv)
                self._written.add(key)


class FlatAdamW:
    """torch.optim.AdamW semantics on the runtime's flat parameter buffer: one kernel per step updates every
    parameter in it (encoder, decoder and -- with TransformerHalf -- the prediction heads) and refreshes the bf16
    shadows (csrc/optim.cu); the kernel reads the bf16 weight gradients backward() left and folds the 1/world of the
    data-parallel mean.  `extra` parameters living outside the flat buffer, if any, go through a regular fused torch
    AdamW with the same hyper-parameters.  CUDA-graph capturable (the step count lives on the device)."""

    refreshes_shadows = True  # engine: no separate master -> bf16 shadow cast after step()

    def __init__(self, P: "FlatParams", extra=(), lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        self.P, self.lr, self.betas, self.eps, self.wd = P, lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(P.m32)
        self.exp_avg_sq = torch.zeros_like(P.m32)
        self.t = torch.zeros(1, dtype=torch.float32, device=P.m32.device)
        P.defer_grad_cast = True  # step() consumes the bf16 weight gradients directly (P.materialize_grads() to peek)
        extra = [p for p in extra if p.requires_grad]
        self.extra = torch.optim.AdamW(extra, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, fused=True,
                                       capturable=True) if extra else None

    def zero_grad(self, set_to_none: bool = False):
        if self.extra is not None:
            self.extra.zero_grad(set_to_none=set_to_none)

    def enable_overlap(self, on: bool = True):
        """The caller promises that every runtime backward() is followed by exactly one step() (a training loop, the
        engine's step graph): the decoder's + heads' share of the update then runs as soon as the decoder backward is
        done, concurrently with the encoder backward (FlatParams.reduce_decoder_grads)."""
        self.P.early_opt = self._early if on else None

    def _launch(self, lo: int, hi: int, bf16_lo: int, bf16_hi: int, scale: float):
        """AdamW on parameters [lo, hi); gradients of [bf16_lo, bf16_hi) (absolute indices, inside [lo, hi)) are bf16."""
        P = self.P
        from . import _lib
        n = hi - lo
        assert n % 4 == 0 and lo % 4 == 0
        has16 = bf16_hi > bf16_lo
        _lib.call("destr_flat_adamw", P.m32.data_ptr() + 4 * lo, P.g32.data_ptr() + 4 * lo, self.exp_avg.data_ptr() + 4 * lo,
                  self.exp_avg_sq.data_ptr() + 4 * lo, P.s16.data_ptr() + 2 * lo, n, float(self.lr), float(self.betas[0]),
                  float(self.betas[1]), float(self.eps), float(self.wd), self.t.data_ptr(),
                  (P.g16.data_ptr() + 2 * lo) if has16 else None, (bf16_lo - lo) if has16 else 0,
                  (bf16_hi - lo) if has16 else 0, float(scale), ops._stream())

    def _early(self):
        """Decoder weights [dec_off, nW) (bf16 gradients) and decoder biases / LayerNorm / heads [dec_s_off, n)."""
        P = self.P
        self.t.add_(1.0)
        scale = 1.0 / getattr(P, "world", 1)
        self._launch(P.dec_off, P.nW, P.dec_off, P.nW, scale)
        n_end = (P.n + 3) // 4 * 4
        self._launch(P.dec_s_off, n_end, 0, 0, scale)
        self._early_done = True

    def step(self):
        P = self.P
        n = (P.n + 3) // 4 * 4
        pend = P.g16_pending  # backward left bf16 weight gradients (else .grad was filled by the caller)
        if getattr(self, "_early_done", False):
            # the decoder's share ran under the encoder backward: finish with the encoder weights and biases
            self._early_done = False
            if P.early_stream is not None:
                torch.cuda.current_stream().wait_stream(P.early_stream)
            b0 = P.g16_begin() if pend else P.dec_off
            self._launch(0, P.dec_off, min(b0, P.dec_off), P.dec_off, float(P.grad_scale) if pend else 1.0)
            self._launch(P.nW, P.dec_s_off, 0, 0, float(P.grad_scale) if pend else 1.0)
        else:
            self.t.add_(1.0)
            self._launch(0, n, P.g16_begin() if pend else 0, P.nW if pend else 0, float(P.grad_scale) if pend else 1.0)
        P.g16_pending = False
        if self.extra is not None:
            self.extra.step()


# ids of the dropout calls ("sites") of the reference, one per call and layer; the mask of a site is a pure function
# of (seed, site, row, column), so forward and backward agree without storing it (csrc/common.cuh)
ENC_SITES = {"attn": 0, "d1": 1, "d2": 2, "d3": 3}                      # encoder_block.py:57-69, 97-109
DEC_SITES = {"sa": 0, "d1a": 1, "d1b": 2, "ca": 3,                       # decoder_block.py:86,132,182-184,230
             "b0.d_ca": 4, "b0.d_relu": 5, "b0.d_fc2": 6, "b1.d_ca": 7, "b1.d_relu": 8, "b1.d_fc2": 9}  # :234,253-256


def enc_site(layer: int, name: str) -> int:
    return 16 * layer + ENC_SITES[name]


def dec_site(layer: int, name: str) -> int:
    return 4096 + 32 * layer + DEC_SITES[name]


class HotPathRuntime:
    """forward()/backward() of encoder -> fine_pos -> decoder on token-major bf16 activations."""

    def __init__(self, encoder: nn.Module, decoder: nn.Module, bbox_embed: nn.Module, device, cls_embed=None):
        heads = []
        if cls_embed is not None:  # the prediction heads share the flat buffer (fp32 region)
            heads = [("cls_w", cls_embed.weight), ("cls_b", cls_embed.bias), ("box0_w", bbox_embed[0].weight),
                     ("box0_b", bbox_embed[0].bias), ("box2_w", bbox_embed[2].weight), ("box2_b", bbox_embed[2].bias)]
        self.P = FlatParams(encoder, decoder, device, heads=heads)
        self.heads_in_flat = bool(heads)
        self.bbox = bbox_embed
        self.enc_mod, self.dec_mod = encoder, decoder
        from .encoder import initial_dropout_seed
        # dropout seed counter (on the device: graph replays see the updates); starts from torch's seed and the rank
        self.seed = torch.full((1,), initial_dropout_seed(), dtype=torch.int32, device=device)
        self.Le, self.Ld = self.P.Le, self.P.Ld
        self.saved = None
        self._bbox_cache = None
        self.anchor = torch.zeros(1, device=device, requires_grad=True)

    # ------------------------------------------------------------------ dropout
    def _drop_config(self):
        """Dropout probabilities in force for this forward, as 16-bit thresholds.  nn.Dropout and
        nn.MultiheadAttention(dropout=) act in training mode only; the reference's SelfAttention builds its
        nn.Dropout inline (self_attention.py:40), so the decoder's attention-probability dropout is ALWAYS on."""
        t = ops.drop_thr16
        eb, db = self.enc_mod._encoder[0], self.dec_mod._decoder[0]
        etr, dtr = self.enc_mod.training, self.dec_mod.training
        cfg = {
            "e.attn": t(eb.self_attn.dropout) if etr else 0,
            "e.d1": t(eb.dropout1.p) if etr else 0, "e.d2": t(eb.dropout2.p) if etr else 0,
            "e.d3": t(eb.dropout3.p) if etr else 0,
            "d.sa": t(db._self_attn._dropout_prob),
            "d.d1": t(db.dropout1.p) if dtr else 0,
            "d.ca": t(db._cls_branch.cross_attn._dropout_prob),
            "d.br": t(db._cls_branch.dropout.p) if dtr else 0,
        }
        return cfg

    def _d(self, thr16: int, site: int, site2: Optional[int] = None):
        """`drop` argument of an op: None when the site is inactive."""
        if not thr16:
            return None
        return (self.seed, thr16, site) if site2 is None else (self.seed, thr16, site, site2)

    def _gen_attn_bits(self, B: int, N: int, device):
        """Bit matrices of the encoder attention dropout mask of every layer (ops.attn_dropout_bits), written by a
        background stream while the first layers' GEMMs run: [(rowbits_l, colbits_l, ready event)]."""
        thr = self.dc["e.attn"]
        if not thr:
            return None
        P, L = self.P, self.Le
        words, Np = ops.mask_words(N), (N + 127) // 128 * 128
        rb = torch.empty(L, B * 8, words, Np, dtype=torch.int32, device=device)  # allocated on the caller's stream
        cb = torch.empty_like(rb)
        bg = P.forks[1] if P.forks else None
        out = []
        if bg is not None:
            bg.wait_stream(torch.cuda.current_stream())
        with (torch.cuda.stream(bg) if bg is not None else contextlib.nullcontext()):
            for l in range(L):
                ops.attn_dropout_bits((self.seed, thr, enc_site(l, "attn")), B * 8, N, device, out=(rb[l], cb[l]))
                ev = None
                if bg is not None:
                    ev = torch.cuda.Event()
                    ev.record(bg)
                out.append((rb[l], cb[l], ev))
        return out

    def _bbox_bf16(self):
        """bf16 copy of the box head's first layer, refreshed once per forward()."""
        if self._bbox_cache is None:
            self._bbox_cache = (self.bbox[0].weight.detach().to(BF16), self.bbox[0].bias.detach().to(BF16))
        return self._bbox_cache

    # ------------------------------------------------------------------ encoder
    def _enc_fwd(self, l: int, x: Tensor, pos: Tensor, bits: Tensor, B: int, N: int):
        P = self.P
        Win, b_in = P.w(f"e{l}.in_w"), P.w(f"e{l}.in_b")
        dc = self.dc
        if _TC:  # tcgen05 GEMM family: biases are read in fp32, the position-scale add lives in a GEMM epilogue
            bf = P.f(f"e{l}.in_b")
            with P.fork(_FORK["ENC_V"], k=0):
                v = ops.gemm(x, Win[512:], bias=bf[512:])
            h1 = ops.gemm(x, P.w("e.ps0_w"), bias=P.f("e.ps0_b"), relu=True)
            xq = ops.gemm(h1, P.w("e.ps2_w"), bias=P.f("e.ps2_b"), mul=pos, add=x)     # x + pos * pos_scale(x)
            qk = ops.gemm(xq, Win[:512], bias=bf[:512])
        else:
            with P.fork(_FORK["ENC_V"], k=0):  # the value projection does not depend on the position-scale chain
                v = _mm_bias(x, Win[512:], b_in[512:])
            h1 = _mm_bias_relu(x, P.w("e.ps0_w"), P.w("e.ps0_b"))
            s = _mm_bias(h1, P.w("e.ps2_w"), P.w("e.ps2_b"))
            xq = ops.pos_mul_add(x, pos, s)
            qk = _mm_bias(xq, Win[:512], b_in[:512])
        P.join(0)
        da_ = self._d(dc["e.attn"], enc_site(l, "attn"))
        # attention dropout mask as bit matrices: rows for the forward kernel, columns kept for the backward
        rb = cb = None
        if da_ is not None:
            rb, cb, ev = self._attn_bits[l]
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
        a, lse = ops.enc_attn_fwd(qk[:, :256], qk[:, 256:], v, bits, B, N, 8, 1.0 / math.sqrt(32), drop=da_, rowbits=rb)
        d1_, d2_, d3_ = (self._d(dc["e.d" + k], enc_site(l, "d" + k)) for k in "123")
        if _TC_LN:
            # out-proj + dropout1 + residual + norm1 in one kernel; z1 = x + dropout1(out_proj(a)) is what norm1's
            # backward needs (it replaces the saved out-proj output)
            x1, o, m1, r1 = ops.gemm_res_ln(a, P.w(f"e{l}.out_w"), P.f(f"e{l}.out_b"), x, P.f(f"e{l}.n1_w"),
                                            P.f(f"e{l}.n1_b"), drop=d1_)
        else:
            o = ops.gemm(a, P.w(f"e{l}.out_w"), bias=P.f(f"e{l}.out_b")) if _TC else \
                _mm_bias(a, P.w(f"e{l}.out_w"), P.w(f"e{l}.out_b"))
            x1, m1, r1 = ops.add_layernorm(x, o, P.f(f"e{l}.n1_w"), P.f(f"e{l}.n1_b"), save_stats=True, drop=d1_)
        if _FUSED_FFN1:  # tcgen05 GEMM with bias + ReLU + dropout2 in its epilogue (csrc/gemm_bias_relu.cu)
            f1 = ops.linear_bias_relu_dropout(x1, P.w(f"e{l}.fc1_w"), P.f(f"e{l}.fc1_b"), d2_)
        else:    # library path: cuBLASLt bias+ReLU epilogue, then the dropout pass
            f1 = _mm_bias_relu(x1, P.w(f"e{l}.fc1_w"), P.w(f"e{l}.fc1_b"))
            ops.dropout_inplace(f1, d2_)
        if _TC_LN:
            # fc2 + dropout3 + residual + norm2, then the encoder's shared norm on x + block(x), all in the epilogue
            x2, g, m2, r2, xo, m3, r3 = ops.gemm_res_ln(f1, P.w(f"e{l}.fc2_w"), P.f(f"e{l}.fc2_b"), x1, P.f(f"e{l}.n2_w"),
                                                        P.f(f"e{l}.n2_b"), drop=d3_, res2=x, gamma2=P.f("e.n_w"),
                                                        beta2=P.f("e.n_b"))
        else:
            g = ops.gemm(f1, P.w(f"e{l}.fc2_w"), bias=P.f(f"e{l}.fc2_b")) if _TC else \
                _mm_bias(f1, P.w(f"e{l}.fc2_w"), P.w(f"e{l}.fc2_b"))
            # norm2(x1 + dropout3(fc2 ..)) and the encoder's shared norm(x + block(x)) in one row pass
            x2, m2, r2, xo, m3, r3 = ops.add_layernorm2(x1, g, P.f(f"e{l}.n2_w"), P.f(f"e{l}.n2_b"), x, P.f("e.n_w"),
                                                        P.f("e.n_b"), drop=d3_)
        return xo, (x, h1, xq, qk, v, a, lse, o, x1, m1, r1, f1, g, x2, m2, r2, m3, r3, cb)

    def _enc_bwd(self, l: int, dxo: Tensor, sv, pos: Tensor, bits: Tensor, B: int, N: int) -> Tensor:
        P = self.P
        x, h1, xq, qk, v, a, lse, o, x1, m1, r1, f1, g, x2, m2, r2, m3, r3, cb = sv
        dc = self.dc
        if _TC_LN:
            # xo = LN(x + x2)
            d3, _, _ = ops.add_layernorm_bwd(dxo, x, x2, P.f("e.n_w"), m3, r3, dgamma=P.g("e.n_w"), dbeta=P.g("e.n_b"))
            # x2 = LN(x1 + fc2(relu(fc1 x1))): `g` holds the pre-LayerNorm sum z (fused-LayerNorm forward), so the backward
            # reads one operand (a = z, b = None); the dropout mask is still applied to the branch gradient
            r_ = ops.add_layernorm_bwd(d3, g, None, P.f(f"e{l}.n2_w"), m2, r2, dgamma=P.g(f"e{l}.n2_w"),
                                       dbeta=P.g(f"e{l}.n2_b"), dbias=P.g(f"e{l}.fc2_b"),
                                       drop=self._d(dc["e.d3"], enc_site(l, "d3")), want_sum=bool(dc["e.d3"]))
            d2, d2s = r_[0], (r_[3] if dc["e.d3"] else r_[0])
        else:
            # xo = LN(x + x2), x2 = LN(x1 + dropout3(fc2 ..)): both LayerNorm backwards in one row pass.  With dropout3
            # active the gradient splits: d2 goes through the mask into fc2, d2s is the un-masked residual gradient
            d3, d2, d2s = ops.add_layernorm2_bwd(dxo, x, x2, P.f("e.n_w"), m3, r3, x1, g, P.f(f"e{l}.n2_w"), m2, r2,
                                                 P.g("e.n_w"), P.g("e.n_b"), P.g(f"e{l}.n2_w"), P.g(f"e{l}.n2_b"),
                                                 dbias=P.g(f"e{l}.fc2_b"), drop=self._d(dc["e.d3"], enc_site(l, "d3")),
                                                 want_sum=bool(dc["e.d3"]))
            if d2s is None:
                d2s = d2
        P.acc_gw(f"e{l}.fc2_w", d2, f1)
        if _TC:
            # dX of fc2 + the ReLU/dropout2 mask + fc1's bias gradient in one kernel, then dX of fc1 + residual gradient
            dpre = ops.gemm_relu_bwd(d2, P.w(f"e{l}.fc2_w"), f1, _drop_scale(dc["e.d2"]), P.g(f"e{l}.fc1_b"))
            P.acc_gw(f"e{l}.fc1_w", dpre, x1)
            dx1 = ops.gemm(dpre, P.w(f"e{l}.fc1_w"), b_kn=True, add=d2s)
        else:
            df1 = torch.mm(d2, P.w(f"e{l}.fc2_w"))
            dpre = ops.relu_bwd_colsum(df1, f1, P.g(f"e{l}.fc1_b"), scale=_drop_scale(dc["e.d2"]))
            P.acc_gw(f"e{l}.fc1_w", dpre, x1)
            dx1 = torch.addmm(d2s, dpre, P.w(f"e{l}.fc1_w"))
        # x1 = LN(x + dropout1(out_proj(attn)));  dx = d3 + d(x1 input)
        d1, _, _, dx = ops.add_layernorm_bwd(dx1, o if _TC_LN else x, None if _TC_LN else o, P.f(f"e{l}.n1_w"), m1, r1,
                                             dgamma=P.g(f"e{l}.n1_w"), dbeta=P.g(f"e{l}.n1_b"), dbias=P.g(f"e{l}.out_b"),
                                             res_in=d3, drop=self._d(dc["e.d1"], enc_site(l, "d1")))
        P.acc_gw(f"e{l}.out_w", d1, a)
        da = ops.gemm(d1, P.w(f"e{l}.out_w"), b_kn=True) if _TC else torch.mm(d1, P.w(f"e{l}.out_w"))
        dqk, dv = ops.enc_attn_bwd(qk[:, :256], qk[:, 256:], v, bits, a, da, lse, B, N, 8, 1.0 / math.sqrt(32),
                                   drop=self._d(dc["e.attn"], enc_site(l, "attn")), colbits=cb)
        gb = P.g(f"e{l}.in_b")
        P.off_path(lambda: (ops.relu_bwd_colsum(dqk, None, gb[:512]), ops.relu_bwd_colsum(dv, None, gb[512:])), dqk, dv)
        Win = P.w(f"e{l}.in_w")
        P.acc_gw(f"e{l}.in_w", dqk, xq, sl=slice(0, 512))
        P.acc_gw(f"e{l}.in_w", dv, x, sl=slice(512, 768))
        if _TC:
            ops.gemm(dv, Win[512:], b_kn=True, add=dx, out=dx)                         # dx += dv Wv
            ds, dx = ops.gemm(dqk, Win[:512], b_kn=True, mul=pos, add2=dx, out2=True)  # ds = dxq * pos, dx += dxq
        else:
            dxq = torch.mm(dqk, Win[:512])
            dx.addmm_(dv, Win[512:])
            ds, dx = ops.pos_mul_add_bwd_acc(dxq, pos, dx)
        dx = self._pos_scale_bwd("e", ds, h1, x, dx)
        P.reduce_encoder_layer(l)  # data parallel: this layer's weight gradients are final
        return dx

    def _pos_scale_bwd(self, pfx: str, ds: Tensor, h1: Tensor, xin: Tensor, dx_acc: Tensor) -> Tensor:
        """backward of s = W2 relu(W0 xin + b0) + b2 (shared MLP): accumulates weight grads, dx_acc += ..."""
        P = self.P
        P.off_path(lambda: ops.relu_bwd_colsum(ds, None, P.g(pfx + ".ps2_b")), ds)
        P.acc_gw(pfx + ".ps2_w", ds, h1)
        if (_TC and pfx == "e") or (_TC_DEC and pfx == "d"):
            dpre = ops.gemm_relu_bwd(ds, P.w(pfx + ".ps2_w"), h1, 1.0, P.g(pfx + ".ps0_b"))
            P.acc_gw(pfx + ".ps0_w", dpre, xin)
            if pfx == "e" or _TC_DEC_ALL:
                ops.gemm(dpre, P.w(pfx + ".ps0_w"), b_kn=True, add=dx_acc, out=dx_acc)
            else:
                dx_acc.addmm_(dpre, P.w(pfx + ".ps0_w"))
            return dx_acc
        dh = torch.mm(ds, P.w(pfx + ".ps2_w"))
        dpre = ops.relu_bwd_colsum(dh, h1, P.g(pfx + ".ps0_b"))
        P.acc_gw(pfx + ".ps0_w", dpre, xin)
        dx_acc.addmm_(dpre, P.w(pfx + ".ps0_w"))
        return dx_acc

    # ------------------------------------------------------------------ decoder
    def _dec_fwd(self, l: int, x: Tensor, kv_all, kpos_all, qkpos_all, sine, centers, bits, B, Q, N, lam, pairs_ov):
        P = self.P
        xr = x[:, 256:]
        # three independent chains start from the layer input; only the box -> pairing chain is on the critical path
        with P.fork(_FORK["DEC_HEAD"], k=0):  # query-position chain: pos_scale MLP -> sin -> its cross-attention projection
            if _TC_DEC:
                t1 = ops.gemm(xr, P.w("d.ps0_w"), bias=P.f("d.ps0_b"), relu=True) if _TC_DEC_ALL else \
                    _mm_bias_relu(xr, P.w("d.ps0_w"), P.w("d.ps0_b"))
                sin = ops.gemm(t1, P.w("d.ps2_w"), bias=P.f("d.ps2_b"), mul=sine)   # pos_scale(x_reg) * sine embedding
                qp = ops.gemm(sin, P.w(f"d{l}.cqp_w")) if _TC_DEC_ALL else torch.mm(sin, P.w(f"d{l}.cqp_w").t())
            else:
                t1 = _mm_bias_relu(xr, P.w("d.ps0_w"), P.w("d.ps0_b"))
                t2 = _mm_bias(t1, P.w("d.ps2_w"), P.w("d.ps2_b"))
                sin = ops.mul(t2, sine)
                qp = torch.mm(sin, P.w(f"d{l}.cqp_w").t())
        with P.fork(_FORK["DEC_HEAD"], k=1):  # packed q|k|v object projection
            qkv_obj = ops.gemm(x, P.w(f"d{l}.q_w", rows=1536)) if _TC_DEC_ALL else torch.mm(x, P.w(f"d{l}.q_w", rows=1536).t())
        bp = self.bbox
        # box refinement of the layer's queries: feeds ONLY the pairing (arg-max indices) and its input is already bf16,
        # so the 256x256 layer runs as a bf16 tensor-core GEMM (bias + ReLU fused) instead of an fp32 SIMT one; the
        # 256 -> 4 output layer stays fp32
        w0, b0 = self._bbox_bf16()
        hbox = ops.gemm(xr, w0, bias=bp[0].bias.detach(), relu=True) if _TC_DEC_ALL else _mm_bias_relu(xr, w0, b0)
        coords = ops.box_head_refine(hbox, bp[2].weight.detach(), bp[2].bias.detach(), centers)  # 256 -> 4 tail + refinement
        pairs = ops.pair_indices(coords.view(B, Q, 4)) if pairs_ov is None else pairs_ov
        P.join(1)
        qkv, cat = ops.dec_qkv_prep(qkv_obj, qkpos_all[:, l * 512:(l + 1) * 512], pairs, B, Q)
        dc = self.dc
        o1, o2, lse1, lse2 = ops.dec_self_pair_attn_fwd(qkv, cat, B, Q, drop=self._d(dc["d.sa"], dec_site(l, "sa")))
        o, st = ops.dual_ln_mix(x, o1, o2, pairs, P.f(f"d{l}.n1_w"), P.f(f"d{l}.n1_b"), P.f(f"d{l}.n2_w"),
                                P.f(f"d{l}.n2_b"), lam, Q,
                                drop=self._d(dc["d.d1"], dec_site(l, "d1a"), dec_site(l, "d1b")))
        qo = ops.gemm(o, P.w(f"d{l}.cq_w")) if _TC_DEC_ALL else torch.mm(o, P.w(f"d{l}.cq_w").t())
        P.join(0)
        ke, vv = kv_all[:, l * 512:l * 512 + 256], kv_all[:, l * 512 + 256:(l + 1) * 512]
        kp = kpos_all[:, l * 256:(l + 1) * 256]
        ca, lse_c = ops.split_cross_attn_fwd(qo, qp, ke, kp, vv, bits, B, Q, N, drop=self._d(dc["d.ca"], dec_site(l, "ca")))
        y = torch.empty_like(x)
        br_saved = []
        for i in (1, 0):  # the class / box branches are independent: branch 1 is forked to the second stream
            sl = slice(i * 256, (i + 1) * 256)
            with P.fork(i == 1):
                xb, mb1, rb1 = ops.add_layernorm(o[:, sl], ca[:, sl], P.f(f"d{l}.b{i}.n1_w"), P.f(f"d{l}.b{i}.n1_b"),
                                                 save_stats=True, drop=self._d(dc["d.br"], dec_site(l, f"b{i}.d_ca")))
                if _TC_DEC:  # fc1 + bias + ReLU + dropout in one tcgen05 GEMM, fc2 on the family too
                    f = ops.linear_bias_relu_dropout(xb, P.w(f"d{l}.b{i}.fc1_w"), P.f(f"d{l}.b{i}.fc1_b"),
                                                     self._d(dc["d.br"], dec_site(l, f"b{i}.d_relu")))
                    g = ops.gemm(f, P.w(f"d{l}.b{i}.fc2_w"), bias=P.f(f"d{l}.b{i}.fc2_b")) if _TC_DEC_ALL else \
                        _mm_bias(f, P.w(f"d{l}.b{i}.fc2_w"), P.w(f"d{l}.b{i}.fc2_b"))
                else:
                    f = _mm_bias_relu(xb, P.w(f"d{l}.b{i}.fc1_w"), P.w(f"d{l}.b{i}.fc1_b"))
                    ops.dropout_inplace(f, self._d(dc["d.br"], dec_site(l, f"b{i}.d_relu")))
                    g = _mm_bias(f, P.w(f"d{l}.b{i}.fc2_w"), P.w(f"d{l}.b{i}.fc2_b"))
                _, mb2, rb2 = ops.add_layernorm(xb, g, P.f(f"d{l}.b{i}.n2_w"), P.f(f"d{l}.b{i}.n2_b"), save_stats=True,
                                                out=y[:, sl], drop=self._d(dc["d.br"], dec_site(l, f"b{i}.d_fc2")))
            br_saved.append((xb, mb1, rb1, f, g, mb2, rb2))
        br_saved.reverse()
        P.join()
        xo, mn, rn = ops.add_layernorm(x, y, P.f("d.n_w"), P.f("d.n_b"), save_stats=True)
        return xo, (x, t1, sin, pairs, qkv, cat, o1, o2, lse1, lse2, o, st, qo, qp, ca, lse_c, y, br_saved, mn, rn), \
            (coords, pairs)

    def _dec_bwd(self, l: int, dxo: Tensor, sv, ctx, d_kv_all, d_kpos_all, d_qkpos_all) -> Tensor:
        P = self.P
        kv_all, kpos_all, sine, bits, kpm, B, Q, N, lam = ctx
        x, t1, sin, pairs, qkv, cat, o1, o2, lse1, lse2, o, st, qo, qp, ca, lse_c, y, br_saved, mn, rn = sv
        d, _, _ = ops.add_layernorm_bwd(dxo, x, y, P.f("d.n_w"), mn, rn, dgamma=P.g("d.n_w"), dbeta=P.g("d.n_b"))
        dc = self.dc
        dca = torch.empty_like(x)                               # d(cross-attention output): through the branch dropout
        dres = torch.empty_like(x) if dc["d.br"] else dca       # d(o) along the residual: not masked
        for i in (1, 0):  # independent branches: branch 1 on the second stream
            sl = slice(i * 256, (i + 1) * 256)
            xb, mb1, rb1, f, g, mb2, rb2 = br_saved[i]
            pf = f"d{l}.b{i}."
            with P.fork(i == 1):
                r_ = ops.add_layernorm_bwd(d[:, sl], xb, g, P.f(pf + "n2_w"), mb2, rb2, dgamma=P.g(pf + "n2_w"),
                                           dbeta=P.g(pf + "n2_b"), dbias=P.g(pf + "fc2_b"),
                                           drop=self._d(dc["d.br"], dec_site(l, f"b{i}.d_fc2")), want_sum=bool(dc["d.br"]))
                d2, d2s = r_[0], (r_[3] if dc["d.br"] else r_[0])
                P.acc_gw(pf + "fc2_w", d2, f)
                if _TC_DEC:
                    df = None
                    dpre = ops.gemm_relu_bwd(d2, P.w(pf + "fc2_w"), f, _drop_scale(dc["d.br"]), P.g(pf + "fc1_b"))
                    P.acc_gw(pf + "fc1_w", dpre, xb)
                    dxb = ops.gemm(dpre, P.w(pf + "fc1_w"), b_kn=True, add=d2s) if _TC_DEC_ALL else \
                        torch.addmm(d2s, dpre, P.w(pf + "fc1_w"))
                else:
                    df = torch.mm(d2, P.w(pf + "fc2_w"))
                    dpre = ops.relu_bwd_colsum(df, f, P.g(pf + "fc1_b"), scale=_drop_scale(dc["d.br"]))
                    P.acc_gw(pf + "fc1_w", dpre, xb)
                    dxb = torch.addmm(d2s, dpre, P.w(pf + "fc1_w"))
                ops.add_layernorm_bwd(dxb, o[:, sl], ca[:, sl], P.f(pf + "n1_w"), mb1, rb1, dgamma=P.g(pf + "n1_w"),
                                      dbeta=P.g(pf + "n1_b"), dx_out=dca[:, sl],
                                      drop=self._d(dc["d.br"], dec_site(l, f"b{i}.d_ca")), want_sum=bool(dc["d.br"]),
                                      res_out=dres[:, sl] if dc["d.br"] else None)
                P._keep += [d2, d2s, df, dpre, dxb]
        P.join()
        ke, vv = kv_all[:, l * 512:l * 512 + 256], kv_all[:, l * 512 + 256:(l + 1) * 512]
        kp = kpos_all[:, l * 256:(l + 1) * 256]
        dqo, dqp, _, _, _ = ops.split_cross_attn_bwd(
            qo, qp, ke, kp, vv, bits, ca, dca, lse_c, B, Q, N, dke_out=d_kv_all[:, l * 512:l * 512 + 256],
            dkp_out=d_kpos_all[:, l * 256:(l + 1) * 256], dv_out=d_kv_all[:, l * 512 + 256:(l + 1) * 512],
            drop=self._d(dc["d.ca"], dec_site(l, "ca")))
        P.acc_gw(f"d{l}.cq_w", dqo, o)
        # d(o) = d(o_cls|o_reg residual) + dq_obj W
        do = ops.gemm(dqo, P.w(f"d{l}.cq_w"), b_kn=True, add=dres) if _TC_DEC_ALL else torch.addmm(dres, dqo, P.w(f"d{l}.cq_w"))
        P.acc_gw(f"d{l}.cqp_w", dqp, sin)
        with P.fork(_FORK["DSIN"], k=0):  # sin = sine * pos_scale(x_reg): independent of the self/pair-attention chain below
            if _TC_DEC:
                dsin = None
                dt2 = ops.gemm(dqp, P.w(f"d{l}.cqp_w"), b_kn=True, mul=sine)
            else:
                dsin = torch.mm(dqp, P.w(f"d{l}.cqp_w"))
                dt2 = ops.mul(dsin, sine)
            xr = x[:, 256:]
            dxr = torch.zeros(x.shape[0], 256, dtype=BF16, device=x.device)
            self._pos_scale_bwd("d", dt2, t1, xr, dxr)
        dx2, do1, do2, delta1, delta2 = ops.dual_ln_mix_bwd(
            do, x, o1, o2, pairs, P.f(f"d{l}.n1_w"), P.f(f"d{l}.n2_w"), st, lam, Q,
            pg=(P.g(f"d{l}.n1_w"), P.g(f"d{l}.n1_b"), P.g(f"d{l}.n2_w"), P.g(f"d{l}.n2_b")), head_major=True,
            drop=self._d(dc["d.d1"], dec_site(l, "d1a"), dec_site(l, "d1b")))
        dx = d + dx2
        d_qkv, d_cat = ops.dec_self_pair_attn_bwd(qkv, cat, do1, do2, lse1, lse2, delta1, delta2, B, Q,
                                                  drop=self._d(dc["d.sa"], dec_site(l, "sa")))
        d_qkv_obj, _ = ops.dec_qkv_prep_bwd(d_qkv, d_cat, pairs, B, Q,
                                            d_pos_out=d_qkpos_all[:, l * 512:(l + 1) * 512])
        P.acc_gw(f"d{l}.q_w", d_qkv_obj, x, rows=1536)
        if _TC_DEC_ALL:
            ops.gemm(d_qkv_obj, P.w(f"d{l}.q_w", rows=1536), b_kn=True, add=dx, out=dx)
        else:
            dx.addmm_(d_qkv_obj, P.w(f"d{l}.q_w", rows=1536))
        P.join(0)
        dx[:, 256:] += dxr
        P._keep += [dsin, dt2, dxr]
        P.reduce_decoder_layer(l)  # data parallel: this layer's own weight gradients are final
        return dx

    # ------------------------------------------------------------------ whole path
    def forward(self, x: Tensor, pos: Tensor, bits: Tensor, kpm: Tensor, sel: Tensor, pos_embed: Tensor, sine: Tensor,
                centers: Tensor, B: int, N: int, Q: int, lam: float = 0.5, pairs_override=None, aux=None):
        """x, pos bf16 [B*N,256]; sel bf16 [B*Q,512]; pos_embed, sine bf16 [B*Q,256]; centers fp32 [B*Q,2].
        Returns (dec_out bf16 [B*Q,512], enc_out bf16 [B*N,256])."""
        P, Ld = self.P, self.Ld
        self._bbox_cache = None  # the box head is trained: re-cast it every step
        self.dc = self._drop_config()  # dropout in force for this forward AND its backward
        enc_saved = []
        self._attn_bits = self._gen_attn_bits(B, N, x.device)
        for l in range(self.Le):
            x, sv = self._enc_fwd(l, x, pos, bits, B, N)
            enc_saved.append(sv)
        enc = x
        # fine_pos = pos * encoder._pos_scale(enc_out)   (model.py:89-92)
        if _TC:
            hfp = ops.gemm(enc, P.w("e.ps0_w"), bias=P.f("e.ps0_b"), relu=True)
            fine = ops.gemm(hfp, P.w("e.ps2_w"), bias=P.f("e.ps2_b"), mul=pos)
        else:
            hfp = _mm_bias_relu(enc, P.w("e.ps0_w"), P.w("e.ps0_b"))
            fine = ops.mul(_mm_bias(hfp, P.w("e.ps2_w"), P.w("e.ps2_b")), pos)
        # projections of layer-invariant inputs for ALL decoder layers: three packed GEMMs
        if _TC_DEC_ALL:
            kv_all = ops.gemm(enc, P.w("d0.ke_w", rows=Ld * 512))
            kpos_all = ops.gemm(fine, P.w("d0.kp_w", rows=Ld * 256))
            qkpos_all = ops.gemm(pos_embed, P.w("d0.sqp_w", rows=Ld * 512))
        else:
            kv_all = torch.mm(enc, P.w("d0.ke_w", rows=Ld * 512).t())
            kpos_all = torch.mm(fine, P.w("d0.kp_w", rows=Ld * 256).t())
            qkpos_all = torch.mm(pos_embed, P.w("d0.sqp_w", rows=Ld * 512).t())
        dec_saved = []
        y = sel
        for l in range(Ld):
            y, sv, a = self._dec_fwd(l, y, kv_all, kpos_all, qkpos_all, sine, centers, bits, B, Q, N, lam,
                                     None if pairs_override is None else pairs_override[l])
            dec_saved.append(sv)
            if aux is not None:
                aux.append(a)
        self.saved = (enc_saved, dec_saved, enc, hfp, fine, pos, bits, kpm, kv_all, kpos_all, pos_embed, sine,
                      B, N, Q, lam)
        return y, enc

    def backward(self, d_dec: Tensor, d_enc_ext: Optional[Tensor] = None) -> Tensor:
        """Overwrites the .grad of every encoder/decoder parameter (= zero_grad + backward) and returns
        the gradient w.r.t. the encoder input tokens (bf16 [B*N,256])."""
        P, Ld = self.P, self.Ld
        enc_saved, dec_saved, enc, hfp, fine, pos, bits, kpm, kv_all, kpos_all, pos_embed, sine, B, N, Q, lam = self.saved
        P.begin_backward()
        d_kv_all = torch.empty_like(kv_all)
        d_kpos_all = torch.empty_like(kpos_all)
        d_qkpos_all = torch.empty(B * Q, Ld * 512, dtype=BF16, device=enc.device)
        ctx = (kv_all, kpos_all, sine, bits, kpm, B, Q, N, lam)
        dy = d_dec.contiguous()
        for l in reversed(range(Ld)):
            dy = self._dec_bwd(l, dy, dec_saved[l], ctx, d_kv_all, d_kpos_all, d_qkpos_all)
        P.acc_gw("d0.ke_w", d_kv_all, enc, rows=Ld * 512)
        d_enc = ops.gemm(d_kv_all, P.w("d0.ke_w", rows=Ld * 512), b_kn=True) if _TC_DEC_ALL else \
            torch.mm(d_kv_all, P.w("d0.ke_w", rows=Ld * 512))
        P.acc_gw("d0.kp_w", d_kpos_all, fine, rows=Ld * 256)
        d_fine = ops.gemm(d_kpos_all, P.w("d0.kp_w", rows=Ld * 256), b_kn=True) if _TC_DEC_ALL else \
            torch.mm(d_kpos_all, P.w("d0.kp_w", rows=Ld * 256))
        P.acc_gw("d0.sqp_w", d_qkpos_all, pos_embed, rows=Ld * 512)
        P.reduce_decoder_grads()  # data parallel: the decoder's gradients are final -> exchange them under the encoder backward
        du = ops.mul(d_fine, pos)
        if d_enc_ext is not None:
            d_enc += d_enc_ext
        dx = self._pos_scale_bwd("e", du, hfp, enc, d_enc)
        for l in reversed(range(self.Le)):
            dx = self._enc_bwd(l, dx, enc_saved[l], pos, bits, B, N)
        P.end_backward()
        self.saved = None
        return dx


class _RuntimeFn(torch.autograd.Function):
    """Splices the hand-written forward/backward into torch autograd (heads + loss stay in autograd)."""

    @staticmethod
    def forward(ctx, rt: HotPathRuntime, anchor, x, pos, bits, kpm, sel, pos_embed, sine, centers, B, N, Q,
                pairs_override, aux):
        # `anchor` is a dummy leaf with requires_grad=True: the parameters are not autograd inputs here, so
        # without it the outputs would not require grad and backward() would never be called.
        ctx.rt = rt
        ctx.set_materialize_grads(False)
        dec, enc = rt.forward(x, pos, bits, kpm, sel, pos_embed, sine, centers, B, N, Q, pairs_override=pairs_override,
                              aux=aux)
        return dec, enc

    @staticmethod
    def backward(ctx, d_dec, d_enc):
        dx = ctx.rt.backward(d_dec, d_enc)
        return (None, None, dx) + (None,) * 12
