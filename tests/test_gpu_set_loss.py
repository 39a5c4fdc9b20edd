"""-m gpu: the fused set-loss kernel (csrc/set_loss.cu) vs (a) the CPU oracle's SetCriterion restatement
(loss values AND, through the oracle's own autograd, gradients) and (b) torch autograd of the batched restatement on the
same device (gradients).
fp32 throughout; tolerance 2e-5 relative to the largest entry (different summation order only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(B, Q, C, seed, empty_image=False, clamp_heavy=False):
    from oracle import destr_oracle as O
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, Q, C, generator=g) * 2
    boxes = torch.cat([0.1 + 0.8 * torch.rand(B, Q, 2, generator=g), 0.03 + 0.4 * torch.rand(B, Q, 2, generator=g)], -1)
    if clamp_heavy:  # boxes hanging over the image border: the clamps in from_cxcyhw_to_xyxy become active
        boxes[:, ::3, 2:] += 0.9
    labels, tboxes = O.make_targets(B, seed=seed, max_t=min(40, Q), num_cls=C)
    if empty_image:
        labels[1], tboxes[1] = labels[1][:0], tboxes[1][:0]
    idx = O.hungarian_match(O.match_cost_blocks(logits, boxes, labels, tboxes, 0.5, 0.0, 0.5, with_l1=False))
    return logits, boxes, labels, tboxes, idx


def _pad(labels, tboxes, idx, B, Q, t_max):
    n = min(Q, t_max)
    tl = torch.ones(B, t_max, dtype=torch.int64)
    tb = torch.zeros(B, t_max, 4)
    pi = torch.full((B, n), Q, dtype=torch.int64)
    ti = torch.zeros(B, n, dtype=torch.int64)
    valid = torch.zeros(B, n, dtype=torch.bool)
    for b in range(B):
        t = labels[b].numel()
        tl[b, :t], tb[b, :t] = labels[b], tboxes[b]
        k = idx[b][0].numel()
        pi[b, :k], ti[b, :k], valid[b, :k] = idx[b][0], idx[b][1], True
    return tl, tb, pi, ti, valid


@pytest.mark.parametrize("B,Q,C,seed,empty,clampy", [(8, 100, 91, 0, False, False), (3, 40, 2, 1, True, False),
                                                     (4, 300, 91, 2, False, True), (1, 7, 5, 3, False, False)])
def test_fused_set_loss_matches_oracle_and_autograd(B, Q, C, seed, empty, clampy):
    from oracle import destr_oracle as O
    from object_detection_destr_b200 import ops
    from object_detection_destr_b200.engine import set_loss_static
    logits, boxes, labels, tboxes, idx = _case(B, Q, C, seed, empty, clampy)
    w = (0.5, 0.3, 0.7)
    ref = O.set_criterion(logits, boxes, labels, tboxes, idx, C)
    tl, tb, pi, ti, valid = (t.cuda() for t in _pad(labels, tboxes, idx, B, Q, 40))
    lg, bx = logits.cuda().requires_grad_(), boxes.cuda().requires_grad_()
    total, losses = ops.set_loss(lg, bx, tl, tb, pi, ti, valid, w)
    total.backward()
    for i, k in enumerate(("class", "bbox", "ciou")):
        assert abs(float(losses[i]) - float(ref[k])) <= 2e-5 * max(1.0, abs(float(ref[k]))), (k, float(losses[i]), float(ref[k]))
    assert abs(float(total) - sum(wi * float(ref[k]) for wi, k in zip(w, ("class", "bbox", "ciou")))) <= 5e-5
    # gradients: torch autograd through the batched restatement
    lg2, bx2 = logits.cuda().requires_grad_(), boxes.cuda().requires_grad_()
    st = set_loss_static(lg2, bx2, tl, tb, pi, ti, valid, C)
    (w[0] * st["class"] + w[1] * st["bbox"] + w[2] * st["ciou"]).backward()
    for got, exp, name in ((lg.grad, lg2.grad, "dlogits"), (bx.grad, bx2.grad, "dboxes")):
        tol = 2e-5 * float(exp.abs().max()) + 1e-9
        assert float((got - exp).abs().max()) <= tol, (name, float((got - exp).abs().max()), float(exp.abs().max()))
    # gradients vs the ORACLE's autograd (CPU fp32, per-image loop exactly as criterion.py:57-79)
    lo, bo = logits.clone().requires_grad_(), boxes.clone().requires_grad_()
    ro = O.set_criterion(lo, bo, labels, tboxes, idx, C)
    (w[0] * ro["class"] + w[1] * ro["bbox"] + w[2] * ro["ciou"]).sum().backward()
    from parity_log import record
    for got, exp, name in ((lg.grad, lo.grad, "dlogits"), (bx.grad, bo.grad, "dboxes")):
        err = float((got.cpu() - exp).abs().max())
        tol = 2e-5 * float(exp.abs().max()) + 1e-9
        record(f"set_loss_B{B}_Q{Q}_C{C}", name + ".max_abs_vs_oracle_autograd", err, tol)
        assert err <= tol, (name, err, float(exp.abs().max()))
    # upstream gradient scaling
    lg3, bx3 = logits.cuda().requires_grad_(), boxes.cuda().requires_grad_()
    t3, _ = ops.set_loss(lg3, bx3, tl, tb, pi, ti, valid, w)
    (2.5 * t3).backward()
    assert torch.allclose(lg3.grad, 2.5 * lg.grad, rtol=1e-6, atol=0) and torch.allclose(bx3.grad, 2.5 * bx.grad, rtol=1e-6, atol=0)
