"""CPU: the batched SetCriterion drop-in against the reference golden losses and the oracle."""
import torch

from oracle import destr_oracle as O


def _crit(num_classes):
    import importlib.util, os, sys, types
    # matcher.py imports the package (which needs the built CUDA lib for `ops`); the criterion itself is
    # pure torch, so load the module file with a stub `ops` when the library is not built.
    try:
        from object_detection_destr_b200.matcher import SetCriterion
    except ImportError:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        pkg = types.ModuleType("object_detection_destr_b200")
        pkg.__path__ = [os.path.join(root, "object_detection_destr_b200")]
        pkg.ops = types.ModuleType("object_detection_destr_b200.ops")
        sys.modules.setdefault("object_detection_destr_b200", pkg)
        sys.modules.setdefault("object_detection_destr_b200.ops", pkg.ops)
        spec = importlib.util.spec_from_file_location("object_detection_destr_b200.matcher",
                                                      os.path.join(root, "object_detection_destr_b200", "matcher.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        SetCriterion = mod.SetCriterion
    return SetCriterion(num_classes, matcher=None)


def test_criterion_matches_reference_golden(golden):
    targets = [{"labels": l, "boxes": b} for l, b in zip(golden["m_labels"], golden["m_tboxes"])]
    outputs = {"pred_class": golden["m_logits"], "pred_boxes": golden["m_boxes"]}
    got = _crit(2)(outputs, targets, indices=golden["m_wol1"])
    for k in ("class", "bbox", "ciou"):
        torch.testing.assert_close(got[k].reshape(()), golden["crit_out"][k].reshape(()), atol=2e-6, rtol=1e-5)


def test_criterion_91_classes_vs_oracle_with_grad():
    g = torch.Generator().manual_seed(3)
    B, Q, C = 4, 30, 91
    logits = torch.randn(B, Q, C, generator=g, requires_grad=True)
    boxes = torch.cat([0.1 + 0.8 * torch.rand(B, Q, 2, generator=g), 0.03 + 0.4 * torch.rand(B, Q, 2, generator=g)], -1).requires_grad_()
    labels, tboxes = O.make_targets(B, seed=5, max_t=40, num_cls=C)
    labels[1], tboxes[1] = labels[1][:0], tboxes[1][:0]  # an image without targets
    idx = O.hungarian_match(O.match_cost_blocks(logits.detach(), boxes.detach(), labels, tboxes, 0.5, 0.0, 0.5, False))
    ref = O.set_criterion(logits, boxes, labels, tboxes, idx, C)
    (ref["class"] + ref["bbox"] + ref["ciou"]).sum().backward()
    gl, gb = logits.grad.clone(), boxes.grad.clone()
    logits.grad = boxes.grad = None
    targets = [{"labels": l, "boxes": b} for l, b in zip(labels, tboxes)]
    got = _crit(C)({"pred_class": logits, "pred_boxes": boxes}, targets, indices=idx)
    for k in ("class", "bbox", "ciou"):
        torch.testing.assert_close(got[k].reshape(()), ref[k].reshape(()), atol=2e-6, rtol=1e-5)
    (got["class"] + got["bbox"] + got["ciou"]).sum().backward()
    torch.testing.assert_close(logits.grad, gl, atol=1e-7, rtol=1e-4)
    torch.testing.assert_close(boxes.grad, gb, atol=1e-6, rtol=1e-4)


def test_static_loss_equals_criterion():
    """engine.set_loss_static (padded, branch-free, CUDA-graph friendly) == SetCriterion == oracle."""
    import importlib.util, os, sys
    crit = _crit(91)
    mod = sys.modules[type(crit).__module__]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        from object_detection_destr_b200.engine import set_loss_static
    except ImportError:
        spec = importlib.util.spec_from_file_location("object_detection_destr_b200.engine",
                                                      os.path.join(root, "object_detection_destr_b200", "engine.py"))
        eng = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(eng)
        set_loss_static = eng.set_loss_static
    g = torch.Generator().manual_seed(4)
    B, Q, C, tm = 4, 30, 91, 40
    logits = torch.randn(B, Q, C, generator=g)
    boxes = torch.cat([0.1 + 0.8 * torch.rand(B, Q, 2, generator=g), 0.03 + 0.4 * torch.rand(B, Q, 2, generator=g)], -1)
    labels, tboxes = O.make_targets(B, seed=6, max_t=tm, num_cls=C)
    labels[2], tboxes[2] = labels[2][:0], tboxes[2][:0]
    idx = O.hungarian_match(O.match_cost_blocks(logits, boxes, labels, tboxes, 0.5, 0.0, 0.5, False))
    ref = O.set_criterion(logits, boxes, labels, tboxes, idx, C)
    n = min(Q, tm)
    pi, ti, valid = torch.full((B, n), Q), torch.zeros(B, n, dtype=torch.int64), torch.zeros(B, n, dtype=torch.bool)
    tl, tb = torch.ones(B, tm, dtype=torch.int64), torch.zeros(B, tm, 4)
    for b, (i, j) in enumerate(idx):
        k = i.numel()
        pi[b, :k], ti[b, :k], valid[b, :k] = i, j, True
        tl[b, :labels[b].numel()], tb[b, :labels[b].numel()] = labels[b], tboxes[b]
    got = set_loss_static(logits, boxes, tl, tb, pi, ti, valid, C)
    for k in ("class", "bbox", "ciou"):
        torch.testing.assert_close(got[k].reshape(()), ref[k].reshape(()), atol=2e-6, rtol=1e-5)
