"""-m gpu: the CUDA-graph training step replays exactly what the eager step computes."""
from argparse import Namespace

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gpu_lsa", [True, False])
def test_graphed_step_matches_eager(gpu_lsa):
    import bench
    from object_detection_destr_b200.encoder import disable_dropout
    from object_detection_destr_b200.engine import GraphedTrainStep
    from object_detection_destr_b200.hotpath import TransformerHalf
    cfg = dict(bench.CFG, B=2, L=2, H=10, W=14, Q=60)
    torch.manual_seed(0)
    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=2, num_decoder_blocks=2, num_cls=cfg["C"]))
    disable_dropout(model).cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=0.0, fused=True, capturable=True)
    eng = GraphedTrainStep(model, opt, B=2, H=10, W=14, Q=60, num_classes=cfg["C"], t_max=40, gpu_lsa=gpu_lsa)
    batches = [bench.make_batch(0, s, 2, cfg, padded=True) for s in range(3)]
    eng.load_batch(*batches[0])
    eng.capture(warmup=2)
    assert eng.launches_per_step > 50
    for bt in batches:
        eng.load_batch(*bt)
        l_graph = float(eng.step())
        g_graph = model._encoder._encoder[0].fc1.weight.grad.clone()
        l_eager = float(eng.eager_step())
        g_eager = model._encoder._encoder[0].fc1.weight.grad
        assert abs(l_graph - l_eager) <= 1e-5 * max(1.0, abs(l_eager)), (l_graph, l_eager)
        assert torch.allclose(g_graph, g_eager, rtol=1e-3, atol=1e-6)
        assert l_graph == l_graph and l_graph > 0


def test_flat_adamw_matches_torch_adamw():
    """The one-launch flat AdamW (csrc/optim.cu) == torch.optim.AdamW on the same gradients, 3 steps."""
    import copy
    from object_detection_destr_b200.hotpath import TransformerHalf
    torch.manual_seed(1)
    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=1, num_decoder_blocks=1, num_cls=7)).cuda()
    opt = model.make_optimizer(lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.05)
    P = model.runtime().P
    ref_p = [p.detach().clone().requires_grad_() for p in P.params]
    ref = torch.optim.AdamW(ref_p, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.05)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(3):
        P.g32.copy_(torch.randn(P.g32.shape, generator=g, device="cuda") * 0.1)
        for rp, p in zip(ref_p, P.params):
            rp.grad = p.grad.detach().clone()
        opt.step()
        ref.step()
    # a step straight after backward(): the weight gradients are still bf16 (P.g16) and everything is a sum over
    # `world` ranks -- the kernel widens, scales by 1/world and leaves the fp32 values in .grad
    P.g16.copy_((torch.randn(P.g16.shape, generator=g, device="cuda") * 0.1).bfloat16())
    P.g32.copy_(torch.randn(P.g32.shape, generator=g, device="cuda") * 0.1)
    b0 = P.g16_begin()  # encoder weight gradients below b0 are accumulated in fp32 by the tcgen05 dW kernel
    expect = torch.cat([P.g32[:b0], P.g16[b0:].float(), P.g32[P.nW:]]) * 0.5
    P.g16_pending, P.grad_scale = True, 0.5
    for rp, p in zip(ref_p, P.params):
        off = p.grad.storage_offset()  # .grad is a view of the flat fp32 gradient buffer
        rp.grad = expect[off:off + p.numel()].view(p.shape).clone()
    opt.step()
    ref.step()
    assert not P.g16_pending and torch.equal(P.g32, expect)
    for rp, p in zip(ref_p, P.params):
        assert torch.allclose(p, rp, rtol=2e-6, atol=2e-7), float((p - rp).abs().max())
    for name in ("e0.fc1_w", "d0.q_w"):  # bf16 shadows follow the masters
        assert torch.equal(P.w(name), P.f(name).to(torch.bfloat16))


def test_device_assignment_step_equals_host_assignment_step():
    """Same weights, same batch: the single-graph step (device LSAP + fused loss) and the reference arrangement
    (host scipy + autograd loss) produce the same loss and gradients."""
    import bench
    from object_detection_destr_b200.encoder import disable_dropout
    from object_detection_destr_b200.engine import GraphedTrainStep
    from object_detection_destr_b200.hotpath import TransformerHalf
    cfg = dict(bench.CFG, B=3, L=1, H=10, W=14, Q=60)
    res = []
    for gpu_lsa, fused in ((True, True), (False, False)):
        torch.manual_seed(0)
        model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=1, num_decoder_blocks=1, num_cls=cfg["C"]))
        disable_dropout(model).cuda().train()
        opt = torch.optim.AdamW(model.parameters(), lr=0.0, fused=True, capturable=True)
        eng = GraphedTrainStep(model, opt, B=3, H=10, W=14, Q=60, num_classes=cfg["C"], t_max=40, gpu_lsa=gpu_lsa,
                               fused_loss=fused)
        eng.load_batch(*bench.make_batch(0, 5, 3, cfg, padded=True))
        loss = float(eng.eager_step())
        eng.raise_if_invalid()
        res.append((loss, model._decoder._decoder[0]._sa_proj_to_q_obj.weight.grad.clone(), eng.s_pi.clone(), eng.s_ti.clone()))
    assert torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3])  # identical assignments
    assert abs(res[0][0] - res[1][0]) <= 2e-5 * max(1.0, abs(res[1][0]))
    assert torch.allclose(res[0][1], res[1][1], rtol=2e-2, atol=1e-6)


def test_prefetched_steps_equal_plain_steps():
    """Input pipelining (prefetch on a copy stream + device hand-over) feeds the graph the same batches."""
    import bench
    from object_detection_destr_b200.encoder import disable_dropout
    from object_detection_destr_b200.engine import GraphedTrainStep
    from object_detection_destr_b200.hotpath import TransformerHalf
    cfg = dict(bench.CFG, B=2, L=1, H=10, W=14, Q=60)
    torch.manual_seed(0)
    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=1, num_decoder_blocks=1, num_cls=cfg["C"]))
    disable_dropout(model).cuda().train()
    opt = model.make_optimizer(lr=0.0)
    eng = GraphedTrainStep(model, opt, B=2, H=10, W=14, Q=60, num_classes=cfg["C"], t_max=40)
    batches = [bench.make_batch(0, s, 2, cfg, padded=True) for s in range(4)]
    pinned = [tuple(t.pin_memory() for t in bt[:4]) + (bt[4], bt[5]) for bt in batches]
    eng.load_batch(*batches[0])
    eng.capture(warmup=2)
    plain = []
    for bt in batches:
        eng.load_batch(*bt)
        plain.append(float(eng.step()))
    eng.prefetch(*pinned[0])
    piped = []
    for s in range(4):
        loss_t = eng.step_prefetched()
        eng.prefetch(*pinned[(s + 1) % 4])
        piped.append(float(loss_t))
    assert piped == plain, (piped, plain)


def test_load_batch_step_back_to_back_without_host_sync():
    """load_batch() re-packs targets into ONE set of pinned arrays and enqueues non-blocking copies; a
    caller looping load_batch()+step() without reading anything back must not overwrite them before the previous
    step's copies ran.  Losses of an un-synchronised loop == losses of the same loop with a sync after every step."""
    import bench
    from object_detection_destr_b200.encoder import disable_dropout
    from object_detection_destr_b200.engine import GraphedTrainStep
    from object_detection_destr_b200.hotpath import TransformerHalf
    cfg = dict(bench.CFG, B=2, L=1, H=10, W=14, Q=60)
    torch.manual_seed(0)
    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=1, num_decoder_blocks=1, num_cls=cfg["C"]))
    disable_dropout(model).cuda().train()
    opt = model.make_optimizer(lr=0.0)   # lr 0: every pass sees the same weights
    eng = GraphedTrainStep(model, opt, B=2, H=10, W=14, Q=60, num_classes=cfg["C"], t_max=40)
    batches = [bench.make_batch(0, s, 2, cfg, padded=True) for s in range(12)]
    dev_batches = [tuple(t.cuda() for t in bt[:4]) + (bt[4], bt[5]) for bt in batches]
    eng.load_batch(*dev_batches[0])
    eng.capture(warmup=2)
    assert eng.graph_kernel_nodes is None or eng.graph_kernel_nodes >= eng.launches_per_step

    def run(sync: bool):
        out = []
        for bt in dev_batches:
            eng.load_batch(*bt)
            out.append(eng.step().clone())   # device-side copy of the loss, no host read
            if sync:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        return [float(t) for t in out]

    assert run(False) == run(True)


def test_out_of_range_label_raises_like_the_reference():
    import bench
    from object_detection_destr_b200.engine import GraphedTrainStep
    from object_detection_destr_b200.hotpath import TransformerHalf
    cfg = dict(bench.CFG, B=2, L=1, H=10, W=14, Q=60)
    model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=1, num_decoder_blocks=1, num_cls=cfg["C"])).cuda()
    eng = GraphedTrainStep(model, None, B=2, H=10, W=14, Q=60, num_classes=cfg["C"], t_max=40)
    f, m, s, c, labels, boxes = bench.make_batch(0, 0, 2, cfg)
    labels[1][0] = cfg["C"]   # one past the last class
    with pytest.raises(IndexError):
        eng.load_batch(f, m, s, c, labels, boxes)


def test_overlapped_optimizer_step_equals_plain_step():
    """FlatAdamW with the decoder's share of the step launched under the encoder backward (engine default) produces
    the same parameters as the single flat launch after the backward: 3 training steps from identical states.
    Conditioning: several fp32 gradient accumulations use atomics (bias / LayerNorm column sums, the split-K dW, the
    TMA reduce-adds of the attention dQ), so two runs of the SAME path differ in the last bits of some gradients; with
    the default eps = 1e-8 Adam turns a 1e-8 wobble of a near-zero gradient into an lr-sized parameter difference
    (measured: either path against itself, fully serialised with CUDA_LAUNCH_BLOCKING=1, lands 1.25e-3 apart on
    d1.q_w about every other run -- tools/dbg_overlap_opt.py).  eps = 1e-4 keeps the update linear in such gradients,
    so what is compared is the split of the launch, not that noise."""
    import bench
    from object_detection_destr_b200.encoder import disable_dropout
    from object_detection_destr_b200.engine import GraphedTrainStep
    from object_detection_destr_b200.hotpath import TransformerHalf
    cfg = dict(bench.CFG, B=2, L=2, H=10, W=14, Q=60)
    batches = [bench.make_batch(0, s, 2, cfg, padded=True) for s in range(3)]
    finals = []
    for overlap in (True, False):
        torch.manual_seed(0)
        model = TransformerHalf(Namespace(hidden_dim=256, num_encoder_blocks=2, num_decoder_blocks=2, num_cls=cfg["C"]))
        disable_dropout(model).cuda().train()
        opt = model.make_optimizer(lr=1e-3, eps=1e-4)
        eng = GraphedTrainStep(model, opt, B=2, H=10, W=14, Q=60, num_classes=cfg["C"], t_max=40)
        eng.overlap_opt = overlap
        for bt in batches:
            eng.load_batch(*bt)
            eng.eager_step()
        torch.cuda.synchronize()
        finals.append((model.runtime().P.m32.clone(), float(opt.t)))
    (pa, ta), (pb, tb) = finals
    assert ta == tb == 3.0
    # (bias / LayerNorm gradients are accumulated with atomics: equal up to summation order)
    assert float((pa - pb).abs().max()) <= 2e-5, float((pa - pb).abs().max())
    assert float((pa - pb).abs().mean()) <= 1e-7
