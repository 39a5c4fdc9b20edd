"""CPU: the packed target block of the engine's batch feed (engine._TargetBlock) -- five typed views over ONE byte buffer,
16-byte aligned, disjoint, in the order of the step graph's static inputs -- and the packer writing through them."""
import numpy as np
import pytest
import torch


def _import_engine():
    try:
        from object_detection_destr_b200 import engine
    except ImportError as e:  # the package refuses to import without the built library
        pytest.skip(str(e))
    return engine


@pytest.mark.parametrize("B,tm", [(8, 40), (2, 7), (1, 1), (64, 300)])
def test_views_are_aligned_disjoint_and_typed(B, tm):
    eng = _import_engine()
    blk = eng._TargetBlock(B, tm, pin=False)
    base = blk.buf.data_ptr()
    spans = []
    for v, dt, shape in zip(blk.views(), (torch.int32, torch.int32, torch.float32, torch.int64, torch.float32),
                            ((B * tm,), (B + 1,), (B * tm, 4), (B, tm), (B, tm, 4))):
        assert v.dtype == dt and tuple(v.shape) == shape and v.is_contiguous()
        off = v.data_ptr() - base
        spans.append((off, off + v.numel() * v.element_size()))
    assert spans[0][0] == 0 and spans[1][0] == spans[0][1]          # ids then offsets: the packer's one int32 array
    for (a0, a1), (b0, b1) in zip(spans[1:], spans[2:]):
        assert b0 >= a1 and b0 % 16 == 0                               # later segments: disjoint, 16-byte aligned
    assert spans[-1][1] <= blk.buf.numel()
    assert blk.ints.data_ptr() == base and blk.ints.numel() == B * tm + B + 1


def test_packer_writes_through_the_views():
    eng = _import_engine()
    B, tm, C = 3, 5, 91
    blk = eng._TargetBlock(B, tm, pin=False)
    blk.buf.fill_(0xAB)
    labels = [torch.tensor([3, 90], dtype=torch.int64), torch.zeros(0, dtype=torch.int64), torch.tensor([7], dtype=torch.int64)]
    boxes = [torch.tensor([[0.1, 0.2, 0.3, 0.4], [0.5, 0.5, 0.9, 0.8]]), torch.zeros(0, 4), torch.tensor([[0.0, 0.0, 1.0, 1.0]])]

    class _E:  # the packer only needs these attributes
        pass
    e = _E()
    e.B, e.t_max, e.C = B, tm, C
    sizes = eng.GraphedTrainStep._pack_targets(e, labels, boxes, blk.ints, blk.flt, blk.tl, blk.tb)
    assert sizes == [2, 0, 1]
    assert blk.ids[:3].tolist() == [3, 90, 7]
    assert blk.offs.tolist() == [0, 2, 2, 3]
    np.testing.assert_allclose(blk.tboxes[:3].numpy(), torch.cat(boxes).numpy())
    assert blk.tl[0, :2].tolist() == [3, 90] and blk.tl[2, 0].item() == 7 and blk.tl[1].tolist() == [1] * tm
    np.testing.assert_allclose(blk.tb[0, :2].numpy(), boxes[0].numpy())
    assert float(blk.tb[1].abs().sum()) == 0.0
    with pytest.raises(IndexError):
        eng.GraphedTrainStep._pack_targets(e, [torch.tensor([C])] * B, [torch.zeros(1, 4)] * B, blk.ints, blk.flt, blk.tl, blk.tb)
