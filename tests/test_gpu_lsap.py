"""-m gpu: destr_lsap_blockdiag == scipy.optimize.linear_sum_assignment, assignment by assignment (bit-exact
requirement of the matcher), including ties, T > Q, empty images and the NaN case."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

pytestmark = pytest.mark.gpu


def _run(blocks, Q, t_max):
    from object_detection_destr_b200 import ops
    B = len(blocks)
    sizes = [blk.shape[1] for blk in blocks]
    offs = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
    flat = torch.cat([torch.as_tensor(blk, dtype=torch.float32).reshape(-1) for blk in blocks] + [torch.zeros(1)])
    pi, ti, valid, status = ops.lsap_blockdiag(flat.cuda(), offs.cuda(), B, Q, t_max)
    return pi.cpu(), ti.cpu(), valid.cpu(), status.cpu()


def _check(blocks, Q, t_max):
    pi, ti, valid, status = _run(blocks, Q, t_max)
    for b, blk in enumerate(blocks):
        assert int(status[b]) == 0
        if blk.shape[1] == 0:
            assert not bool(valid[b].any())
            continue
        r, c = linear_sum_assignment(np.asarray(blk, dtype=np.float32).astype(np.float64))
        k = len(r)
        assert int(valid[b].sum()) == k and bool(valid[b, :k].all())
        assert pi[b, :k].tolist() == list(r) and ti[b, :k].tolist() == list(c), (b, blk.shape)
        assert (pi[b, k:] == Q).all() and (ti[b, k:] == 0).all()


def test_random_costs_config2_and_config5_shapes():
    rng = np.random.default_rng(0)
    for Q, t_max, B in ((100, 40, 8), (300, 40, 16), (60, 40, 5)):
        for _ in range(4):
            sizes = rng.integers(0, t_max + 1, size=B)
            _check([rng.standard_normal((Q, t)).astype(np.float32) for t in sizes], Q, t_max)


def test_more_targets_than_queries_and_ties():
    rng = np.random.default_rng(1)
    _check([rng.standard_normal((10, t)).astype(np.float32) for t in (25, 10, 9, 40)], 10, 40)
    for _ in range(40):  # tie-heavy small integer costs: the tie rule decides the assignment
        Q = int(rng.integers(1, 14))
        _check([rng.integers(0, 3, size=(Q, int(rng.integers(1, 14)))).astype(np.float32) for _ in range(6)], Q, 14)
    _check([np.zeros((6, 6), np.float32), np.ones((9, 4), np.float32)[:6], np.ones((6, 9), np.float32)], 6, 9)


def test_nan_block_is_reported_not_assigned():
    blocks = [np.random.default_rng(2).standard_normal((20, 5)).astype(np.float32) for _ in range(3)]
    blocks[1][3, 2] = np.nan
    pi, ti, valid, status = _run(blocks, 20, 8)
    assert status.tolist() == [0, 1, 0] and not bool(valid[1].any()) and int(valid[0].sum()) == 5
