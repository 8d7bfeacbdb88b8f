"""CPU: host-side logic (views, ragged batching, synthetic generator, sharding plan)."""
import numpy as np
import torch

from lc_b200 import _native as nat
from lc_b200.pnp import cer_solver
from lc_b200.synth import make_correspondences, planar_view, full_icov_from_inv_std


def test_view_of_keeps_strides_of_planar_and_broadcast_tensors():
    t = planar_view(torch.arange(2 * 5 * 3, dtype=torch.float32).reshape(2, 5, 3))
    assert t.stride() == (15, 1, 5)
    v = nat.view_of(t)
    assert list(v.stride)[:3] == [15, 1, 5] and v.ptr == t.data_ptr()
    g = torch.zeros(5, 2).expand(4, 5, 2)
    assert list(nat.view_of(g).stride)[:3] == [0, 2, 1]
    assert nat.view_of(None).ptr is None


def test_batch_tensors_pads_ragged_lists_like_the_reference():
    p3 = [torch.ones(3, 3), 2 * torch.ones(5, 3), 3 * torch.ones(4, 3)]
    st = [torch.arange(7.0), torch.arange(7.0) + 1, torch.arange(7.0) + 2]
    K = torch.eye(3).expand(3, 3, 3)
    Kb, P, S, n = cer_solver._batch_tensors(K, p3, st, [3, 5, 4])
    assert Kb is K
    assert P.shape == (3, 5, 3) and S.shape == (3, 7) and n.tolist() == [3, 5, 4]
    assert P[0, 3:].abs().sum() == 0 and P[1].eq(2).all() and P[2, :4].eq(3).all() and P[2, 4].abs().sum() == 0


def test_make_args_type_checks():
    import pytest
    t32 = torch.zeros(2, 4, 3)
    with pytest.raises(TypeError):
        nat.make_args(2, 4, torch.float64, pts3d=t32)
    with pytest.raises(TypeError):
        nat.make_args(2, 4, torch.float16)
    a = nat.make_args(2, 4, torch.float32, pts3d=t32, max_iter=7, grad_scale=0.5)
    assert a.B == 2 and a.N == 4 and a.max_iter == 7 and a.grad_scale == 0.5 and a.pts3d.ptr == t32.data_ptr()


def test_synth_is_deterministic_and_well_formed():
    a, b = make_correspondences(3, 17, 5), make_correspondences(3, 17, 5)
    for k in a.__dataclass_fields__:
        assert torch.equal(getattr(a, k), getattr(b, k))
    assert torch.allclose(a.pose[:, :4].norm(dim=-1), torch.ones(3, dtype=torch.float64))
    assert (a.pose[:, 6] > 300).all() and (a.inv_std > 0).all()
    ic = full_icov_from_inv_std(a.inv_std, 0)
    assert (torch.linalg.eigvalsh(ic) > 0).all()


def test_shard_plan_covers_the_batch_exactly():
    from lc_b200.sharded import shard_bounds
    for B in (1, 7, 8, 1024, 1025):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
