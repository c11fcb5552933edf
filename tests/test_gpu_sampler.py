"""Device samplers (csrc/sampler.cu) against the semantics of the reference's host loop
(LightGCN_SPEX/code/utility1/dataloader.py:250-265): a negative is NEVER a training item of its user,
negatives are uniform over the non-interacted items, the draw is deterministic in the seed; BPR triples:
the positive IS a training item, the negative is not (upstream LightGCN sampler semantics)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _graph(cuda_device, nu=300, m=200, ni=6000, seed=5):
    from helpers import random_graph
    from spex_b200 import ops
    from spex_b200.graph import build_norm_adj

    u, i = random_graph(nu, m, ni, seed)
    g = build_norm_adj(u, i, nu + 1, m)
    dg = ops.DeviceGraph.from_host(g, cuda_device)
    pairs = set(zip(u.tolist(), i.tolist()))
    return dg, u, i, pairs, nu, m


def test_negatives_never_hit_training_items_and_are_uniform(cuda_device):
    from spex_b200.dataloader import sample_negatives_device

    dg, u, i, pairs, nu, m = _graph(cuda_device)
    users = torch.from_numpy(u).to(cuda_device)
    neg = sample_negatives_device(dg.rowptr, dg.col, nu + 1, m, users, 5, seed=7)
    assert neg.shape == (u.size, 5) and int(neg.min()) >= 0 and int(neg.max()) < m
    nn = neg.cpu().numpy()
    assert not any((int(uu), int(j)) in pairs for uu, row in zip(u, nn) for j in row)
    # deterministic in the seed, different across seeds
    assert torch.equal(neg, sample_negatives_device(dg.rowptr, dg.col, nu + 1, m, users, 5, seed=7))
    assert not torch.equal(neg, sample_negatives_device(dg.rowptr, dg.col, nu + 1, m, users, 5, seed=8))
    # uniformity over the allowed items of one user with many draws (chi-square, 5 sigma)
    u0 = int(u[0])
    allowed = np.array([j for j in range(m) if (u0, j) not in pairs])
    draws = sample_negatives_device(dg.rowptr, dg.col, nu + 1, m, torch.full((40000,), u0, device=cuda_device), 5, 3)
    cnt = np.bincount(draws.cpu().numpy().ravel(), minlength=m)
    assert cnt[[j for j in range(m) if (u0, j) in pairs]].sum() == 0
    exp = 200000 / allowed.size
    chi2 = float(((cnt[allowed] - exp) ** 2 / exp).sum())
    assert abs(chi2 - allowed.size) < 5 * np.sqrt(2 * allowed.size), chi2
    # the hot flag in bit 31 of the column ids must not matter
    dg.col.bitwise_or_(torch.tensor(-(2 ** 31), dtype=torch.int32, device=cuda_device) * (dg.col % 3 == 0).int())
    assert torch.equal(neg, sample_negatives_device(dg.rowptr, dg.col, nu + 1, m, users, 5, seed=7))


def test_dense_user_falls_back_to_the_forward_walk(cuda_device):
    """A user who interacted with all items but two: the rejection loop (64 draws) may fail, the walk must
    still return one of the two free items; a user with every item gets -1."""
    from spex_b200 import ops
    from spex_b200.dataloader import sample_negatives_device
    from spex_b200.graph import build_norm_adj

    m = 500
    free = {17, 333}
    u = np.concatenate([np.zeros(m - 2, np.int64), np.ones(m, np.int64)])
    i = np.concatenate([np.array([j for j in range(m) if j not in free]), np.arange(m)])
    dg = ops.DeviceGraph.from_host(build_norm_adj(u, i, 3, m), cuda_device)
    neg = sample_negatives_device(dg.rowptr, dg.col, 3, m, torch.zeros(2000, dtype=torch.int64, device=cuda_device), 5, 1)
    vals = set(neg.cpu().numpy().ravel().tolist())
    assert vals <= free and len(vals) == 2
    neg1 = sample_negatives_device(dg.rowptr, dg.col, 3, m, torch.ones(10, dtype=torch.int64, device=cuda_device), 5, 1)
    assert bool((neg1 == -1).all())


def test_bpr_triples(cuda_device):
    from spex_b200.dataloader import uniform_sample_bpr_device

    dg, u, i, pairs, nu, m = _graph(cuda_device)
    users, pos, neg = uniform_sample_bpr_device(dg.rowptr, dg.col, nu + 1, m, nu, 50000, seed=9)
    uu, pp, nn = users.cpu().numpy(), pos.cpu().numpy(), neg.cpu().numpy()
    assert uu.min() >= 0 and uu.max() < nu
    assert all((int(a), int(b)) in pairs for a, b in zip(uu, pp))
    assert not any((int(a), int(b)) in pairs for a, b in zip(uu, nn))
    # users are uniform over users WITH interactions
    deg = np.bincount(u, minlength=nu)
    cnt = np.bincount(uu, minlength=nu)
    assert cnt[deg == 0].sum() == 0
    active = int((deg > 0).sum())
    exp = 50000 / active
    chi2 = float(((cnt[deg > 0] - exp) ** 2 / exp).sum())
    assert abs(chi2 - active) < 5 * np.sqrt(2 * active), chi2


def test_light_train_data_device_epoch(cuda_device):
    from spex_b200.dataloader import LightTrainData

    dg, u, i, pairs, nu, m = _graph(cuda_device)
    feats = np.stack([u, i], 1).tolist()
    td = LightTrainData(feats, m, None)
    users, items, labels = td.ng_sample_device(dg, nu + 1, seed=4)
    n = len(feats)
    assert users.numel() == 6 * n and float(labels.sum()) == n
    uu, ii, ll = users.cpu().numpy(), items.cpu().numpy(), labels.cpu().numpy()
    assert all(((int(a), int(b)) in pairs) == bool(c) for a, b, c in zip(uu, ii, ll))
