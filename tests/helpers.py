"""Shared helpers for the parity tests (oracle = checker, CUDA path = thing under test)."""
import argparse

import numpy as np
import torch

from oracle import lightgcn_oracle as O
from spex_b200.graph import build_interaction_csr, build_norm_adj


def make_args(**kw):
    d = dict(recdim=64, layer=3, keepprob=0.6, A_split=0, dropout=0, a_fold=4, dataset="synthetic",
             lr=1e-3, seed=2020)
    d.update(kw)
    return argparse.Namespace(**d)


def random_graph(n_users, m_items, n_inter, seed, hub_items=0, hub_degree=0):
    """(users, items) unique pairs; optional hub items with a huge degree (long rows)."""
    u, i = O.random_bipartite(n_users, m_items, n_inter, seed)
    if hub_items:
        rng = np.random.default_rng(seed + 1)
        hu = rng.integers(0, n_users, hub_items * hub_degree)
        hi = np.repeat(np.arange(hub_items), hub_degree)
        key = np.unique(np.concatenate([u * m_items + i, hu.astype(np.int64) * m_items + hi]))
        u, i = key // m_items, key % m_items
    return u, i


def oracle_graph(u, i, n_user_rows, m_items):
    return O.to_sparse_tensor(O.norm_adj_scipy(u, i, n_user_rows, m_items))


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check_topk_against_scores(idx, val, scores, k, masked, atol):
    """idx/val [B,k] is a valid top-k of `scores` [B,m] (fp64) modulo ties within atol."""
    idx = idx.cpu().numpy()
    val = val.cpu().numpy().astype(np.float64)
    s = scores.astype(np.float64).copy()
    B, m = s.shape
    for r in range(B):
        s[r, masked[r]] = -np.inf
        n_valid = int(np.isfinite(s[r]).sum())
        kk = min(k, n_valid)
        assert (idx[r, kk:] == -1).all(), "slots beyond the unmasked items must be -1"
        ids = idx[r, :kk]
        assert len(set(ids.tolist())) == kk, "duplicate item in top-k"
        assert not np.isin(ids, masked[r]).any(), "masked item returned"
        # reported scores match the oracle's scores of the same items
        assert np.abs(val[r, :kk] - s[r, ids]).max() <= atol
        # order: non-increasing
        assert (np.diff(val[r, :kk]) <= 0).all()
        # completeness: nothing outside beats the k-th by more than atol
        if kk:
            rest = s[r].copy()
            rest[ids] = -np.inf
            assert rest.max() <= s[r, ids].min() + atol
