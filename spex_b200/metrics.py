"""Ranking metrics with the exact definitions of the reference's utility1/metrics.py, vectorised.

    recall_at_k (metrics.py:88-94)   sum(r[:k]) / all_pos_num           (0 if all_pos_num == 0)
    dcg_at_k    (metrics.py:43-58)   sum(r[:k] / log2(2..k+1))          (method 1)
    ndcg_at_k   (metrics.py:61-71)   dcg(r, k) / dcg(sorted(r, desc), k) (0 if the ideal is 0)

`r` is the 0/1 hit list of the ranked candidates (length K_max = 50 in Test()); note that the
ideal DCG is computed from the hits INSIDE r only, exactly as the reference does.  All arithmetic
is float64 like numpy's default there (np.asfarray -> float64).
"""
from __future__ import annotations

import numpy as np


def _as_f64(r):
    return np.asarray(r, dtype=np.float64)


def dcg_at_k(r, k, method=1):
    r = _as_f64(r)[:k]
    if not r.size:
        return 0.0
    if method == 1:
        return float(np.sum(r / np.log2(np.arange(2, r.size + 2))))
    if method == 0:
        return float(r[0] + np.sum(r[1:] / np.log2(np.arange(2, r.size + 1))))
    raise ValueError("method must be 0 or 1.")


def ndcg_at_k(r, k, method=1):
    ideal = dcg_at_k(sorted(r, reverse=True), k, method)
    if not ideal:
        return 0.0
    return dcg_at_k(r, k, method) / ideal


def recall_at_k(r, k, all_pos_num):
    if all_pos_num == 0:
        return 0.0
    return float(np.sum(_as_f64(r)[:k]) / all_pos_num)


def hit_at_k(r, k):
    return 1.0 if np.sum(np.asarray(r)[:k]) > 0 else 0.0


def precision_at_k(r, k):
    assert k >= 1
    return float(np.mean(np.asarray(r)[:k]))


def batch_recall_ndcg(R: np.ndarray, n_pos: np.ndarray, Ks):
    """Per-user recall/ndcg for a hit matrix R [n_users, K_max] (0/1): two float64 [n_users, len(Ks)].

    Row-wise identical to calling recall_at_k / ndcg_at_k on every row.
    """
    R = np.asarray(R, dtype=np.float64)
    n, kmax = R.shape
    disc = 1.0 / np.log2(np.arange(2, kmax + 2))
    ideal_r = -np.sort(-R, axis=1)
    recall = np.zeros((n, len(Ks)))
    ndcg = np.zeros((n, len(Ks)))
    n_pos = np.asarray(n_pos, dtype=np.float64)
    for j, k in enumerate(Ks):
        kk = min(k, kmax)
        hits = R[:, :kk].sum(1)
        recall[:, j] = np.where(n_pos > 0, hits / np.where(n_pos > 0, n_pos, 1.0), 0.0)
        # row-wise sums in the same left-to-right order as np.sum over a short vector
        dcg = (R[:, :kk] / np.log2(np.arange(2, kk + 2))).sum(1)
        idcg = (ideal_r[:, :kk] / np.log2(np.arange(2, kk + 2))).sum(1)
        ndcg[:, j] = np.where(idcg > 0, dcg / np.where(idcg > 0, idcg, 1.0), 0.0)
    return recall, ndcg


def fullrank_recall_ndcg(topk_idx: np.ndarray, truth_rowptr: np.ndarray, truth_col: np.ndarray, k: int):
    """Recall@k / NDCG@k of full-ranking lists against per-user ground-truth sets (CSR), upstream
    LightGCN definitions: recall = hits / |truth|, ndcg = dcg / idcg(min(k, |truth|))."""
    n = topk_idx.shape[0]
    hits = np.zeros((n, k), dtype=np.float64)
    for u in range(n):
        t = truth_col[truth_rowptr[u]: truth_rowptr[u + 1]]
        if t.size:
            hits[u] = np.isin(topk_idx[u, :k], t)
    n_t = np.diff(truth_rowptr).astype(np.float64)
    disc = 1.0 / np.log2(np.arange(2, k + 2))
    dcg = (hits * disc).sum(1)
    cum = np.concatenate([[0.0], np.cumsum(disc)])
    idcg = cum[np.minimum(n_t, k).astype(np.int64)]
    valid = n_t > 0
    recall = np.where(valid, hits.sum(1) / np.where(valid, n_t, 1.0), 0.0)
    ndcg = np.where(idcg > 0, dcg / np.where(idcg > 0, idcg, 1.0), 0.0)
    return recall, ndcg
