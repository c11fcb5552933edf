"""Sampled-candidate evaluation with the semantics of the reference's utility1/batch_test.py,
hoisted: ONE propagation for all test users instead of one per user (batch_test.py:33 calls the
model, hence computer(), once per user).

Per user (batch_test.py:28-40,72-90): candidates = 99 negatives + the held-out positive, positive
LAST; ranking = heapq.nlargest(50, rating, key=rating.get) over a dict keyed by item id, i.e. a
stable descending sort in insertion order (the positive loses exact ties, a repeated item id keeps
its first position); r = hit list; Recall@{10,20,50} and NDCG@{10,20,50}; result = sum over users
of metric / n_test_users, accumulated in user order in float64.
"""
from __future__ import annotations

import numpy as np
import torch

from . import metrics, ops

Ks = [10, 20, 50]
BATCH_SIZE = 256


def _candidate_matrix(users, testRatings, testNegatives):
    lens = {len(testNegatives[u]) + len(testRatings[u]) for u in users}
    if len(lens) != 1:
        return None
    n_c = lens.pop()
    cand = np.empty((len(users), n_c), dtype=np.int32)
    for r, u in enumerate(users):
        cand[r, : n_c - len(testRatings[u])] = testNegatives[u]
        cand[r, n_c - len(testRatings[u]):] = testRatings[u]
    return cand


def rank_hits(scores: np.ndarray, cand: np.ndarray, positives, k_max: int = 50) -> np.ndarray:
    """Hit matrix [n_users, k_max] under the reference's dict + heapq.nlargest semantics."""
    n_u, n_c = cand.shape
    order = np.argsort(-scores, axis=1, kind="stable")  # stable desc: earlier candidate wins ties
    R = np.zeros((n_u, k_max), dtype=np.float64)
    srt = np.sort(cand, axis=1)
    has_dup = (srt[:, 1:] == srt[:, :-1]).any(1)
    for r in range(n_u):
        pos = positives[r]
        if has_dup[r]:
            # dict semantics: a repeated item keeps its first slot (its score is identical anyway)
            _, first = np.unique(cand[r], return_index=True)
            keep = np.zeros(n_c, bool)
            keep[first] = True
            o = order[r][keep[order[r]]][:k_max]
        else:
            o = order[r][:k_max]
        R[r, : o.size] = np.isin(cand[r][o], pos)
    return R


@torch.no_grad()
def test(model, testRatings, testNegatives):
    """Same result dict as the reference test(): {'recall': float64[3], 'ndcg': float64[3]}."""
    users = list(testRatings.keys())
    result = {"recall": np.zeros(len(Ks)), "ndcg": np.zeros(len(Ks))}
    if not users:
        return result
    cand = _candidate_matrix(users, testRatings, testNegatives)
    # ONE propagation (+ gate for the multi-task model): the table forward(flag=1) scores against
    all_users, all_items = getattr(model, "final_embeddings", model.computer)()
    if cand is not None:
        scores = ops.score_candidates(all_users, all_items, np.asarray(users, np.int64),
                                      torch.from_numpy(cand)).cpu().numpy()
    else:  # ragged candidate lists: score user by user on the shared propagation
        scores, cand_rows = [], []
        for u in users:
            c = np.asarray(testNegatives[u] + testRatings[u], dtype=np.int32)[None]
            scores.append(ops.score_candidates(all_users, all_items, np.asarray([u], np.int64),
                                               torch.from_numpy(c)).cpu().numpy()[0])
            cand_rows.append(c[0])
        return _accumulate_ragged(users, scores, cand_rows, testRatings)
    positives = [testRatings[u] for u in users]
    R = rank_hits(scores, cand, positives, max(Ks))
    n_pos = np.array([len(p) for p in positives])
    rec, ndcg = metrics.batch_recall_ndcg(R, n_pos, Ks)
    n = len(users)
    # user-order float64 accumulation of metric / n, like result[...] += re[...] / n_test_users
    result["recall"] = np.cumsum(rec / n, axis=0)[-1]
    result["ndcg"] = np.cumsum(ndcg / n, axis=0)[-1]
    return result


def _accumulate_ragged(users, scores, cand_rows, testRatings):
    result = {"recall": np.zeros(len(Ks)), "ndcg": np.zeros(len(Ks))}
    n = len(users)
    for u, s, c in zip(users, scores, cand_rows):
        R = rank_hits(s[None], c[None], [testRatings[u]], max(Ks))[0]
        r = R[: min(len(np.unique(c)), max(Ks))]
        result["recall"] += np.array([metrics.recall_at_k(r, k, len(testRatings[u])) for k in Ks]) / n
        result["ndcg"] += np.array([metrics.ndcg_at_k(list(r), k) for k in Ks]) / n
    return result


def rec_test(model, testRatings, testNegatives):
    """Multi-task entry points call this name (batch_test.py:43-55); same computation."""
    return test(model, testRatings, testNegatives)


@torch.no_grad()
def test_fullrank(model, users, truth, k: int = 20, precision: str = "bf16"):
    """Full-ranking Recall@k / NDCG@k (north_star (3)): fused score + mask + top-k on the GPU,
    metrics on the host.  `truth` maps user -> list of held-out items."""
    users = np.asarray(list(users), dtype=np.int64)
    idx, _ = model.rank_topk(users, k=k, exclude_train=True, precision=precision)
    idx = idx.cpu().numpy()
    rp = np.zeros(users.size + 1, np.int64)
    cols = []
    for r, u in enumerate(users):
        t = truth.get(int(u), [])
        rp[r + 1] = rp[r] + len(t)
        cols.extend(t)
    rec, ndcg = metrics.fullrank_recall_ndcg(idx, rp, np.asarray(cols, np.int64), k)
    return {"recall": float(rec.mean()) if users.size else 0.0,
            "ndcg": float(ndcg.mean()) if users.size else 0.0}
