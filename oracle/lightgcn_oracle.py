"""CPU oracle for the LightGCN_SPEX path — TEST INFRASTRUCTURE, never shipped or measured as product.

A restatement, in the reference's own numeric backend (PyTorch CPU ops + numpy/scipy), of the
functions SURVEY.md §8(a) lists.  Each function cites the reference lines it follows (paths are
relative to /root/reference/).  The reference is pure Python whose arithmetic lives in PyTorch /
scipy library calls, so the faithful restatement calls the same library ops
(torch.sparse.mm, torch.stack/mean, BCEWithLogitsLoss, scipy sparse products) on the CPU.

Pinning: the reference has no tests or golden vectors.  This oracle is pinned against
outputs of the reference code itself, imported from /root/reference in the build container by
tests/golden/make_golden.py; the vectors are committed under tests/golden/ and checked by
tests/test_oracle_golden.py.  The north_star additions (bpr_loss, getUsersRating + top-k) do not
exist in the reference: their oracle is the upstream-LightGCN formula written with torch ops and is
labelled "north-star semantics, not reference-pinned".
"""
from __future__ import annotations

import heapq

import numpy as np
import scipy.sparse as sp
import torch

Ks = [10, 20, 50]


# ---- a1: graph -------------------------------------------------------------------------------------
def norm_adj_scipy(users, items, n_user_rows, m_items):
    """LightGCN_SPEX/code/utility1/dataloader.py:110-111,196-212 with block assembly instead of
    lil slicing: same float32 operands, same product order d_mat.dot(adj).dot(d_mat)."""
    users = np.asarray(users)
    items = np.asarray(items)
    R = sp.csr_matrix((np.ones(len(users)), (users, items)), shape=(n_user_rows, m_items))
    adj = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
    rowsum = np.array(adj.sum(axis=1))
    with np.errstate(divide="ignore"):
        d_inv = np.power(rowsum, -0.5).flatten()
    d_inv[np.isinf(d_inv)] = 0.0
    d_mat = sp.diags(d_inv)
    norm = d_mat.dot(adj).dot(d_mat).tocsr()
    norm.sort_indices()
    return norm


def to_sparse_tensor(X):
    """dataloader.py:179-185 + :221 coalesce."""
    coo = X.tocoo().astype(np.float32)
    index = torch.stack([torch.from_numpy(coo.row).long(), torch.from_numpy(coo.col).long()])
    return torch.sparse_coo_tensor(index, torch.from_numpy(coo.data), torch.Size(coo.shape)).coalesce()


# ---- a3: edge dropout ------------------------------------------------------------------------------
def dropout_graph(graph, keep_prob, rand=None):
    """utility1/model.py:46-55.  `rand` = the torch.rand(nnz) draw (None: draw it here)."""
    index = graph.indices().t()
    values = graph.values()
    if rand is None:
        rand = torch.rand(len(values))
    keep = (rand + keep_prob).int().bool()
    return torch.sparse_coo_tensor(index[keep].t(), values[keep] / keep_prob, graph.size())


# ---- a2: propagation -------------------------------------------------------------------------------
def computer(user_w, item_w, graph, n_layers):
    """utility1/model.py:66-97 (non-split branch)."""
    all_emb = torch.cat([user_w, item_w])
    embs = [all_emb]
    for _ in range(n_layers):
        all_emb = torch.sparse.mm(graph, all_emb)
        embs.append(all_emb)
    light_out = torch.mean(torch.stack(embs, dim=1), dim=1)
    return torch.split(light_out, [user_w.shape[0], item_w.shape[0]])


def computer_split(user_w, item_w, folds, n_layers):
    """utility1/model.py:84-89: serial row folds (A_split)."""
    all_emb = torch.cat([user_w, item_w])
    embs = [all_emb]
    for _ in range(n_layers):
        all_emb = torch.cat([torch.sparse.mm(f, all_emb) for f in folds], dim=0)
        embs.append(all_emb)
    light_out = torch.mean(torch.stack(embs, dim=1), dim=1)
    return torch.split(light_out, [user_w.shape[0], item_w.shape[0]])


# ---- a4: forward -----------------------------------------------------------------------------------
def gamma(all_users, all_items, users, items):
    """utility1/model.py:115-118."""
    return torch.sum(torch.mul(all_users[users], all_items[items]), dim=1)


def bce_forward(user_w, item_w, graph, n_layers, users, items, labels):
    """utility1/model.py:111-121 flag=0."""
    all_users, all_items = computer(user_w, item_w, graph, n_layers)
    return torch.nn.BCEWithLogitsLoss()(gamma(all_users, all_items, users, items), labels.float())


# ---- a5: north-star semantics, not reference-pinned ------------------------------------------------
def bpr_loss(user_w, item_w, graph, n_layers, users, pos, neg):
    """Upstream LightGCN bpr_loss: (mean softplus(neg-pos), 0.5*(|u0|^2+|p0|^2+|n0|^2)/B)."""
    all_users, all_items = computer(user_w, item_w, graph, n_layers)
    u, p, n = all_users[users], all_items[pos], all_items[neg]
    u0, p0, n0 = user_w[users], item_w[pos], item_w[neg]
    reg = 0.5 * (u0.norm(2).pow(2) + p0.norm(2).pow(2) + n0.norm(2).pow(2)) / float(len(users))
    pos_s = torch.sum(u * p, dim=1)
    neg_s = torch.sum(u * n, dim=1)
    return torch.mean(torch.nn.functional.softplus(neg_s - pos_s)), reg


def users_rating(all_users, all_items, users):
    """Upstream getUsersRating: sigmoid(U[users] . I^T)."""
    return torch.sigmoid(torch.matmul(all_users[users], all_items.t()))


def topk_masked(scores, train_rowptr, train_col, users, k):
    """Full-ranking top-k with train items excluded; order: score desc, ties by ascending id.
    Returns (idx int64 [B,k] with -1 padding, val [B,k] with -inf padding)."""
    s = scores.clone().double()
    B, m = s.shape
    for r, u in enumerate(users):
        s[r, torch.as_tensor(train_col[train_rowptr[u]: train_rowptr[u + 1]], dtype=torch.long)] = -np.inf
    order = np.lexsort((np.arange(m)[None].repeat(B, 0), -s.numpy()), axis=1)[:, :k]
    idx = torch.from_numpy(order.copy())
    val = torch.gather(s, 1, idx).float()
    idx[val == -np.inf] = -1
    return idx, val


# ---- a8: sampled evaluation ------------------------------------------------------------------------
def dcg_at_k(r, k):
    """utility1/metrics.py:43-58 (method 1)."""
    r = np.asarray(r, dtype=float)[:k]
    return np.sum(r / np.log2(np.arange(2, r.size + 2))) if r.size else 0.0


def ndcg_at_k(r, k):
    """utility1/metrics.py:61-71."""
    dcg_max = dcg_at_k(sorted(r, reverse=True), k)
    return dcg_at_k(r, k) / dcg_max if dcg_max else 0.0


def recall_at_k(r, k, all_pos_num):
    """utility1/metrics.py:88-94."""
    return 0.0 if all_pos_num == 0 else np.sum(np.asarray(r, dtype=float)[:k]) / all_pos_num


def test_sampled(all_users, all_items, testRatings, testNegatives):
    """utility1/batch_test.py:12-40,72-90 on already-propagated tables (eval mode is deterministic,
    so hoisting computer() out of the per-user loop does not change any value)."""
    result = {"recall": np.zeros(len(Ks)), "ndcg": np.zeros(len(Ks))}
    users = list(testRatings.keys())
    for u in users:
        pos = testRatings[u]
        test_items = testNegatives[u] + pos
        uu = torch.full((len(test_items),), u, dtype=torch.long)
        pred = gamma(all_users, all_items, uu, torch.tensor(test_items, dtype=torch.long)).tolist()
        rating = {}
        for i, item in enumerate(test_items):
            rating[item] = pred[i]
        top = heapq.nlargest(max(Ks), rating, key=rating.get)
        r = [1 if i in pos else 0 for i in top]
        result["recall"] += np.array([recall_at_k(r, k, len(pos)) for k in Ks]) / len(users)
        result["ndcg"] += np.array([ndcg_at_k(r, k) for k in Ks]) / len(users)
    return result


# ---- a9: expert gating -----------------------------------------------------------------------------
def expert_gate(e0, e_out, W):
    """utility1/model_expert_s.py:154-161: softmax([E0|Eout].W) convex mix."""
    att = torch.softmax(torch.matmul(torch.cat([e0, e_out], dim=1), W), dim=1)
    return e0 * att[:, 0:1] + e_out * att[:, 1:2]


# ---- a10: NGCF layer -------------------------------------------------------------------------------
def ngcf_layer(graph, ego, W1, b1, W2, b2, slope=0.2):
    """NGCF_SPEX/code/main_rec.py:76-82 without message dropout (eval): returns (ego', norm)."""
    side = torch.sparse.mm(graph, ego)
    lrelu = torch.nn.functional.leaky_relu
    out = lrelu(torch.nn.functional.linear(side, W1, b1), slope) + \
        lrelu(torch.nn.functional.linear(ego * side, W2, b2), slope)
    return out, torch.nn.functional.normalize(out, p=2, dim=1)


# ---- synthetic graphs for bench.py's CPU legs --------------------------------------------------------
def random_bipartite(n_users, m_items, n_inter, seed=2020):
    rng = np.random.default_rng(seed)
    u = np.concatenate([np.arange(n_users), rng.integers(0, n_users, max(n_inter - n_users, 0))])
    i = rng.integers(0, m_items, u.size)
    key = np.unique(u.astype(np.int64) * m_items + i)
    return (key // m_items).astype(np.int64), (key % m_items).astype(np.int64)
