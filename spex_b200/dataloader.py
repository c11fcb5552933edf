"""Datasets for the LightGCN_SPEX path: same public surface as the reference's
utility1/dataloader.py (BasicDataset :10-63, Loader :65-236, LightTrainData :239-277), built
with vectorised numpy so that it also works at 10^9 interactions.

On-disk format (written by the reference's Data_process/rec/data_process_rec.py:401-416,516-519):
    <ds>.train.rating    "user item 1" per line, space separated (only the first two columns used)
    <ds>.test.rating     "user item ..." per line; the LAST line of a user wins (dataloader.py:139-148)
    <ds>.test.negative   "user n1 n2 ... n99" per line
    s_pre_adj_mat.npz    scipy CSR cache of the normalised adjacency (dataloader.py:191,215)

``getSparseGraph()`` keeps returning a coalesced fp32 torch sparse COO tensor of shape
(n_users+1+m_items)^2 (other code reads .indices()/.values()/.size(), model.py:47-49); the CSR the
kernels read is the additional ``getCSR()``.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np
import torch
from torch.utils.data import Dataset

from .graph import CSRGraph, build_interaction_csr, build_norm_adj, csr_from_coo, fold_rows


class BasicDataset(Dataset):
    """Abstract dataset: the extension point the reference model is written against."""

    def __init__(self):
        pass

    @property
    def n_users(self):
        raise NotImplementedError

    @property
    def m_items(self):
        raise NotImplementedError

    @property
    def trainDataSize(self):
        raise NotImplementedError

    @property
    def testDict(self):
        raise NotImplementedError

    @property
    def allPos(self):
        raise NotImplementedError

    def getUserItemFeedback(self, users, items):
        raise NotImplementedError

    def getUserPosItems(self, users):
        raise NotImplementedError

    def getUserNegItems(self, users):
        raise NotImplementedError

    def getSparseGraph(self):
        raise NotImplementedError


class _PairSet:
    """Membership test over the training pairs with the ``(u, i) in train_mat`` syntax of the
    reference's dok_matrix (dataloader.py:97-100,258), backed by a sorted int64 key array."""

    def __init__(self, users, items, m_items: int, shape):
        self.m_items = int(m_items)
        self.shape = shape
        self.keys = np.unique(np.asarray(users, np.int64) * self.m_items + np.asarray(items, np.int64))

    def __contains__(self, ui) -> bool:
        k = int(ui[0]) * self.m_items + int(ui[1])
        p = np.searchsorted(self.keys, k)
        return bool(p < self.keys.size and self.keys[p] == k)

    def contains(self, users: np.ndarray, items: np.ndarray) -> np.ndarray:
        k = np.asarray(users, np.int64) * self.m_items + np.asarray(items, np.int64)
        p = np.searchsorted(self.keys, k)
        p[p == self.keys.size] = 0
        return self.keys[p] == k if self.keys.size else np.zeros(k.shape, bool)

    def __len__(self):
        return int(self.keys.size)


class _GraphMixin:
    """Shared graph construction for file-backed and synthetic datasets."""

    Graph = None
    _csr: Optional[CSRGraph] = None
    _icsr = None
    split = 0
    folds = 1
    path: Optional[str] = None
    graph_device = "cpu"

    def getCSR(self) -> CSRGraph:
        """Normalised adjacency as host CSR (rowptr int64, col int32, val fp32) + transpose map."""
        if self._csr is None:
            cached = self._load_cached_csr()
            if cached is not None:
                self._csr = cached
            else:
                self._csr = build_norm_adj(self.trainUser, self.trainItem, self.n_users + 1, self.m_items)
                self._save_cached_csr(self._csr)
        return self._csr

    def getInteractionCSR(self):
        if self._icsr is None:
            self._icsr = build_interaction_csr(self.trainUser, self.trainItem, self.n_users + 1,
                                               self.m_items)
        return self._icsr

    def _cache_file(self):
        return None if self.path is None else os.path.join(self.path, "s_pre_adj_mat.npz")

    def _load_cached_csr(self) -> Optional[CSRGraph]:
        f = self._cache_file()
        if f is None or not os.path.exists(f):
            return None
        import scipy.sparse as sp

        try:
            m = sp.load_npz(f).tocsr().astype(np.float32)
        except Exception:
            return None
        N = self.n_users + 1 + self.m_items
        if m.shape != (N, N):
            return None
        m.sort_indices()
        g = CSRGraph(N, N, m.indptr.astype(np.int64), m.indices.astype(np.int32),
                     m.data.astype(np.float32))
        try:
            from .graph import transpose_positions

            g.tpos = transpose_positions(g)
        except ValueError:
            g.tpos = None
        return g

    def _save_cached_csr(self, g: CSRGraph):
        f = self._cache_file()
        if f is None:
            return
        try:
            import scipy.sparse as sp

            m = sp.csr_matrix((g.val, g.col, g.rowptr), shape=(g.n_rows, g.n_cols))
            sp.save_npz(f, m)
        except OSError:
            pass  # read-only data dir: the cache is an optimisation only

    def _coo_tensor(self, g: CSRGraph, r0: int = 0, r1: Optional[int] = None) -> torch.Tensor:
        r1 = g.n_rows if r1 is None else r1
        lo, hi = int(g.rowptr[r0]), int(g.rowptr[r1])
        rows = np.repeat(np.arange(r0, r1, dtype=np.int64), np.diff(g.rowptr[r0: r1 + 1])) - r0
        idx = torch.from_numpy(np.stack([rows, g.col[lo:hi].astype(np.int64)]))
        val = torch.from_numpy(np.ascontiguousarray(g.val[lo:hi]))
        t = torch.sparse_coo_tensor(idx, val, (r1 - r0, g.n_cols), is_coalesced=True)
        dev = torch.device(self.graph_device)
        return t.to(dev) if dev.type != "cpu" else t

    def getSparseGraph(self):
        """Coalesced fp32 sparse COO of D^-1/2 A D^-1/2 (list of row folds if A_split)."""
        if self.Graph is None:
            g = self.getCSR()
            if self.split:
                self.Graph = [self._coo_tensor(g, a, b) for a, b in fold_rows(g.n_rows, self.folds)]
            else:
                self.Graph = self._coo_tensor(g)
        return self.Graph


class Loader(_GraphMixin, BasicDataset):
    """File-backed dataset (epinion2 / weibo / twitter layout)."""

    def __init__(self, config, path: Optional[str] = None):
        super().__init__()
        dataset = config.dataset
        if path is None:
            base = getattr(config, "data_path", None) or "../data/"
            path = os.path.join(base, dataset) + "/"
        self.path = path
        self.split = getattr(config, "A_split", 0)
        self.folds = getattr(config, "a_fold", 100)
        self.mode_dict = {"train": 0, "test": 1}
        self.mode = self.mode_dict["train"]
        self.graph_device = "cuda" if torch.cuda.is_available() else "cpu"
        train_file = os.path.join(path, "rec", f"{dataset}.train.rating")
        test_rating_file = os.path.join(path, "rec", f"{dataset}.test.rating")
        test_negative_file = os.path.join(path, "rec", f"{dataset}.test.negative")

        pairs = _read_int_columns(train_file, 2)
        self.trainUser = pairs[:, 0].astype(np.int32)
        self.trainItem = pairs[:, 1].astype(np.int32)
        self.n_user = int(self.trainUser.max()) + 1
        self.m_item = int(self.trainItem.max()) + 1
        self.trainUniqueUsers = np.unique(self.trainUser)
        self.traindataSize = 0  # the reference never fills these two (dataloader.py:86-87)
        self.testDataSize = 0
        self.rec_train_data = pairs[:, :2].tolist()
        self.train_mat = _PairSet(self.trainUser, self.trainItem, self.m_item,
                                  (self.n_user + 1, self.m_item))
        self.testRatings = self.load_test_rating_as_dict(test_rating_file)
        self.testNegatives = self.load_test_negative_as_dict(test_negative_file)
        self.Graph = None
        print(dataset)
        print("use:", self.n_user)
        print("item:", self.m_item)
        print("----------------")
        rp, col = self.getInteractionCSR()
        deg_u = np.diff(rp).astype(np.float64)
        self.users_D = np.where(deg_u == 0, 1.0, deg_u)
        deg_i = np.bincount(col, minlength=self.m_item).astype(np.float64)
        self.items_D = np.where(deg_i == 0, 1.0, deg_i)

    @property
    def n_users(self):
        return self.n_user

    @property
    def m_items(self):
        return self.m_item

    @property
    def trainDataSize(self):
        return self.traindataSize

    @staticmethod
    def load_test_rating_as_dict(filename) -> Dict[int, List[int]]:
        arr = _read_int_columns(filename, 2)
        return {int(u): [int(i)] for u, i in arr[:, :2]}  # later lines overwrite earlier ones

    @staticmethod
    def load_test_negative_as_dict(filename) -> Dict[int, List[int]]:
        out = {}
        with open(filename, "r") as f:
            for line in f:
                parts = line.split()
                if parts:
                    out[int(parts[0])] = [int(x) for x in parts[1:]]
        return out

    def getUserItemFeedback(self, users, items):
        return self.train_mat.contains(np.asarray(users), np.asarray(items)).astype("uint8").reshape(-1)

    def getUserPosItems(self, users):
        rp, col = self.getInteractionCSR()
        return [col[rp[u]: rp[u + 1]] for u in users]


def _read_int_columns(filename: str, min_cols: int) -> np.ndarray:
    """Whitespace-separated integer table -> int64 [rows, >=min_cols] (blank lines skipped)."""
    import pandas as pd

    df = pd.read_csv(filename, sep=r"\s+", header=None, usecols=list(range(min_cols)),
                     dtype=np.int64, engine="c", skip_blank_lines=True)
    return df.to_numpy()


class LightTrainData(Dataset):
    """1 positive + num_ng sampled negatives per training pair (dataloader.py:239-277).

    ``ng_sample()`` is vectorised: all negatives are drawn at once and only collisions with the
    training set are redrawn.  With ``exact_stream=True`` it consumes ``np.random`` draws in the
    reference's order (one draw per slot, redraw immediately on collision), so a shared
    ``np.random.seed`` yields the same negatives as the reference's Python loop.
    """

    def __init__(self, features, num_item, train_mat=None, num_ng: int = 5, exact_stream: bool = True):
        super().__init__()
        self.features_ps = features
        self.num_item = int(num_item)
        self.train_mat = train_mat
        self.num_ng = num_ng
        self.exact_stream = exact_stream
        self.labels = [0 for _ in range(len(features))]
        self._ps = np.asarray(features, dtype=np.int64).reshape(-1, 2) if len(features) else np.zeros((0, 2), np.int64)
        self._users = self._items = self._labels = None

    def _contains(self, users, items):
        tm = self.train_mat
        if hasattr(tm, "contains"):
            return tm.contains(users, items)
        return np.fromiter(((int(u), int(j)) in tm for u, j in zip(users, items)), bool, len(users))

    def ng_sample_device(self, graph, n_user_rows: int, seed: int = 2020):
        """ng_sample() on the GPU: the epoch's (users, items, labels) as device tensors, negatives from
        spex_sample_negatives against the device adjacency `graph` (same distribution as ng_sample, not the
        same np.random stream - the exact-stream CPU path stays the parity oracle).  Returns the tensors
        and also keeps them for arrays_device()."""
        import torch

        dev = graph.device
        ps = torch.from_numpy(self._ps).to(dev)
        neg = sample_negatives_device(graph.rowptr, graph.col, n_user_rows, self.num_item, ps[:, 0], self.num_ng,
                                      seed)
        users = torch.cat([ps[:, 0], ps[:, 0].repeat_interleave(self.num_ng)])
        items = torch.cat([ps[:, 1], neg.reshape(-1)])
        labels = torch.cat([torch.ones(ps.shape[0], device=dev), torch.zeros(neg.numel(), device=dev)])
        self._device_arrays = (users, items, labels)
        return self._device_arrays

    def ng_sample(self):
        n_pos = self._ps.shape[0]
        slots_u = np.repeat(self._ps[:, 0], self.num_ng)
        n = slots_u.size
        neg = np.empty(n, dtype=np.int64)
        if self.exact_stream:
            _sample_stream_exact(slots_u, neg, self.num_item, self._contains)
        else:
            todo = np.arange(n)
            while todo.size:
                draw = np.random.randint(self.num_item, size=todo.size)
                neg[todo] = draw
                bad = self._contains(slots_u[todo], draw)
                todo = todo[bad]
        self._users = np.concatenate([self._ps[:, 0], slots_u])
        self._items = np.concatenate([self._ps[:, 1], neg])
        self._labels = np.concatenate([np.ones(n_pos, np.int64), np.zeros(n, np.int64)])
        self.features_ng = np.stack([slots_u, neg], 1).tolist() if n < 5_000_000 else None

    # reference attribute names, materialised lazily
    @property
    def features_fill(self):
        return np.stack([self._users, self._items], 1).tolist()

    @property
    def labels_fill(self):
        return self._labels.tolist()

    def arrays(self):
        """(users, items, labels) int64 arrays of the current epoch (batched loaders use these)."""
        return self._users, self._items, self._labels

    def __len__(self):
        return (self.num_ng + 1) * len(self.labels)

    def __getitem__(self, idx):
        return int(self._users[idx]), int(self._items[idx]), int(self._labels[idx])


def _sample_stream_exact(slots_u, neg, num_item, contains, chunk: int = 65536, window: int = 1024):
    """Sequential-equivalent rejection sampling over the legacy np.random stream.

    Reference loop (dataloader.py:253-260): for every slot in order, draw np.random.randint(num_item)
    until the pair is not a training pair.  In stream terms: slot k takes the first acceptable draw
    at or after its start position, slot k+1 starts right after it.  np.random.randint(n, size=m)
    yields the same m values as m scalar calls, and a chunk never holds more draws than there are
    open slots, so the generator ends in exactly the state the reference loop leaves it in.
    """
    n = slots_u.size
    s = 0  # next open slot
    while s < n:
        m = min(chunk, n - s)
        d = np.random.randint(num_item, size=m)
        p = 0  # next unconsumed draw of this chunk
        while p < m:
            take = min(m - p, window)
            bad = contains(slots_u[s: s + take], d[p: p + take])
            if not bad.any():
                neg[s: s + take] = d[p: p + take]
                s += take
                p += take
                continue
            f = int(np.argmax(bad))  # first collision: slots before it are settled
            neg[s: s + f] = d[p: p + f]
            s += f
            p += f + 1  # the colliding draw is consumed; slot s retries with the next draw


def uniform_sample_bpr(train_users, train_items, n_users: int, m_items: int, n_samples: Optional[int] = None,
                       seed: int = 2020) -> np.ndarray:
    """(user, positive, negative) triples for ``LightGCN.bpr_loss`` - int64 [n, 3].

    The reference has no BPR path (SURVEY "five facts" #2); the semantics are those of the upstream
    LightGCN-PyTorch sampler the reference was derived from (README.md:29): draw ``n_samples``
    (default: the number of training pairs) users uniformly WITH replacement, skip users without
    interactions, take one of the user's training items uniformly as the positive and rejection-
    sample a negative that is not a training item of the user.  Vectorised: positives come from the
    interaction CSR, negatives are drawn for all triples at once and only collisions are redrawn.
    North-star semantics, not reference-pinned; deterministic in ``seed``.
    """
    tu = np.asarray(train_users, dtype=np.int64).ravel()
    ti = np.asarray(train_items, dtype=np.int64).ravel()
    rng = np.random.default_rng(seed)
    n = tu.size if n_samples is None else int(n_samples)
    key = np.unique(tu * m_items + ti)                   # sorted (user, item) pairs, duplicates dropped
    ku, ki = key // m_items, key % m_items
    deg = np.bincount(ku, minlength=n_users)
    rp = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(deg, out=rp[1:])
    users = rng.integers(0, n_users, n)
    users = users[deg[users] > 0]
    pos = ki[rp[users] + (rng.random(users.size) * deg[users]).astype(np.int64)]
    neg = np.empty(users.size, dtype=np.int64)
    todo = np.arange(users.size)
    while todo.size:
        draw = rng.integers(0, m_items, todo.size)
        neg[todo] = draw
        k = users[todo] * m_items + draw
        p = np.searchsorted(key, k)
        p[p == key.size] = 0
        todo = todo[key[p] == k] if key.size else todo[:0]
    return np.stack([users, pos, neg], axis=1)


def sample_negatives_device(rowptr, col, n_user_rows: int, m_items: int, users, n_neg: int = 5, seed: int = 2020):
    """GPU counterpart of LightTrainData.ng_sample's rejection loop (dataloader.py:250-265): for every
    entry of `users` (int64, device) draw `n_neg` items uniformly among the items the user has NOT
    interacted with -> int64 [n, n_neg].  (rowptr, col) is the device adjacency CSR whose user rows list
    items as n_user_rows + item (spex_b200.ops.DeviceGraph; the hot flag in bit 31 is masked).
    Counter-based randomness: deterministic in `seed`, independent of thread scheduling."""
    import torch

    from ._capi import call, ptr, stream_ptr

    if not users.is_cuda:
        raise RuntimeError("sample_negatives_device runs on sm_100a only (the CPU sampler is LightTrainData)")
    users = users.to(torch.int64).contiguous()
    out = torch.empty(users.numel(), n_neg, dtype=torch.int64, device=users.device)
    call("spex_sample_negatives", ptr(rowptr), ptr(col), int(n_user_rows), int(m_items), ptr(users), users.numel(),
         int(n_neg), int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(out), stream_ptr())
    return out


def uniform_sample_bpr_device(rowptr, col, n_user_rows: int, m_items: int, n_users: int, n_samples: int,
                              seed: int = 2020):
    """GPU counterpart of uniform_sample_bpr: (users, pos, neg) int64 [n] each, on the device."""
    import torch

    from ._capi import call, ptr, stream_ptr

    dev = rowptr.device
    users = torch.empty(n_samples, dtype=torch.int64, device=dev)
    pos = torch.empty_like(users)
    neg = torch.empty_like(users)
    call("spex_sample_bpr", ptr(rowptr), ptr(col), int(n_user_rows), int(m_items), int(n_users), int(n_samples),
         int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(users), ptr(pos), ptr(neg), stream_ptr())
    return users, pos, neg


class SyntheticDataset(_GraphMixin, BasicDataset):
    """Seeded synthetic bipartite interactions with the dataset surface the model and the
    evaluation harness read (n_users, m_items, trainUser/trainItem, testRatings, testNegatives).

    Users have uniform activity, items Zipf(alpha) popularity, duplicates removed, every user has
    at least one interaction, ids randomly permuted.  One interaction per user is
    held out as the test positive when ``with_test``.
    """

    def __init__(self, n_users: int, m_items: int, n_interactions: int, seed: int = 2020,
                 zipf_alpha: float = 1.3, with_test: bool = True, n_test_neg: int = 99,
                 A_split: int = 0, a_fold: int = 1):
        super().__init__()
        rng = np.random.default_rng(seed)
        self.n_user, self.m_item = int(n_users), int(m_items)
        self.split, self.folds = A_split, a_fold
        self.path = None
        self.graph_device = "cpu"
        # item popularity: Zipf over a random permutation of item ids
        ranks = np.arange(1, m_items + 1, dtype=np.float64)
        p = ranks ** (-zipf_alpha)
        p /= p.sum()
        item_perm = rng.permutation(m_items)
        u = np.concatenate([np.arange(n_users), rng.integers(0, n_users, max(n_interactions - n_users, 0))])
        it = item_perm[rng.choice(m_items, size=u.size, p=p)]
        key = np.unique(u.astype(np.int64) * m_items + it)
        u = (key // m_items).astype(np.int32)
        it = (key % m_items).astype(np.int32)
        self.testRatings, self.testNegatives = {}, {}
        if with_test:
            # hold out the last interaction of every user with >= 2
            rp = np.zeros(n_users + 1, np.int64)
            np.cumsum(np.bincount(u, minlength=n_users), out=rp[1:])
            deg = np.diff(rp)
            hold = rp[1:][deg >= 2] - 1
            keep = np.ones(u.size, bool)
            keep[hold] = False
            pset = _PairSet(u, it, m_items, (n_users + 1, m_items))
            for h in hold:
                uu = int(u[h])
                self.testRatings[uu] = [int(it[h])]
                negs = []
                while len(negs) < n_test_neg:
                    c = rng.integers(0, m_items, n_test_neg)
                    ok = ~pset.contains(np.full(c.size, uu), c)
                    negs.extend(int(x) for x in c[ok][: n_test_neg - len(negs)])
                self.testNegatives[uu] = negs
            u, it = u[keep], it[keep]
        self.trainUser, self.trainItem = u, it
        self.trainUniqueUsers = np.unique(u)
        self.train_mat = _PairSet(u, it, m_items, (n_users + 1, m_items))
        self.Graph = None

    @property
    def rec_train_data(self):
        return np.stack([self.trainUser, self.trainItem], 1).tolist()

    @property
    def n_users(self):
        return self.n_user

    @property
    def m_items(self):
        return self.m_item

    @property
    def trainDataSize(self):
        return int(self.trainUser.size)
