#!/usr/bin/env python
"""Golden vectors on the reference's REAL dataset (BASELINE.json configs[0]: LightGCN_SPEX
main_rec.py on epinion2, 3 layers, dim 64), produced by the UNMODIFIED reference code.

Build container only (the reference does not travel to the GPU box):
    PYTHONHASHSEED=0 python tests/golden/make_epinion2.py
Stage 1 runs /root/reference/Data_process/rec/data_process_rec.py (change_format .. split, to_NGCF,
to_NCF, to_LightGCN) on copies of the shipped Data_process/rec/epinion2/*.mat inside a scratch tree.
Shims: random.sample(set, k) raises on py >= 3.11 (data_process_rec.py:249,270) -> sample from
sorted(set); PYTHONHASHSEED=0 + random.seed(2020) pin the set-order dependent re-indexing
(data_process_rec.py:143-148).  Result: 3 185 users x 12 407 items, 209 304 train interactions.
Stage 2 imports LightGCN_SPEX/code (shims as in make_golden.py) and records
  * the processed interactions / held-out positives / 99 negatives      -> epinion2_data.npz
  * Loader.getSparseGraph(): nnz, value sum, 4096 sampled entries        \
  * LightGCN.computer() K=3 on seeded weights: every 13th row, col sums   > epinion2_golden.npz
  * forward(flag=0) BCE loss + gradient rows, batch_test.test on 160 users /
Weights are not stored: both sides fill them from numpy default_rng(7) (see `seeded_weights`).
"""
import os
import random
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
N_TEST_USERS = 160


def seeded_weights(n_user_rows, m_items, D=64, seed=7):
    """U(-a, a) with the xavier bound of model.py:34-35, from numpy so that both sides agree."""
    rng = np.random.default_rng(seed)
    au, ai = np.sqrt(6.0 / (n_user_rows + D)), np.sqrt(6.0 / (m_items + D))
    U = rng.uniform(-au, au, (n_user_rows, D)).astype(np.float32)
    I = rng.uniform(-ai, ai, (m_items, D)).astype(np.float32)
    return U, I


def stage1(work):
    import importlib.util

    dp = os.path.join(work, "Data_process", "rec")
    os.makedirs(os.path.join(dp, "epinion2"))
    for f in ("rating_with_timestamp.mat", "trust_with_timestamp.mat"):
        shutil.copy(os.path.join(REF, "Data_process", "rec", "epinion2", f), os.path.join(dp, "epinion2", f))
    os.chdir(dp)
    orig = random.sample

    def sample(pop, k, **kw):
        if isinstance(pop, (set, frozenset)):
            pop = sorted(pop)
        return orig(pop, k, **kw)

    random.sample = sample
    random.seed(2020)
    sys.argv = ["data_process_rec.py", "--root", "epinion2"]
    spec = importlib.util.spec_from_file_location("dpr", os.path.join(REF, "Data_process/rec/data_process_rec.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.args = m.parse_args()
    for step in ("change_format", "select_items", "select_users", "unify_index", "sort", "split", "to_NGCF",
                 "to_NCF", "to_LightGCN"):
        getattr(m, step)()
    random.sample = orig
    return os.path.join(work, "LightGCN_SPEX", "data")


def main():
    assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0"
    work = os.environ.get("SPEX_EP2_WORK") or tempfile.mkdtemp(prefix="spex_ep2_")
    data_root = os.path.join(work, "LightGCN_SPEX", "data")
    if not os.path.exists(os.path.join(data_root, "epinion2", "rec", "epinion2.train.rating")):
        data_root = stage1(work)
    code = os.path.join(work, "LightGCN_SPEX", "code")
    os.makedirs(code, exist_ok=True)
    os.chdir(code)   # Loader reads "../data/<dataset>/" (dataloader.py:74)

    if not hasattr(np, "asfarray"):
        np.asfarray = lambda a, dtype=float: np.asarray(a, dtype=dtype)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.argv = ["main_rec.py", "--dataset", "epinion2", "--layer", "3", "--recdim", "64"]
    sys.path.insert(0, os.path.join(REF, "LightGCN_SPEX", "code"))
    import utility1.dataloader as ref_dl
    import utility1.model as ref_model
    import utility1.batch_test as ref_bt
    from lg_parser import parse_args_r

    args = parse_args_r()
    dataset = ref_dl.Loader(args)
    nu, m = dataset.n_users, dataset.m_items
    users = sorted(dataset.testRatings)
    data = {
        "train_user": dataset.trainUser.astype(np.int16), "train_item": dataset.trainItem.astype(np.int16),
        "test_user": np.array(users, dtype=np.int16),
        "test_pos": np.array([dataset.testRatings[u][0] for u in users], dtype=np.int16),
        "test_neg": np.array([dataset.testNegatives[u] for u in users], dtype=np.int16),
    }
    assert nu < 32768 and m < 32768
    np.savez_compressed(os.path.join(HERE, "epinion2_data.npz"), **data)

    G = {"n_users": nu, "m_items": m}
    A = dataset.getSparseGraph()
    idx, val = A.indices().numpy(), A.values().numpy()
    G["adj_nnz"] = idx.shape[1]
    G["adj_value_sum"] = np.float64(val.astype(np.float64).sum())
    pick = np.random.default_rng(1).choice(idx.shape[1], size=4096, replace=False)
    pick.sort()
    G["adj_pick"], G["adj_pick_rc"], G["adj_pick_val"] = pick, idx[:, pick], val[pick]
    G["adj_rowsum"] = np.asarray(np.bincount(idx[0], weights=val.astype(np.float64), minlength=A.shape[0]))[::13]

    model = ref_model.LightGCN(args, dataset)
    U, I = seeded_weights(nu + 1, m)
    with torch.no_grad():
        model.embedding_user.weight.copy_(torch.from_numpy(U))
        model.embedding_item.weight.copy_(torch.from_numpy(I))
    model.eval()
    with torch.no_grad():
        cu, ci = model.computer()
    allrows = torch.cat([cu, ci]).numpy()
    G["computer_rows_13"] = allrows[::13].copy()
    G["computer_colsum"] = allrows.astype(np.float64).sum(0)

    rng = np.random.default_rng(3)
    B = 256
    bu = rng.integers(0, nu, B)
    bi = rng.integers(0, m, B)
    bl = rng.integers(0, 2, B)
    G["batch_users"], G["batch_items"], G["batch_labels"] = bu, bi, bl
    model.train()
    model.zero_grad()
    loss = model(torch.from_numpy(bu), torch.from_numpy(bi), torch.from_numpy(bl), flag=0)
    loss.backward()
    G["bce_loss"] = np.float32(loss.item())
    gu, gi = model.embedding_user.weight.grad.numpy(), model.embedding_item.weight.grad.numpy()
    G["grad_user_rows_13"], G["grad_item_rows_13"] = gu[::13].copy(), gi[::13].copy()
    G["grad_user_abs_sum"], G["grad_item_abs_sum"] = np.float64(np.abs(gu).sum()), np.float64(np.abs(gi).sum())

    model.eval()
    sub = users[:N_TEST_USERS]
    with torch.no_grad():
        ret = ref_bt.test(model, {u: dataset.testRatings[u] for u in sub}, {u: dataset.testNegatives[u] for u in sub})
    G["test_users"] = np.array(sub)
    G["test_recall"], G["test_ndcg"] = np.asarray(ret["recall"]), np.asarray(ret["ndcg"])
    np.savez_compressed(os.path.join(HERE, "epinion2_golden.npz"), **G)
    if os.environ.get("SPEX_EP2_FULL"):
        # second fixture (epinion2_full.npz): EVERY row of computer() and of both gradients through two
        # float64 numbers per row (sum and sum of squares), and the reference's Test() over ALL test users
        # (one propagation per user in the reference: ~10 minutes on 8 CPU threads)
        F = {"computer_row_sum": allrows.astype(np.float64).sum(1), "computer_row_sq": (allrows.astype(np.float64) ** 2).sum(1),
             "grad_user_row_sum": gu.astype(np.float64).sum(1), "grad_user_row_sq": (gu.astype(np.float64) ** 2).sum(1),
             "grad_item_row_sum": gi.astype(np.float64).sum(1), "grad_item_row_sq": (gi.astype(np.float64) ** 2).sum(1)}
        with torch.no_grad():
            ret_all = ref_bt.test(model, dataset.testRatings, dataset.testNegatives)
        F["test_recall_all"], F["test_ndcg_all"] = np.asarray(ret_all["recall"]), np.asarray(ret_all["ndcg"])
        F["n_test_users"] = len(dataset.testRatings)
        np.savez_compressed(os.path.join(HERE, "epinion2_full.npz"), **F)
        print("epinion2 full: test users", F["n_test_users"], "recall", F["test_recall_all"], "ndcg", F["test_ndcg_all"])
    print("epinion2: users", nu, "items", m, "train", len(dataset.trainUser), "nnz", G["adj_nnz"],
          "loss %.6f" % G["bce_loss"], "recall", G["test_recall"], "ndcg", G["test_ndcg"])


if __name__ == "__main__":
    main()
