#!/usr/bin/env python
"""Golden vectors for BASELINE.json configs[2] (NGCF_SPEX propagation) from the UNMODIFIED reference:
/root/reference/NGCF_SPEX/code/main_rec.py::Model_Wrapper + utility/load_data.py::Data on the
epinion2 data written by tests/golden/make_epinion2.py's stage 1 (to_NGCF).

    SPEX_EP2_WORK=<scratch tree of make_epinion2.py> PYTHONHASHSEED=0 python tests/golden/make_ngcf.py

Shims: random.sample(set, k) (load_data.py:20 on py >= 3.11), Tensor.cuda/Module.cuda -> identity on
this GPU-less host, sys.argv + cwd before importing main_rec.py (argparse, Data() and the log
directory are created at import time: main_rec.py:2-3,29-33, utility/batch_test.py:9-13).
Weights are not stored: both sides fill them from numpy default_rng(11) (`seeded_ngcf_weights`).
Output: tests/golden/ngcf_epinion2.npz
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/NGCF_SPEX/code"


def seeded_ngcf_weights(n_user_rows, n_items, D=64, seed=11):
    rng = np.random.default_rng(seed)
    au, ai, aw = np.sqrt(6.0 / (n_user_rows + D)), np.sqrt(6.0 / (n_items + D)), 1.0 / np.sqrt(D)
    return {
        "user": rng.uniform(-au, au, (n_user_rows, D)).astype(np.float32),
        "item": rng.uniform(-ai, ai, (n_items, D)).astype(np.float32),
        "W1": rng.uniform(-aw, aw, (D, D)).astype(np.float32), "b1": rng.uniform(-aw, aw, D).astype(np.float32),
        "W2": rng.uniform(-aw, aw, (D, D)).astype(np.float32), "b2": rng.uniform(-aw, aw, D).astype(np.float32),
    }


def main():
    work = os.environ["SPEX_EP2_WORK"]
    code = os.path.join(work, "NGCF_SPEX", "code")
    os.makedirs(code, exist_ok=True)
    os.chdir(code)
    orig = random.sample
    random.sample = lambda pop, k, **kw: orig(sorted(pop) if isinstance(pop, (set, frozenset)) else pop, k, **kw)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.argv = ["main_rec.py", "--dataset", "epinion2", "--data_path", "../data/"]
    sys.path.insert(0, REF)
    import main_rec as ref   # noqa: E402  (runs the reference's import-time set-up)

    dg = ref.data_generator
    nu, ni = dg.n_users, dg.n_items
    _, norm_adj, _ = dg.create_adj_mat()
    config = {"n_users": nu, "n_items": ni, "norm_adj": norm_adj}
    model = ref.Model_Wrapper(data_config=config, device=torch.device("cpu"))
    W = seeded_ngcf_weights(nu + 1, ni)
    with torch.no_grad():
        model.user_embedding.weight.copy_(torch.from_numpy(W["user"]))
        model.item_embedding.weight.copy_(torch.from_numpy(W["item"]))
        model.GC_Linear_list[0].weight.copy_(torch.from_numpy(W["W1"]))
        model.GC_Linear_list[0].bias.copy_(torch.from_numpy(W["b1"]))
        model.Bi_Linear_list[0].weight.copy_(torch.from_numpy(W["W2"]))
        model.Bi_Linear_list[0].bias.copy_(torch.from_numpy(W["b2"]))
    G = {"n_users": nu, "n_items": ni, "n_layers": model.n_layers}
    coo = norm_adj.tocoo().astype(np.float32)
    order = np.lexsort((coo.col, coo.row))
    r, c, v = coo.row[order], coo.col[order], coo.data[order]
    G["adj_nnz"] = r.size
    pick = np.sort(np.random.default_rng(2).choice(r.size, size=4096, replace=False))
    G["adj_pick"], G["adj_pick_rc"], G["adj_pick_val"] = pick, np.stack([r[pick], c[pick]]), v[pick]
    G["adj_value_sum"] = np.float64(v.astype(np.float64).sum())
    # the interactions must be the ones LightGCN's files hold (same ids): R block of the adjacency
    R = dg.R.tocoo()
    G["train_pairs_checksum"] = np.int64((R.row.astype(np.int64) * 1000003 + R.col.astype(np.int64)).sum())
    G["n_train"] = R.row.size

    model.eval()
    with torch.no_grad():
        ua, ia = model(None, None, None, 1)
    rows = torch.cat([ua, ia]).numpy()
    G["out_rows_29"] = rows[::29].copy()
    G["out_colsum"] = rows.astype(np.float64).sum(0)
    rng = np.random.default_rng(5)
    B = 256
    bu, bi = rng.integers(0, nu, B), rng.integers(0, ni, B)
    bl = rng.integers(0, 2, B).astype(np.float32)
    G["batch_users"], G["batch_items"], G["batch_labels"] = bu, bi, bl
    model.zero_grad()
    loss = model(torch.from_numpy(bu), torch.from_numpy(bi), torch.from_numpy(bl), 0)   # eval: no dropout
    loss.backward()
    G["bce_loss"] = np.float32(loss.item())
    G["grad_W1"] = model.GC_Linear_list[0].weight.grad.numpy().copy()
    G["grad_W2"] = model.Bi_Linear_list[0].weight.grad.numpy().copy()
    G["grad_b1"] = model.GC_Linear_list[0].bias.grad.numpy().copy()
    G["grad_b2"] = model.Bi_Linear_list[0].bias.grad.numpy().copy()
    G["grad_user_rows_29"] = model.user_embedding.weight.grad.numpy()[::29].copy()
    G["grad_item_rows_29"] = model.item_embedding.weight.grad.numpy()[::29].copy()
    np.savez_compressed(os.path.join(HERE, "ngcf_epinion2.npz"), **G)
    print("ngcf: users", nu, "items", ni, "nnz", G["adj_nnz"], "layers", model.n_layers, "out", rows.shape,
          "loss %.6f" % G["bce_loss"])


if __name__ == "__main__":
    main()
