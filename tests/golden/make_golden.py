#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference code (imported from /root/reference)
on a tiny dataset written in the reference's on-disk format.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Outputs (committed): tests/golden/tiny/rec/tiny.{train.rating,test.rating,test.negative}
                     tests/golden/lightgcn_tiny.npz

Shims installed here (never by editing the reference): np.asfarray (removed in numpy 2, used at
utility1/metrics.py:50,75), Tensor.cuda -> identity on this GPU-less host (dataloader.py:176,222),
sys.argv set before import (argparse at import time, utility1/batch_test.py:5-6), cwd such that
"../data/<dataset>/" resolves (dataloader.py:74).
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/LightGCN_SPEX/code"
DS = "tiny"


def write_dataset(root):
    rng = np.random.default_rng(2020)
    n_users, m_items = 60, 150
    rec = os.path.join(root, "rec")
    os.makedirs(rec, exist_ok=True)
    pairs = set()
    for u in range(n_users):
        deg = int(rng.integers(3, 20))
        for i in rng.choice(m_items, size=deg, replace=False):
            pairs.add((u, int(i)))
    # a few very popular items so degrees are skewed
    for u in range(0, n_users, 2):
        pairs.add((u, 0))
        pairs.add((u, 149))
    pairs = sorted(pairs)
    per_user = {}
    for u, i in pairs:
        per_user.setdefault(u, []).append(i)
    train, test = [], {}
    for u, items in per_user.items():
        held = items[int(rng.integers(len(items)))]
        test[u] = held
        train += [(u, i) for i in items if i != held]
    order = rng.permutation(len(train))  # file order is not sorted in the real data either
    with open(os.path.join(rec, f"{DS}.train.rating"), "w") as f:
        for k in order:
            f.write(f"{train[k][0]} {train[k][1]} 1\n")
    with open(os.path.join(rec, f"{DS}.test.rating"), "w") as f:
        for u in sorted(test):
            f.write(f"{u} {test[u]} 1\n")
    with open(os.path.join(rec, f"{DS}.test.negative"), "w") as f:
        for u in sorted(test):
            seen = set(per_user[u])
            cand = [i for i in range(m_items) if i not in seen]
            negs = rng.choice(cand, size=99, replace=False)
            f.write(" ".join([str(u)] + [str(int(x)) for x in negs]) + "\n")


def write_trust(root, n_users):
    """Synthetic trust paths in the pickle layout main_11.py:31-32 reads:
    train.txt = (paths, targets); test2.txt = (paths, targets, candidates with the positive LAST)."""
    import pickle

    rng = np.random.default_rng(99)
    tdir = os.path.join(root, "trust")
    os.makedirs(tdir, exist_ok=True)

    def paths(n_per_user):
        P, T = [], []
        for u in range(n_users):
            for _ in range(n_per_user):
                L = int(rng.integers(1, 6))
                p = [u] + [int(x) for x in rng.integers(0, n_users, L - 1)]
                P.append(p)
                T.append(int(rng.integers(0, n_users)))
        return P, T

    P, T = paths(3)
    pickle.dump((P, T), open(os.path.join(tdir, "train.txt"), "wb"))
    P2, T2 = paths(1)
    negs = []
    for t in T2:
        cand = [x for x in range(n_users) if x != t]
        negs.append([int(x) for x in rng.choice(cand, size=49, replace=False)] + [t])
    pickle.dump((P2, T2, negs), open(os.path.join(tdir, "test2.txt"), "wb"))


def make_expert_golden(args, dataset, work):
    """Multi-task model (utility1/model_expert_s.py) on the tiny dataset + synthetic trust paths."""
    import pickle
    import utility1.model_expert_s as ref_me
    import utility1.batch_test as ref_bt
    import utility1.utils as ref_utils
    from utility2.utils import Data
    from utility2.batch_test_gnn import trust_test5

    tdir = os.path.join(HERE, DS, "trust")
    train_raw = pickle.load(open(os.path.join(tdir, "train.txt"), "rb"))
    test_raw = pickle.load(open(os.path.join(tdir, "test2.txt"), "rb"))
    train_paths = Data(train_raw, dataset.n_users, shuffle=False)
    test_paths = Data(test_raw, dataset.n_users, shuffle=False, test=True)
    args.batchSize = 32
    ref_utils.set_seed(2020)
    model = ref_me.LightGCN(args, dataset)
    E = {}
    for k, v in model.state_dict().items():
        E["sd." + k] = v.detach().numpy().copy()
    rng = np.random.default_rng(11)
    B = 64
    users = rng.integers(0, dataset.n_users, B)
    items = rng.integers(0, dataset.m_items, B)
    labels = rng.integers(0, 2, B)
    sl = rng.choice(len(train_raw[0]), size=40, replace=False)
    E["users"], E["items"], E["labels"], E["slice"] = users, items, labels, sl
    model.train()
    model.zero_grad()
    l1, l2 = model(torch.from_numpy(users), torch.from_numpy(items), torch.from_numpy(labels), sl, train_paths, flag=0)
    (l1 + l2).backward()
    E["loss1"], E["loss2"] = np.float32(l1.item()), np.float32(l2.item())
    for name, p in model.named_parameters():
        if p.grad is not None:
            E["grad." + name] = p.grad.numpy().copy()
    model.eval()
    with torch.no_grad():
        E["gamma"] = model(torch.from_numpy(users), torch.from_numpy(items), None, None, None, flag=1).numpy().copy()
        sc, negs = model(None, None, None, np.arange(10), test_paths, 2)
        E["trust_scores"] = sc.numpy().copy()
        ret = ref_bt.rec_test(model, dataset.testRatings, dataset.testNegatives)
        E["rec_recall"], E["rec_ndcg"] = ret["recall"], ret["ndcg"]
        E["trust_metrics"] = np.array(trust_test5(model, test_paths))
    np.savez_compressed(os.path.join(HERE, "expert_tiny.npz"), **E)
    print("wrote expert_tiny.npz: loss1 %.6f loss2 %.6f trust %s" % (E["loss1"], E["loss2"], E["trust_metrics"]))


def main():
    out_ds = os.path.join(HERE, DS)
    if os.path.exists(out_ds):
        shutil.rmtree(out_ds)
    write_dataset(out_ds)
    write_trust(out_ds, 60)

    work = tempfile.mkdtemp(prefix="spex_golden_")
    os.makedirs(os.path.join(work, "code"))
    shutil.copytree(out_ds, os.path.join(work, "data", DS))
    os.chdir(os.path.join(work, "code"))

    if not hasattr(np, "asfarray"):
        np.asfarray = lambda a, dtype=float: np.asarray(a, dtype=dtype)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.argv = ["main_rec.py", "--dataset", DS, "--layer", "3", "--recdim", "64"]
    sys.path.insert(0, REF)
    import utility1.dataloader as ref_dl
    import utility1.model as ref_model
    import utility1.utils as ref_utils
    import utility1.batch_test as ref_bt
    from lg_parser import parse_args_r

    G = {}
    args = parse_args_r()
    ref_utils.set_seed(2020)
    dataset = ref_dl.Loader(args)
    A = dataset.getSparseGraph()
    G["adj_indices"] = A.indices().numpy()
    G["adj_values"] = A.values().numpy()
    G["n_users"], G["m_items"] = dataset.n_users, dataset.m_items

    ref_utils.set_seed(2020)
    model = ref_model.LightGCN(args, dataset)
    G["user_w"] = model.embedding_user.weight.detach().numpy().copy()
    G["item_w"] = model.embedding_item.weight.detach().numpy().copy()

    # computer() for K = 0..4 (eval, no dropout)
    model.eval()
    for K in range(5):
        model.n_layers = K
        with torch.no_grad():
            u, i = model.computer()
        G[f"computer_users_K{K}"] = u.numpy().copy()
        G[f"computer_items_K{K}"] = i.numpy().copy()
    model.n_layers = 3

    # A_split: serial row folds give the same result (dataloader.py:167-177, model.py:84-89)
    args.A_split, args.a_fold = 1, 4
    ds_split = ref_dl.Loader(args)
    m_split = ref_model.LightGCN(args, ds_split)
    m_split.load_state_dict(model.state_dict())
    m_split.eval()
    with torch.no_grad():
        u, i = m_split.computer()
    G["split_users_K3"] = u.numpy().copy()
    G["split_items_K3"] = i.numpy().copy()
    args.A_split = 0

    # edge dropout in training mode: CPU torch.rand(nnz) under a fixed seed (model.py:46-55)
    args.dropout = 1
    model.train()
    torch.manual_seed(123)
    G["dropout_rand"] = torch.rand(A._nnz()).numpy()
    torch.manual_seed(123)
    u, i = model.computer()
    G["dropout_users_K3"] = u.detach().numpy().copy()
    G["dropout_items_K3"] = i.detach().numpy().copy()
    G["dropout_keepprob"] = args.keepprob
    args.dropout = 0

    # forward(flag=0): BCE loss + gradients; batch with duplicate users and items
    rng = np.random.default_rng(7)
    B = 96
    users = rng.integers(0, dataset.n_users, B)
    users[:20] = users[0]
    items = rng.integers(0, dataset.m_items, B)
    items[40:55] = items[40]
    labels = rng.integers(0, 2, B)
    G["batch_users"], G["batch_items"], G["batch_labels"] = users, items, labels
    model.zero_grad()
    loss = model(torch.from_numpy(users), torch.from_numpy(items), torch.from_numpy(labels), flag=0)
    loss.backward()
    G["bce_loss"] = np.float32(loss.item())
    G["bce_grad_user"] = model.embedding_user.weight.grad.numpy().copy()
    G["bce_grad_item"] = model.embedding_item.weight.grad.numpy().copy()
    model.eval()
    with torch.no_grad():
        G["gamma"] = model(torch.from_numpy(users), torch.from_numpy(items), None, flag=1).numpy().copy()

    # Test(): sampled ranking metrics through the reference's own batch_test.test
    with torch.no_grad():
        ret = ref_bt.test(model, dataset.testRatings, dataset.testNegatives)
    G["test_recall"], G["test_ndcg"] = ret["recall"], ret["ndcg"]

    # three training steps with Adam (main_rec.py:23,30-37), fixed batches
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=args.lr)
    losses = []
    for step in range(3):
        opt.zero_grad()
        sl = slice(step * 32, step * 32 + 32)
        l = model(torch.from_numpy(users[sl]), torch.from_numpy(items[sl]), torch.from_numpy(labels[sl]), flag=0)
        l.backward()
        opt.step()
        losses.append(l.item())
    G["adam_losses"] = np.array(losses, dtype=np.float32)
    G["adam_user_w"] = model.embedding_user.weight.detach().numpy().copy()
    G["adam_item_w"] = model.embedding_item.weight.detach().numpy().copy()
    model.eval()
    with torch.no_grad():
        ret = ref_bt.test(model, dataset.testRatings, dataset.testNegatives)
    G["test_recall_after"], G["test_ndcg_after"] = ret["recall"], ret["ndcg"]

    # expert gating of the multi-task model (model_expert_s.py:154-161): formula on reference tensors
    torch.manual_seed(5)
    W = torch.empty(128, 2)
    torch.nn.init.xavier_uniform_(W, gain=1)
    e0 = torch.from_numpy(G["user_w"])
    ex = torch.from_numpy(G["computer_users_K3"])
    att = torch.softmax(torch.matmul(torch.cat([e0, ex], 1), W), 1)
    G["gate_W"] = W.numpy()
    G["gate_out"] = (torch.mul(e0, att[:, 0].unsqueeze(1)) + torch.mul(ex, att[:, 1].unsqueeze(1))).numpy()

    np.savez_compressed(os.path.join(HERE, "lightgcn_tiny.npz"), **G)
    make_expert_golden(args, dataset, work)
    shutil.rmtree(work)
    print("wrote", os.path.join(HERE, "lightgcn_tiny.npz"), "keys:", len(G))
    print("test_recall", G["test_recall"], "bce_loss", G["bce_loss"], "nnz", A._nnz())


if __name__ == "__main__":
    main()
