"""Rec-only training / evaluation entry point: `python -m spex_b200.main_rec --dataset epinion2 ...`

Same flow, same printed lines as /root/reference/LightGCN_SPEX/code/main_rec.py:
    per epoch: Train() prints '%d,%.5f' % (epoch, total_loss)                     (main_rec.py:38)
               Test()  prints 'Rec:  Epoch %d : recall=[...],  ndcg=[...]'        (main_rec.py:51-54)
    at the end '--- Train Best ---' and the best line                             (main_rec.py:72-74)
Differences in HOW: batches are sliced from the epoch's arrays with one torch.randperm (the same
draw DataLoader(shuffle=True)'s RandomSampler makes, so a shared seed visits the same batches), the
loss is accumulated on the device and read once per epoch instead of `.item()` per step
(main_rec.py:36), Test() propagates once instead of once per user (batch_test.py:33), and the
optimiser is the fused dense Adam.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import batch_test, dataloader, utils
from .lg_parser import parse_args_r
from .model import LightGCN
from .optim import FusedAdam

BATCH = 256  # hard-coded in the reference (main_rec.py:20)


def epoch_batches(n: int, batch: int, generator=None):
    """Index batches in the order DataLoader(shuffle=True) would produce them, consuming the default
    generator identically: the DataLoader iterator draws its base seed (one int64), then the
    RandomSampler draws its own seed (one int64) and shuffles with torch.randperm(n, seeded)."""
    torch.empty((), dtype=torch.int64).random_(generator=generator)  # DataLoader's _base_seed
    seed = int(torch.empty((), dtype=torch.int64).random_(generator=generator).item())
    g = torch.Generator()
    g.manual_seed(seed)
    perm = torch.randperm(n, generator=g)
    return [perm[i: i + batch] for i in range(0, n, batch)]


def Train(train_dataset, recommend_model, epoch, optimizer, device, batch=BATCH):
    train_dataset.ng_sample()
    recommend_model.train()
    users, items, labels = (torch.from_numpy(a) for a in train_dataset.arrays())
    # the epoch's samples live on the device; each step slices its batch there
    users_d, items_d = users.to(device), items.to(device)
    labels_d = labels.to(device=device, dtype=torch.float32)
    total = torch.zeros((), dtype=torch.float64, device=device)
    for idx in epoch_batches(users.numel(), batch):
        idx = idx.to(device)
        optimizer.zero_grad(set_to_none=True)
        loss = recommend_model(users=users_d[idx], items=items_d[idx], labels=labels_d[idx], flag=0)
        loss.backward()
        total += loss.detach().double()
        optimizer.step()
    total_loss = float(total.item())
    print("%d,%.5f" % (epoch, total_loss))
    return total_loss


def Test(dataset, Recmodel, epoch, best_recall, best_ndcg, best_iter):
    Recmodel = Recmodel.eval()
    with torch.no_grad():
        ret = batch_test.test(Recmodel, dataset.testRatings, dataset.testNegatives)
    perf_str = "Rec:  Epoch %d : recall=[%.4f, %.4f, %.4f],  ndcg=[%.4f, %.4f, %.4f]" % (
        epoch, ret["recall"][0], ret["recall"][1], ret["recall"][2], ret["ndcg"][0], ret["ndcg"][1],
        ret["ndcg"][2])
    print(perf_str)
    if ret["recall"][0] > best_recall[0]:
        best_recall, best_iter[0] = ret["recall"], epoch
    if ret["ndcg"][0] > best_ndcg[0]:
        best_ndcg, best_iter[1] = ret["ndcg"], epoch
    return best_recall, best_ndcg, best_iter


def main(argv=None):
    args = parse_args_r(argv)
    utils.set_seed(args.seed)
    if not torch.cuda.is_available():
        raise SystemExit("spex_b200.main_rec needs a B200 (sm_100a); there is no CPU fallback")
    device = torch.device("cuda", int(args.cuda_id))
    torch.cuda.set_device(device)
    dataset = dataloader.Loader(args)
    train_dataset = dataloader.LightTrainData(dataset.rec_train_data, dataset.m_item, dataset.train_mat)
    Recmodel = LightGCN(args, dataset).to(device)
    optimizer = FusedAdam(Recmodel.parameters(), lr=args.lr)
    best_recall, best_ndcg, best_iter = [0, 0, 0], [0, 0, 0], [0, 0]
    for epoch in range(args.epochs):
        start = time.time()
        Train(train_dataset, Recmodel, epoch, optimizer, device)
        best_recall, best_ndcg, best_iter = Test(dataset, Recmodel, epoch, best_recall, best_ndcg, best_iter)
        _ = time.time() - start
    print("--- Train Best ---")
    print("Rec:  recall=[%.4f, %.4f, %.4f],  ndcg=[%.4f, %.4f, %.4f]" % (
        best_recall[0], best_recall[1], best_recall[2], best_ndcg[0], best_ndcg[1], best_ndcg[2]))


if __name__ == "__main__":
    main()
