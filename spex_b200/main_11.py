"""Rec + trust-path multi-task entry point: `python -m spex_b200.main_11 --dataset weibo ...`

Same flow and printed lines as /root/reference/LightGCN_SPEX/code/main_11.py: per batch of 256
rec samples the training paths of the batch's users are gathered (capped at
trust_batch_size = n_paths // n_batches by random.sample, main_11.py:54-59), the model returns
(loss1, loss2), loss = loss1 + loss2 (:69); Train prints
'%d,%.5f,%.5f,%.5f,%.5f' % (epoch, precision1, precision2, total_loss1, total_loss2) (:74) and Test
prints the 'Rec:' and 'Trust:' lines (:88-99).
"""
from __future__ import annotations

import os
import pickle
import random
import time
from collections import defaultdict

import numpy as np
import torch

from . import batch_test, dataloader, utils
from .batch_test_gnn import trust_test5
from .lg_parser import parse_args_r
from .main_rec import BATCH, epoch_batches
from .model_expert_s import LightGCN
from .optim import FusedAdam
from .path_data import Data


def load_trust(args, n_users):
    base = os.path.join(args.data_path, args.dataset, "trust")
    train_raw = pickle.load(open(os.path.join(base, "train.txt"), "rb"))
    test_raw = pickle.load(open(os.path.join(base, "test2.txt"), "rb"))
    user_path_indx = defaultdict(list)
    for i, p in enumerate(train_raw[0]):
        user_path_indx[p[0]].append(i)
    return (Data(train_raw, n_users, shuffle=False), Data(test_raw, n_users, shuffle=False, test=True),
            user_path_indx, len(train_raw[0]))


def Train(train_dataset, train_paths, user_path_indx, trust_batch_size, Recmodel, epoch, optimizer, device):
    train_dataset.ng_sample()
    Recmodel.train()
    users, items, labels = (torch.from_numpy(a) for a in train_dataset.arrays())
    users_d, items_d = users.to(device), items.to(device)
    labels_d = labels.to(device=device, dtype=torch.float32)
    tot = torch.zeros(2, dtype=torch.float64, device=device)
    for idx in epoch_batches(users.numel(), BATCH):
        optimizer.zero_grad(set_to_none=True)
        unique_user = set(users[idx].tolist())
        path_index = []
        for u in unique_user:
            path_index.extend(user_path_indx[u])
        if len(path_index) > trust_batch_size:
            path_index = random.sample(path_index, trust_batch_size)
        idx_d = idx.to(device)
        loss1, loss2 = Recmodel(users=users_d[idx_d], items=items_d[idx_d], labels=labels_d[idx_d],
                                slice_indices=np.array(list(path_index), dtype=int),
                                trust_data=train_paths, flag=0)
        loss = loss1 + loss2
        loss.backward()
        tot += torch.stack([loss1.detach(), loss2.detach()]).double()
        optimizer.step()
    precision1 = torch.exp(-2 * Recmodel.task_weights[0])
    precision2 = torch.exp(-2 * Recmodel.task_weights[1])
    t1, t2 = (float(x) for x in tot.tolist())
    print("%d,%.5f,%.5f,%.5f,%.5f" % (epoch, precision1, precision2, t1, t2))
    return t1, t2


def Test(dataset, test_paths, Recmodel, epoch, best_recall, best_ndcg, best_iter, best_result):
    Recmodel = Recmodel.eval()
    with torch.no_grad():
        ret = batch_test.rec_test(Recmodel, dataset.testRatings, dataset.testNegatives)
        print("Rec:  Epoch %d : recall=[%.4f, %.4f, %.4f],  ndcg=[%.4f, %.4f, %.4f]" % (
            epoch, ret["recall"][0], ret["recall"][1], ret["recall"][2], ret["ndcg"][0], ret["ndcg"][1],
            ret["ndcg"][2]))
        if ret["recall"][0] > best_recall[0]:
            best_recall, best_iter[0] = ret["recall"], epoch
        if ret["ndcg"][0] > best_ndcg[0]:
            best_ndcg, best_iter[1] = ret["ndcg"], epoch
        r10, r20, r50, n10, n20, n50 = trust_test5(Recmodel, test_paths)
        print("Trust:Epoch %d : recall=[%.4f, %.4f, %.4f],  ndcg=[%.4f, %.4f, %.4f]" % (
            epoch, r10, r20, r50, n10, n20, n50))
        if r10 >= best_result[0]:
            best_result[:3] = [r10, r20, r50]
        if n10 >= best_result[3]:
            best_result[3:] = [n10, n20, n50]
    return best_recall, best_ndcg, best_iter, best_result


def main(argv=None):
    args = parse_args_r(argv)
    utils.set_seed(args.seed)
    if not torch.cuda.is_available():
        raise SystemExit("spex_b200.main_11 needs a B200 (sm_100a); there is no CPU fallback")
    device = torch.device("cuda", int(args.cuda_id))
    torch.cuda.set_device(device)
    dataset = dataloader.Loader(args)
    train_dataset = dataloader.LightTrainData(dataset.rec_train_data, dataset.m_item, dataset.train_mat)
    train_paths, test_paths, user_path_indx, n_paths = load_trust(args, dataset.n_users)
    n_batches = -(-len(train_dataset) // BATCH)
    trust_batch_size = n_paths // n_batches
    Recmodel = LightGCN(args, dataset).to(device)
    optimizer = FusedAdam(Recmodel.parameters(), lr=args.lr)
    best_recall, best_ndcg, best_iter = [0, 0, 0], [0, 0, 0], [0, 0]
    best_result = [0, 0, 0, 0, 0, 0]
    for epoch in range(args.epochs):
        start = time.time()
        Train(train_dataset, train_paths, user_path_indx, trust_batch_size, Recmodel, epoch, optimizer, device)
        best_recall, best_ndcg, best_iter, best_result = Test(dataset, test_paths, Recmodel, epoch, best_recall,
                                                              best_ndcg, best_iter, best_result)
        _ = time.time() - start
    print("--- Train Best ---")
    print("Rec:  recall=[%.4f, %.4f, %.4f],  ndcg=[%.4f, %.4f, %.4f]" % (
        best_recall[0], best_recall[1], best_recall[2], best_ndcg[0], best_ndcg[1], best_ndcg[2]))
    return best_recall, best_ndcg, best_result


if __name__ == "__main__":
    main()
