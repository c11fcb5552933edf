"""world_size-2 row-partitioned propagation on CPU (gloo): the orchestration of spex_b200.dist
(partition by nnz, per-layer exchange, local layer-mean) with the oracle's SpMM injected as the
local multiply, compared with the unpartitioned oracle.  The CUDA kernel takes that slot on GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import oracle_graph, random_graph
from oracle import lightgcn_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spex_b200.dist import PartitionedPropagator
        from spex_b200.graph import build_norm_adj, partition_rows_by_nnz

        nu, m, D = 300, 120, 16
        u, i = random_graph(nu, m, 3000, 21, hub_items=2, hub_degree=200)
        g = build_norm_adj(u, i, nu + 1, m)
        bounds = partition_rows_by_nnz(g.rowptr, world)
        blk = g.row_block(bounds[rank], bounds[rank + 1])
        rows = torch.from_numpy(blk.rows_of_entries())
        Ablk = torch.sparse_coo_tensor(torch.stack([rows, torch.from_numpy(blk.col.astype(np.int64))]),
                                       torch.from_numpy(blk.val), (blk.n_rows, g.n_cols)).coalesce()
        torch.manual_seed(0)
        E = torch.randn(g.n_rows, D)
        prop = PartitionedPropagator(Ablk, bounds, D, K, mode="nccl", device=torch.device("cpu"),
                                     local_spmm=lambda A, X: torch.sparse.mm(A, X))
        mine = E[bounds[rank]: bounds[rank + 1]].clone()
        out = prop.propagate(mine)
        out2 = prop.propagate(E[bounds[rank]: bounds[rank + 1]].clone())  # buffers are reusable
        # ring of three tables + publish of the NEXT call's table before the last layer: a sweep over
        # different tables, each announced one call ahead, and one unannounced (stale publish) call
        tables = [mine, mine * 2.0 - 0.5, mine * -1.0 + 0.25, mine]
        outs = []
        for j, t in enumerate(tables):
            nxt = tables[j + 1] if j + 1 < len(tables) else mine * 3.0   # the last announcement is never used
            outs.append(prop.propagate(t, next_E0_local=nxt))
        stale = prop.propagate(mine)   # announced table was mine * 3: must be ignored
        direct1 = prop.propagate(tables[1].clone())
        direct2 = prop.propagate(tables[2].clone())
        ring_ok = (torch.equal(outs[0], out) and torch.equal(outs[3], out) and torch.equal(stale, out)
                   and torch.equal(outs[1], direct1) and torch.equal(outs[2], direct2))
        ru, ri = O.computer(E[: nu + 1], E[nu + 1:], oracle_graph(u, i, nu + 1, m), K)
        want = torch.cat([ru, ri])[bounds[rank]: bounds[rank + 1]]
        err = float((out - want).abs().max())
        q.put((rank, err, bool(torch.equal(out, out2)) and bool(ring_ok), bounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K", [1, 3])
def test_row_partition_world2(K):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same, bounds in res:
        assert err < 1e-6, (rank, err)
        assert same
        assert 0 < bounds[1] < bounds[2]


def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spex_b200.dist import PartitionedPropagator, PartitionedTrainer, _TorchTrainOps
        from spex_b200.graph import build_norm_adj, partition_rows_by_nnz

        nu, m, D, K = 120, 80, 16, 3
        u, i = random_graph(nu, m, 1500, 5)
        g = build_norm_adj(u, i, nu + 1, m)
        bounds = partition_rows_by_nnz(g.rowptr, world)
        r0, r1 = bounds[rank], bounds[rank + 1]
        blk = g.row_block(r0, r1)
        rows = torch.from_numpy(blk.rows_of_entries())
        Ablk = torch.sparse_coo_tensor(torch.stack([rows, torch.from_numpy(blk.col.astype(np.int64))]),
                                       torch.from_numpy(blk.val), (blk.n_rows, g.n_cols)).coalesce()
        torch.manual_seed(3)
        W = torch.randn(g.n_rows, D, dtype=torch.float64).float() * 0.1
        prop = PartitionedPropagator(Ablk, bounds, D, K, mode="nccl", device=torch.device("cpu"),
                                     local_spmm=lambda A, X: torch.sparse.mm(A, X))
        tr = PartitionedTrainer(prop, W[r0:r1].clone(), nu + 1, lr=1e-2, train_ops=_TorchTrainOps())
        rng = np.random.default_rng(0)
        losses = []
        for step in range(3):
            users = torch.from_numpy(rng.integers(0, nu, 64))
            users[:10] = users[0]                                  # duplicates
            items = torch.from_numpy(rng.integers(0, m, 64))
            labels = torch.from_numpy(rng.integers(0, 2, 64)).float()
            losses.append(float(tr.step(users, items, labels)))
        # single-process oracle: the reference's forward + autograd + torch.optim.Adam on the same batches
        uw = W[: nu + 1].clone().requires_grad_(True)
        iw = W[nu + 1:].clone().requires_grad_(True)
        opt = torch.optim.Adam([uw, iw], lr=1e-2)
        A = oracle_graph(u, i, nu + 1, m)
        rng = np.random.default_rng(0)
        ref_losses = []
        for step in range(3):
            users = torch.from_numpy(rng.integers(0, nu, 64))
            users[:10] = users[0]
            items = torch.from_numpy(rng.integers(0, m, 64))
            labels = torch.from_numpy(rng.integers(0, 2, 64))
            opt.zero_grad()
            loss = O.bce_forward(uw, iw, A, K, users, items, labels)
            loss.backward()
            opt.step()
            ref_losses.append(float(loss))
        want = torch.cat([uw, iw]).detach()[r0:r1]
        err = float((tr.W - want).abs().max() / want.abs().max())
        lerr = max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses))
        q.put((rank, err, lerr))
    finally:
        dist.destroy_process_group()


def test_partitioned_training_world2_matches_single_process_oracle():
    """Row-owned table / gradient / Adam moments, batch rows assembled by the exact all-reduce,
    backward = the same partitioned propagation on gradients: after three steps the weights match the
    reference's forward + autograd + torch.optim.Adam run in one process."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, lerr in res:
        assert err < 2e-5, (rank, err)
        assert lerr < 1e-5, (rank, lerr)
