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
