"""Row-partitioned propagation on real GPUs (needs >= 2 devices, skipped otherwise): the NCCL
all-gather exchange and the fused SpMM + P2P push exchange (spex_spmm_csr_f32_push,
spex_push_rows_f32) against the single-GPU kernel on the same graph.  Same row partition as the
reference's serial folds (/root/reference/LightGCN_SPEX/code/utility1/dataloader.py:167-177)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import random_graph

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from spex_b200 import ops
        from spex_b200.dist import PartitionedPropagator
        from spex_b200.graph import build_norm_adj, partition_rows_by_nnz

        nu, m, D, K = 5000, 3000, 64, 3
        u, i = random_graph(nu, m, 120000, 21, hub_items=3, hub_degree=2500)
        g = build_norm_adj(u, i, nu + 1, m)
        N = g.n_rows
        full = ops.DeviceGraph(torch.from_numpy(g.rowptr).to(dev), torch.from_numpy(g.col).to(dev),
                               torch.from_numpy(g.val).to(dev), N)
        torch.manual_seed(0)
        E = (torch.randn(N, D) * 0.1).to(dev)
        want = ops.propagate_mean(E, full, K)
        bounds = partition_rows_by_nnz(g.rowptr, world)
        r0, r1 = bounds[rank], bounds[rank + 1]
        lo, hi = int(g.rowptr[r0]), int(g.rowptr[r1])
        lg = ops.DeviceGraph((full.rowptr[r0: r1 + 1] - lo).contiguous(), full.col[lo:hi].clone(),
                             full.val[lo:hi].clone(), N, None, full.seg_len, row_offset=r0)
        res = {}
        modes = [("nccl", "nccl"), ("push", "push"), ("push", "copy"), ("push", "nccl")]
        try:   # NVLS multicast needs an NVSwitch box
            probe = PartitionedPropagator(lg, bounds, D, K, mode="mcast", device=dev)
            probe.close()
            modes += [("mcast", "mcast"), ("mcast", "nccl")]
        except RuntimeError as e:
            res["mcast_unavailable"] = (True, True, ["layer3", "e0_exchange", str(e)])
        for mode, e0 in modes:
            prop = PartitionedPropagator(lg, bounds, D, K, mode=mode, device=dev)
            prop.e0_exchange = e0
            a = prop.propagate(E[r0:r1].clone())
            b = prop.propagate(E[r0:r1].clone())   # buffers are reusable, result reproducible
            prop.timing = []
            c = prop.propagate(E[r0:r1].clone())
            phases = [n for n, _ in prop.phase_ms()]
            prop.close()
            res[f"{mode}/{e0}"] = (bool(torch.equal(a, want[r0:r1])), bool(torch.equal(a, b) and torch.equal(a, c)),
                                   phases)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_partitioned_propagation_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, res in out:
        for key, (exact, reproducible, phases) in res.items():
            # the row partition does not change any row's summation order: bit-identical
            assert exact and reproducible, (rank, key)
            assert "layer3" in phases and "e0_exchange" in phases
