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
            # a sweep over different tables, each announced one call ahead (published to the peers by the
            # background kernel during the previous call's last layer), then an unannounced call
            ok_ring = True
            if prop.can_prefetch():
                tabs = [E[r0:r1].clone(), (E[r0:r1] * 2.0 - 0.5).clone(), (E[r0:r1] * -1.0).clone(), E[r0:r1].clone()]
                outs = [prop.propagate(t, next_E0_local=tabs[j + 1] if j + 1 < len(tabs) else tabs[1])
                        for j, t in enumerate(tabs)]
                stale = prop.propagate(tabs[0])           # the announced table was tabs[1]: must be ignored
                d1 = prop.propagate(tabs[1].clone())
                d2 = prop.propagate(tabs[2].clone())
                ok_ring = (torch.equal(outs[0], a) and torch.equal(outs[3], a) and torch.equal(stale, a)
                           and torch.equal(outs[1], d1) and torch.equal(outs[2], d2))
            prop.close()
            res[f"{mode}/{e0}"] = (bool(torch.equal(a, want[r0:r1])),
                                   bool(torch.equal(a, b) and torch.equal(a, c) and ok_ring), phases)
        # a width that takes the generic-D kernel (its epilogue ignored the multicast table)
        D2 = 48
        E2 = (torch.randn(N, D2, generator=torch.Generator().manual_seed(1)) * 0.1).to(dev)
        want2 = ops.propagate_mean(E2, full, K)
        for mode in [m for m, e in modes if m == e and m != "nccl"]:
            prop = PartitionedPropagator(lg, bounds, D2, K, mode=mode, device=dev)
            a2 = prop.propagate(E2[r0:r1].clone())
            prop.close()
            res[f"{mode}/D48"] = (bool(torch.equal(a2, want2[r0:r1])), True, ["layer3", "e0_exchange"])
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
