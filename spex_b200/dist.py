"""Row-partitioned propagation across the GPUs of one NVSwitch box.

The reference has no distributed code; its only scale-out hook is the serial row folding
``A_split`` (/root/reference/LightGCN_SPEX/code/utility1/dataloader.py:167-177, model.py:84-89):
cut the adjacency into row blocks, multiply each block by the full table, concatenate.  This module
is that partition with one process per GPU: rank p owns the contiguous node range
[bounds[p], bounds[p+1]) (balanced by nnz), holds the CSR row block of those rows, and every layer

    1. assembles the full E^(k) [N, D] on every rank (the exchange),
    2. runs the local CSR SpMM over its rows, accumulating the layer mean in its own slice.

Three exchange modes:
  "nccl"  one all-gather of the [rows_p, D] slices per layer (torch.distributed over NCCL/NVLink);
  "push"  the SpMM epilogue itself stores every output row into all peers' next-layer tables with
          P2P stores over NVLink (spex_spmm_csr_f32_push): transfer overlaps the gather-bound
          math row by row; layers are separated by a stream-ordered 4-byte all-reduce (barrier).
          E^(0) has no producing kernel to fuse with and travels by NCCL all-gather (measured
          faster than P2P stores or copy engines, which stay selectable: e0_exchange).

  "mcast" the same fusion with ONE store per row to an NVSwitch multicast mapping of the tables
          (NVLS; torch symmetric memory provides the mapping): a row leaves the GPU once instead of
          P-1 times, which removes the NVLink egress bound of "push" (spex_spmm_csr_f32_mcast,
          spex_mcast_rows_f32 for E^(0)).

The local multiply is injectable (``local_spmm``) so the orchestration can be tested on CPU with the
oracle's SpMM under the gloo backend; the default is the CUDA kernel and there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def slice_bounds(bounds: Sequence[int], rank: int):
    return int(bounds[rank]), int(bounds[rank + 1])


class _CudaArray:
    """Expose a raw device allocation to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape),
                                         "typestr": typestr, "version": 3, "strides": None}


class PartitionedPropagator:
    """K-layer propagation + layer mean over a row partition; returns this rank's rows of the mean."""

    def __init__(self, local_graph, bounds: Sequence[int], D: int, K: int, group=None,
                 mode: str = "nccl", local_spmm: Optional[Callable] = None, device=None):
        self.g = local_graph
        self.bounds = [int(b) for b in bounds]
        self.D, self.K = int(D), int(K)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if len(self.bounds) != self.world + 1:
            raise ValueError("bounds must have world_size + 1 entries")
        self.N = self.bounds[-1]
        self.r0, self.r1 = slice_bounds(self.bounds, self.rank)
        self.mode = mode
        self.local_spmm = local_spmm
        self.device = device if device is not None else getattr(local_graph, "device", torch.device("cpu"))
        self._X: List[torch.Tensor] = []
        self._peer_ptrs = None
        self._raw = []
        self._flag = None
        # push mode: how E^(0) travels: "nccl" (all-gather; NVLS multicast makes it the fastest on an
        # NVSwitch box: 6.2 ms for 3.84 GB at 8 GPUs), "push" (SM stores, 10.5 ms) or "copy" (copy
        # engines, 13.2 ms)
        self.e0_exchange = "nccl"
        self._copy_streams = []
        self.timing = None        # set to [] to collect per-phase CUDA-event pairs (bring-up)
        self._mc = None           # multicast pointers of the two tables (mode "mcast")
        self._symm = []
        if mode == "push":
            if local_spmm is not None:
                raise ValueError("push mode is CUDA-only")
            self._setup_push()
        elif mode == "mcast":
            if local_spmm is not None:
                raise ValueError("mcast mode is CUDA-only")
            self._setup_mcast()
            self.e0_exchange = "mcast"
        elif mode == "nccl":
            self._X = [torch.empty(self.N, self.D, dtype=torch.float32, device=self.device)
                       for _ in range(2 if self.K > 1 else 1)]
        else:
            raise ValueError("mode must be 'nccl', 'push' or 'mcast'")

    # ---- push mode: IPC-mapped double-buffered tables --------------------------------------------
    def _setup_push(self):
        from ._capi import call

        nbytes = self.N * self.D * 4
        handles = []
        for _ in range(2):
            p = C.c_void_p()
            h = C.create_string_buffer(64)
            call("spex_ipc_alloc", nbytes, C.byref(p), h)
            self._raw.append(p)
            handles.append(h.raw)
            self._X.append(torch.as_tensor(_CudaArray(p.value, (self.N, self.D)), device=self.device))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=self.group)
        self._peer_ptrs = []  # [buf][rank] raw pointers of every rank's table (own included)
        self._opened = []
        for b in range(2):
            ptrs = []
            for r in range(self.world):
                if r == self.rank:  # our own table is just one more destination of the epilogue
                    ptrs.append(self._raw[b].value)
                    continue
                q = C.c_void_p()
                call("spex_ipc_open", gathered[r][b], C.byref(q))
                self._opened.append(q)
                ptrs.append(q.value)
            self._peer_ptrs.append((C.c_void_p * max(len(ptrs), 1))(*ptrs))
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        dist.barrier(group=self.group)

    # ---- mcast mode: symmetric-memory tables with an NVSwitch multicast mapping -------------------
    def _setup_mcast(self):
        import torch.distributed._symmetric_memory as symm

        group = self.group if self.group is not None else dist.group.WORLD
        self._mc = []
        for _ in range(2):
            t = symm.empty(self.N * self.D, dtype=torch.float32, device=self.device)
            h = symm.rendezvous(t, group)
            if not h.multicast_ptr:
                raise RuntimeError("this box has no NVLS multicast support: use mode='push'")
            self._symm.append((t, h))
            self._X.append(t.view(self.N, self.D))
            self._mc.append(int(h.multicast_ptr))
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        dist.barrier(group=self.group)

    def close(self):
        if self.mode == "mcast" and self._symm:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            self._X, self._symm, self._mc = [], [], None
        if self.mode == "push" and self._raw:
            from ._capi import call

            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            for q in self._opened:
                call("spex_ipc_close", q)
            self._X = []
            for p in self._raw:
                call("spex_ipc_free", p)
            self._raw = []

    # ---- exchange ---------------------------------------------------------------------------------
    def _all_gather_rows(self, X_full: torch.Tensor, local: torch.Tensor):
        if self.world == 1:
            X_full[self.r0: self.r1].copy_(local)
            return
        outs = [X_full[self.bounds[r]: self.bounds[r + 1]] for r in range(self.world)]
        if dist.get_backend(self.group) == "nccl":
            # uneven slices: NCCL issues one grouped broadcast per rank straight into the views
            dist.all_gather(outs, local.contiguous(), group=self.group)
        else:  # gloo (CPU tests): all_gather needs equal sizes, so broadcast slice by slice
            outs[self.rank].copy_(local)
            for r in range(self.world):
                dist.broadcast(outs[r], src=dist.get_global_rank(self.group, r) if self.group else r,
                               group=self.group)

    def _stream_barrier(self):
        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)

    # ---- one layer --------------------------------------------------------------------------------
    def _layer(self, X_full, Y_local, addend, Z_local, z_scale, push_buf=None):
        if self.local_spmm is not None:
            acc = self.local_spmm(self.g, X_full)
            if Y_local is not None:
                Y_local.copy_(acc)
            Z_local.copy_((addend + acc) * z_scale)
            return
        from . import ops
        from ._capi import call, ptr, stream_ptr

        if push_buf is None:
            ops.spmm(self.g, X_full, Y=Y_local, addend=addend, addend_scale=1.0, Z=Z_local, z_scale=z_scale)
            return
        if self.mode == "mcast":   # fused SpMM + all-gather, one multicast store per row
            call("spex_spmm_csr_f32_mcast", ptr(self.g.rowptr), ptr(self.g.col), ptr(self.g.val), ptr(X_full),
                 self.g.n_rows, self.D, self.r0, C.c_void_p(self._mc[push_buf]),
                 ptr(addend), 1.0, ptr(Z_local), float(z_scale), self.g.plan(self.D), stream_ptr())
            return
        # fused SpMM + all-gather: rows go to every peer's (and our own) next-layer table
        call("spex_spmm_csr_f32_push", ptr(self.g.rowptr), ptr(self.g.col), ptr(self.g.val), ptr(X_full),
             self.g.n_rows, self.D, self.r0, self._peer_ptrs[push_buf], self.world,
             ptr(addend), 1.0, ptr(Z_local), float(z_scale), self.g.plan(self.D), stream_ptr())

    def _mark(self, name):
        if self.timing is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timing.append((name, ev))

    def phase_ms(self):
        """[(phase, ms)] between consecutive marks of the last propagate() (timing enabled)."""
        torch.cuda.synchronize()
        t = self.timing or []
        return [(t[i + 1][0], t[i][1].elapsed_time(t[i + 1][1])) for i in range(len(t) - 1)]

    def _exchange_e0(self, E0_local):
        if self.mode == "mcast" and self.e0_exchange == "mcast":
            from ._capi import call, ptr, stream_ptr

            self._stream_barrier()   # everybody is done reading buffer 0 from the previous call
            call("spex_mcast_rows_f32", ptr(E0_local), E0_local.shape[0], self.D, self.r0,
                 C.c_void_p(self._mc[0]), stream_ptr())
        elif self.mode == "push" and self.e0_exchange in ("push", "copy"):
            from ._capi import call, ptr, stream_ptr

            # everybody must be done reading buffer 0 (layer K-1 or K-2 of the previous call)
            self._stream_barrier()
            if self.e0_exchange == "push":   # SM stores: one read, P stores per element
                call("spex_push_rows_f32", ptr(E0_local), E0_local.shape[0], self.D, self.r0,
                     self._peer_ptrs[0], self.world, stream_ptr())
                return
            # copy engines: one cudaMemcpyAsync per peer on its own stream, peers visited in
            # rotated order so that the ranks do not all target the same GPU at the same time
            main = torch.cuda.current_stream()
            if not self._copy_streams:
                self._copy_streams = [torch.cuda.Stream() for _ in range(self.world - 1)]
            start = torch.cuda.Event()
            start.record(main)
            nbytes = E0_local.numel() * 4
            off = self.r0 * self.D * 4
            E0c = E0_local.contiguous()
            done = []
            for j, st in enumerate(self._copy_streams):
                peer = (self.rank + 1 + j) % self.world
                st.wait_event(start)
                call("spex_memcpy_peer_async", C.c_void_p(self._peer_ptrs[0][peer] + off), ptr(E0c), nbytes,
                     C.c_void_p(st.cuda_stream))
                ev = torch.cuda.Event()
                ev.record(st)
                done.append(ev)
            self._X[0][self.r0: self.r1].copy_(E0c)
            for ev in done:
                main.wait_event(ev)
        else:
            self._all_gather_rows(self._X[0], E0_local)

    def propagate(self, E0_local: torch.Tensor) -> torch.Tensor:
        """E0_local: this rank's rows [r0, r1) of the fused table.  Returns mean_k E^(k)[r0:r1]."""
        K = self.K
        out = torch.empty_like(E0_local)
        if K == 0:
            out.copy_(E0_local)
            return out
        inv = 1.0 / (K + 1)
        if self.timing is not None:
            self.timing = []
        self._mark("start")
        self._exchange_e0(E0_local)
        self._mark("e0_exchange")
        if self.mode in ("push", "mcast"):
            self._stream_barrier()  # E^(0) complete everywhere; nobody still reads buffer 1
        self._mark("barrier")
        Y = None
        for k in range(K):
            last = k == K - 1
            X_full = self._X[k & 1]
            addend = E0_local if k == 0 else out
            if self.mode in ("push", "mcast"):
                self._layer(X_full, None, addend, out, inv if last else 1.0,
                            push_buf=None if last else (k + 1) & 1)
                self._mark(f"layer{k + 1}")
                if not last:
                    self._stream_barrier()
                    self._mark("barrier")
            else:
                if not last and Y is None:
                    Y = torch.empty_like(E0_local)
                self._layer(X_full, None if last else Y, addend, out, inv if last else 1.0)
                self._mark(f"layer{k + 1}")
                if not last:
                    self._all_gather_rows(self._X[(k + 1) & 1], Y)
                    self._mark("all_gather")
        return out
