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

In the two fused modes the tables are a ring of THREE buffers: E^(0) of a call lives in slot a, layer k
reads slot a+k and writes slot a+k+1 (mod 3), the next call's E^(0) goes to slot a+K.  That slot is idle
while the LAST layer runs (which stores nothing over NVLink), so `propagate(E0, next_E0_local=...)`
publishes the next call's table from a small high-priority side-stream kernel DURING the last layer:
the E^(0) exchange (5.4 ms + 2.5 ms of skew at 8 GPUs in round 1, un-overlapped) disappears from the
critical path, and the ring makes the "everybody is done reading the buffer" barrier of round 1
unnecessary (every slot's previous readers are separated from its next writer by a barrier that
already exists).

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
        self.n_tables = 3         # ring of table buffers (module docstring)
        self._peer_ptrs = None
        self._raw = []
        self._flag = None
        # push mode: how E^(0) travels: "nccl" (all-gather; NVLS multicast makes it the fastest on an
        # NVSwitch box: 6.2 ms for 3.84 GB at 8 GPUs), "push" (SM stores, 10.5 ms) or "copy" (copy
        # engines, 13.2 ms)
        self.e0_exchange = "nccl"
        self._copy_streams = []
        self.timing = None        # set to [] to collect per-phase CUDA-event pairs (bring-up)
        self._mc = None           # multicast pointers of the tables (mode "mcast")
        self._symm = []
        self._slot = 0            # ring slot that holds / receives E^(0) of the next propagate()
        self._staged = None       # (key, event): E^(0) already published into slot self._slot
        self._side = None         # high-priority stream of the background publish
        self.bg_ctas = 148        # CTAs of the background publish kernel (it must not crowd out the SpMM)
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
                       for _ in range(self.n_tables)]
        else:
            raise ValueError("mode must be 'nccl', 'push' or 'mcast'")

    # ---- push mode: IPC-mapped double-buffered tables --------------------------------------------
    def _setup_push(self):
        from ._capi import call

        nbytes = self.N * self.D * 4
        handles = []
        for _ in range(self.n_tables):
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
        for b in range(self.n_tables):
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
        for _ in range(self.n_tables):
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

    def _exchange_e0(self, E0_local, slot: int = 0, background: bool = False):
        """All-gather this rank's rows of E^(0) into ring slot `slot` of every rank.  No barrier is
        needed before it in the fused modes: the ring guarantees that the slot's last readers finished
        before a barrier every rank has already passed (module docstring)."""
        if self.mode == "mcast" and self.e0_exchange == "mcast":
            from ._capi import call, ptr, stream_ptr

            call("spex_mcast_rows_f32_ex", ptr(E0_local), E0_local.shape[0], self.D, self.r0,
                 C.c_void_p(self._mc[slot]), self.bg_ctas if background else 0, stream_ptr())
        elif self.mode == "push" and self.e0_exchange in ("push", "copy"):
            from ._capi import call, ptr, stream_ptr

            if self.e0_exchange == "push":   # SM stores: one read, P stores per element
                call("spex_push_rows_f32_ex", ptr(E0_local), E0_local.shape[0], self.D, self.r0,
                     self._peer_ptrs[slot], self.world, self.bg_ctas if background else 0, stream_ptr())
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
                call("spex_memcpy_peer_async", C.c_void_p(self._peer_ptrs[slot][peer] + off), ptr(E0c), nbytes,
                     C.c_void_p(st.cuda_stream))
                ev = torch.cuda.Event()
                ev.record(st)
                done.append(ev)
            self._X[slot][self.r0: self.r1].copy_(E0c)
            for ev in done:
                main.wait_event(ev)
        else:
            self._all_gather_rows(self._X[slot], E0_local)

    @staticmethod
    def _key(t: torch.Tensor):
        return (t.data_ptr(), t._version, tuple(t.shape))

    def can_prefetch(self) -> bool:
        """Can the next call's E^(0) be published during this call's last layer?  In the fused modes that
        needs one of our own exchange kernels (NCCL on a side stream would race the stream-ordered
        barriers on the same communicator); in "nccl" mode the publish is a plain all-gather issued
        before the last layer (same ring bookkeeping, no overlap)."""
        if self.mode == "nccl":
            return True
        return (self.mode == "mcast" and self.e0_exchange == "mcast") or \
               (self.mode == "push" and self.e0_exchange == "push")

    def propagate(self, E0_local: torch.Tensor, out: torch.Tensor = None,
                  next_E0_local: torch.Tensor = None, next_ready=None) -> torch.Tensor:
        """E0_local: this rank's rows [r0, r1) of the fused table.  Returns mean_k E^(k)[r0:r1]
        (written into `out` if given).  `next_E0_local`: this rank's rows of the table of the NEXT
        call, if the caller already has it (an inference / evaluation sweep over many tables, or
        the optimiser's output): it is published to all ranks while the last layer of this call runs
        (`next_ready`: CUDA event after which next_E0_local may be read, e.g. the end of its upload)."""
        K = self.K
        if out is None:
            out = torch.empty_like(E0_local)
        if K == 0:
            out.copy_(E0_local)
            return out
        inv = 1.0 / (K + 1)
        if self.timing is not None:
            self.timing = []
        fused = self.mode in ("push", "mcast")
        nt = self.n_tables
        a = self._slot
        self._mark("start")
        staged, self._staged = self._staged, None
        if staged is not None and staged[1] is not None:
            torch.cuda.current_stream().wait_event(staged[1])   # the publish of the last call (even a stale one
                                                                # must be over before this slot is written again)
        if staged is None or staged[0] != self._key(E0_local):
            self._exchange_e0(E0_local, slot=a)
        self._mark("e0_exchange")
        if fused:
            self._stream_barrier()  # E^(0) complete everywhere
        self._mark("barrier")
        Y = None
        for k in range(K):
            last = k == K - 1
            X_full = self._X[(a + k) % nt]
            addend = E0_local if k == 0 else out
            if last and next_E0_local is not None and self.can_prefetch():
                self._publish_next(next_E0_local, (a + K) % nt, next_ready)
            if fused:
                self._layer(X_full, None, addend, out, inv if last else 1.0,
                            push_buf=None if last else (a + k + 1) % nt)
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
                    self._all_gather_rows(self._X[(a + k + 1) % nt], Y)
                    self._mark("all_gather")
        self._slot = (a + K) % nt
        return out

    def _publish_next(self, next_E0_local, slot, next_ready=None):
        """Exchange of the next call's E^(0) into ring slot `slot`.  Fused modes: on the high-priority
        side stream; it starts once everything enqueued so far on the main stream (the barrier before
        the last layer) is done and runs next to the last layer, which stores nothing over NVLink."""
        if self.mode == "nccl":
            if next_ready is not None:
                torch.cuda.current_stream().wait_event(next_ready)
            self._all_gather_rows(self._X[slot], next_E0_local)
            self._staged = (self._key(next_E0_local), None)
            return
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(priority=-1)
        ready = torch.cuda.Event()
        ready.record(main)
        self._side.wait_event(ready)
        if next_ready is not None:
            self._side.wait_event(next_ready)
        with torch.cuda.stream(self._side):
            self._exchange_e0(next_E0_local, slot=slot, background=True)
            done = torch.cuda.Event()
            done.record(self._side)
        self._staged = (self._key(next_E0_local), done)
