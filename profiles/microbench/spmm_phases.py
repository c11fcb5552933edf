"""spmm_phases.py — where does a propagation layer of the 1B-interaction graph spend its time?

Times (the r02 runs also took a --variants list selecting the persistent software-pipelined
kernels of profiles/experiments/r02_spmm_persistent_pipelined.cu.txt; they lost by 30-45 % and
were removed from the library, results in profiles/r02_spmm_phases.txt):
    user rows alone   (short-row kernel over rows [0, n_user_rows): gathers of Zipf-popular item rows)
    item rows alone   (short-row kernel over the item rows of degree <= seg_len: random user rows)
    full layer        (short rows + column-blocked segments + fix-up) -> segments by difference
and checks that every variant returns bit-identical tables.  Run on a B200:
    python profiles/microbench/spmm_phases.py [--scale 1.0]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from spex_b200 import _capi, synthetic  # noqa: E402
from spex_b200._capi import ptr, stream_ptr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    nu, m, ni = int(10_000_000 * args.scale), int(5_000_000 * args.scale), int(1_000_000_000 * args.scale)
    D = 64
    keys = synthetic.generate_interactions(nu, m, ni, seed=2020, device=dev)
    g, _, _ = synthetic.build_norm_adj_device(keys, nu, m)
    del keys
    g.mark_hot_columns(D)
    nur, N = nu + 1, nu + 1 + m
    X = synthetic.xavier_table(nur, m, D, 2020, dev)
    Y = torch.empty_like(X)
    lib = _capi.lib
    lib.spex_debug_spmm_rows.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                                          C.POINTER(_capi.LongPlan), C.c_void_p]
    lib.spex_debug_spmm_rows.restype = C.c_int

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    def rows(a, b):
        rc = lib.spex_debug_spmm_rows(ptr(g.rowptr), ptr(g.col), ptr(g.val), ptr(X), a, b, D, ptr(Y), g.plan(D),
                                      stream_ptr())
        assert rc == 0, rc

    def full():
        _capi.call("spex_spmm_csr_f32", ptr(g.rowptr), ptr(g.col), ptr(g.val), ptr(X), N, D, ptr(Y), None, 1.0,
                   None, 1.0, g.plan(D), stream_ptr())

    deg = g.rowptr[1:] - g.rowptr[:-1]
    short = deg <= g.seg_len
    e_user = int(deg[:nur].sum())
    e_item_short = int(deg[nur:][short[nur:]].sum())
    e_long = g.nnz - e_user - e_item_short
    out = {"workload": f"{nu} users x {m} items, nnz(A)={g.nnz}", "edges": {"user_rows": e_user,
           "item_rows_short": e_item_short, "long_rows": e_long}, "n_long": g.n_long, "n_seg": g.n_seg, "variants": {}}
    ref = None
    for v in [0]:
        t_user = timed(lambda: rows(0, nur))
        t_item = timed(lambda: rows(nur, N))
        t_full = timed(full)
        full()
        torch.cuda.synchronize()
        if ref is None:
            ref = Y.clone()
            same = True
        else:
            same = bool(torch.equal(ref, Y))
        r = {"user_rows_ms": round(t_user, 3), "item_rows_ms": round(t_item, 3), "full_layer_ms": round(t_full, 3),
             "segments_by_difference_ms": round(t_full - t_user - t_item, 3),
             "user_rows_gather_TBps": round(e_user * 264 / t_user / 1e9, 2),
             "item_rows_gather_TBps": round(e_item_short * 264 / t_item / 1e9, 2),
             "bit_identical_to_first": same}
        out["variants"][str(v)] = r
        print(json.dumps({str(v): r}), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
