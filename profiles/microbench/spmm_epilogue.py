"""spmm_epilogue.py — what the fused epilogue of a propagation layer costs on the 1B-interaction graph:
one layer with Y only / Y + Z = addend + acc (layers 1..K-1 of computer()) / Z only in place (last layer),
and the K = 3 spex_propagate_mean_f32 call, back to back.  Run on a B200:
    python profiles/microbench/spmm_epilogue.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from spex_b200 import ops, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    nu, m, ni, D = 10_000_000, 5_000_000, 1_000_000_000, 64
    keys = synthetic.generate_interactions(nu, m, ni, seed=2020, device=dev)
    g, _, _ = synthetic.build_norm_adj_device(keys, nu, m)
    del keys
    g.mark_hot_columns(D)
    X = synthetic.xavier_table(nu + 1, m, D, 2020, dev)
    Y, Z, A = torch.empty_like(X), torch.empty_like(X), torch.randn_like(X)

    def timed(fn, reps=4):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return round(best, 3)

    out = {
        "Y_only_ms": timed(lambda: ops.spmm(g, X, Y=Y)),
        "Y_and_Z_addend_ms": timed(lambda: ops.spmm(g, X, Y=Y, addend=A, Z=Z)),
        "Y_and_Z_inplace_ms": timed(lambda: ops.spmm(g, X, Y=Y, addend=Z, Z=Z)),
        "Z_only_inplace_ms": timed(lambda: ops.spmm(g, X, addend=Z, Z=Z, z_scale=0.25)),
        "propagate_mean_K3_ms": timed(lambda: ops.propagate_mean(X, g, 3)),
    }
    # 10 K=3 calls back to back (what bench.py times): clocks settle under the power cap
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.propagate_mean(X, g, 3)
    e1.record()
    torch.cuda.synchronize()
    out["propagate_mean_K3_ms_avg_of_10_back_to_back"] = round(e0.elapsed_time(e1) / 10, 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
