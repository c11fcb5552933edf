"""spmm_blocking_sweep.py — one propagation layer of the 1B-interaction graph against the long-row
blocking parameters (DeviceGraph._build_plan: SPEX_L2_WINDOW_MB = table bytes per column block,
SPEX_HUB_EPB = average edges per (row, block) from which a long row is column-blocked; an optional
third field sets seg_len, the degree above which a row takes the long-row path).
    python profiles/microbench/spmm_blocking_sweep.py [--configs 16:64,32:32,64:16]
"""
import argparse
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
    ap.add_argument("--configs", default="16:64,32:32,64:16,64:32,32:16,48:24,96:16,16:64")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    nu, m, ni = int(10_000_000 * args.scale), int(5_000_000 * args.scale), int(1_000_000_000 * args.scale)
    D = 64
    keys = synthetic.generate_interactions(nu, m, ni, seed=2020, device=dev)
    g, _, _ = synthetic.build_norm_adj_device(keys, nu, m)
    del keys
    g.mark_hot_columns(D)
    N = nu + 1 + m
    X = synthetic.xavier_table(nu + 1, m, D, 2020, dev)
    Y = torch.empty_like(X)
    ref = None
    for cfg in args.configs.split(","):
        win, epb, *rest = cfg.split(":")
        os.environ["SPEX_L2_WINDOW_MB"], os.environ["SPEX_HUB_EPB"] = win, epb
        g.seg_len = int(rest[0]) if rest else 1024   # long-row threshold = segment cap (third field, optional)
        g._build_plan(D)
        torch.cuda.synchronize()

        def full():
            _capi.call("spex_spmm_csr_f32", ptr(g.rowptr), ptr(g.col), ptr(g.val), ptr(X), N, D, ptr(Y), None, 1.0,
                       None, 1.0, g.plan(D), stream_ptr())
        full()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            full()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        if ref is None:
            ref = Y.clone()
        err = float((Y - ref).abs().max())
        print(json.dumps({"window_mb": int(win), "hub_edges_per_block": int(epb), "seg_len": g.seg_len, "n_long": g.n_long, "layer_ms": round(best, 3),
                          "n_seg": g.n_seg, "n_hub": getattr(g, "n_hub", 0), "partial_mb": g.n_seg * D * 4 >> 20,
                          "max_abs_diff_vs_first": err}), flush=True)


if __name__ == "__main__":
    main()
