"""train_step_kernels.py — the SpMM launches of ONE receptive-field training step on the 1B-interaction graph
(forward: full layer, layer on S + N(S), layer on S; backward: masked layer on S + N(S), masked full layer, full
layer), for an ncu launch list:
    ncu -k regex:spmm --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\\
        l1tex__m_xbar2l1tex_read_bytes.sum,smsp__inst_executed.sum --clock-control none --csv \\
        --log-file out.csv python profiles/microbench/train_step_kernels.py
Template arguments in the kernel names: <D, U, hot, v8, masked>."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spex_b200 import ops, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    nu, m, ni, D = 10_000_000, 5_000_000, 1_000_000_000, 64
    keys = synthetic.generate_interactions(nu, m, ni, seed=2020, device=dev)
    g, _, _ = synthetic.build_norm_adj_device(keys, nu, m)
    del keys
    g.mark_hot_columns(D)
    nur = nu + 1
    W = synthetic.xavier_table(nur, m, D, 2020, dev).requires_grad_(True)
    gen = torch.Generator(device=dev)
    gen.manual_seed(11)
    users, items, labels = bench.sample_train_batch(torch, g, nur, m, 256, gen)
    ops.enable_persistent_workspaces(True)
    steps = int(os.environ.get("STEPS", 1))
    for _ in range(steps):
        W.grad = None
        out = ops.propagate_mean(W, g, 3, rows_needed=torch.cat([users, items + nur]))
        loss = ops.bce_loss(out, nur, users, items, labels)
        loss.backward()
    torch.cuda.synchronize()
    print("loss", float(loss))


if __name__ == "__main__":
    main()
