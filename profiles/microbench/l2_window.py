"""How large a gather window does the B200 L2 keep resident for the SpMM's 256-byte row gathers?
Rows of degree 128 whose columns are uniform in a window of W MB of the table; algorithmic GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spex_b200 import ops

dev = torch.device("cuda:0")
D = 64
deg = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_rows = (1 << 27) // deg
n_cols = (1 << 30) // (D * 4)          # 1 GB table
X = torch.randn(n_cols, D, device=dev)
rowptr = torch.arange(n_rows + 1, device=dev, dtype=torch.int64) * deg
val = torch.ones(n_rows * deg, device=dev)
Y = torch.empty(n_rows, D, device=dev)
g = torch.Generator(device=dev); g.manual_seed(0)
print("degree", deg)
for W in (8, 32, 128, 1024):
    wcols = (W << 20) // (D * 4)
    col = torch.randint(0, wcols, (n_rows * deg,), device=dev, generator=g, dtype=torch.int32)
    col = torch.sort(col.view(n_rows, deg), dim=1).values.reshape(-1).contiguous()
    G = ops.DeviceGraph(rowptr, col, val, n_cols)
    for _ in range(2): ops.spmm(G, X, Y=Y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.spmm(G, X, Y=Y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = (n_rows * deg * (8 + D * 4) + n_rows * (8 + D * 4)) / 1e9
    print(f"window {W:5d} MB: {ms:7.3f} ms  {gb / ms * 1e3:8.0f} GB/s algorithmic")
