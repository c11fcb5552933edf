"""Bring-up: per-tile timeline of the tcgen05 scorer (CTA 0, tiles 2000..2063), clock64 cycles."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spex_b200 import ops, _capi

dev = torch.device("cuda:0")
torch.manual_seed(0)
n_u, m, D = 148 * 2 * 128, 1_000_000, 64
U = torch.randn(n_u, D, device=dev) * 0.1
I = torch.randn(m, D, device=dev) * 0.1
users = torch.arange(n_u, device=dev)
Ib, m_pad = ops.pack_bf16(I, None, 128)
Ub, b_pad = ops.pack_bf16(U, users, 128)
trace = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
_capi.lib.spex_debug_tc_trace.argtypes = [ctypes.c_void_p]
_capi.lib.spex_debug_tc_trace(ctypes.c_void_p(trace.data_ptr()))
ops.score_topk_bf16(Ub, n_u, b_pad, Ib, m, m_pad, 20, users, None, None)
torch.cuda.synchronize()
t = trace.cpu().view(64, 8).numpy()
base = t[0, 0]
print("tile | MMA thread: loop_top  B_tile_ready  buffer_free  chain_issued | epilogue warp 2: acc_ready  drained   (cycles)")
for i in range(0, 40):
    r = t[i] - base
    print(f"{2000+i:5d} | {r[5]:8d} {r[6]:8d} {r[0]:8d} {r[1]:8d} | {r[2]:8d} {r[3]:8d}   wait_B {r[6]-r[5]:5d} wait_buf {r[0]-r[6]:5d} issue {r[1]-r[0]:4d} ->ready {r[2]-r[1]:4d} epilogue {r[3]-r[2]:5d}")
d = t[1:40, 0] - t[0:39, 0]
print("mean period per tile:", d.mean())
