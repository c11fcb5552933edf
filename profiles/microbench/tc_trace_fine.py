"""Bring-up (library built with EXTRA=-DSPEX_TC_TRACE=2): cost of each synchronisation step of
epilogue warp 2, tiles 2000.., in cycles."""
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
Ib, m_pad = ops.pack_bf16(I, None, ops.TC_ITEM_MULTIPLE)
Ub, b_pad = ops.pack_bf16(U, users, ops.TC_USER_MULTIPLE)
trace = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
_capi.lib.spex_debug_tc_trace.argtypes = [ctypes.c_void_p]
_capi.lib.spex_debug_tc_trace(ctypes.c_void_p(trace.data_ptr()))
ops.score_topk_bf16(Ub, n_u, b_pad, Ib, m, m_pad, 20, users, None, None)
torch.cuda.synchronize()
t = trace.cpu().view(64, 8).numpy()
base = t[0, 7]
print("tile | acquire: start  poll  syncwarp  fence_after | release: start fence_before syncwarp arrive")
for i in range(0, 24):
    r = t[i] - base
    print(f"{2000+i:5d} | {r[7]:7d} +{r[0]-r[7]:4d} +{r[1]-r[0]:4d} +{r[2]-r[1]:4d} | {r[3]:7d} +{r[4]-r[3]:4d} +{r[5]-r[4]:4d} +{r[6]-r[5]:4d}")
