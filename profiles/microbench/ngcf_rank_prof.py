import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from spex_b200 import ops
from spex_b200.dataloader import SyntheticDataset
from spex_b200.ngcf import Model_Wrapper, build_ngcf_norm_adj
dev = torch.device('cuda:0')
nu2 = 8930; m2 = int(nu2*3.9)
ds2 = SyntheticDataset(nu2, m2, nu2*66, seed=2021, with_test=False)
adj = build_ngcf_norm_adj(ds2.trainUser, ds2.trainItem, nu2, m2)
torch.manual_seed(2020)
ng = Model_Wrapper({"n_users": nu2, "n_items": m2, "norm_adj": adj}, dev).to(dev)
ng.eval()
def T(fn, n=5):
    fn(); torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); a.record()
    for _ in range(n): r=fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b)/n,3), round((time.perf_counter()-t0)/n*1e3,3)
users = torch.arange(nu2, device=dev)
with torch.no_grad():
    print('propagate', T(lambda: ng.propagate()))
    ua, ia = ng.propagate(); ua, ia = ua.contiguous(), ia.contiguous()
    print('pack items', T(lambda: ops.pack_f16(ia, None, ops.TC_ITEM_MULTIPLE)))
    print('pack users', T(lambda: ops.pack_f16(ua, users, ops.TC_USER_MULTIPLE)))
    Ih, m_pad, imeta = ops.pack_f16(ia, None, ops.TC_ITEM_MULTIPLE)
    Uh, b_pad, umeta = ops.pack_f16(ua, users, ops.TC_USER_MULTIPLE)
    print('score f16 D=%d'%ua.shape[1], T(lambda: ops.score_topk_f16(Uh, umeta, nu2, b_pad, Ih, imeta, ia.shape[0], m_pad, ua.shape[1], 20, users, None, None)))
    print('score f32', T(lambda: ops.score_topk_f32(ua, ia, users, 20, None, None)))
    print('rank_topk', T(lambda: ng.rank_topk(users, k=20)))
    # D=64 for comparison on same shapes
    ua64, ia64 = ua[:, :64].contiguous(), ia[:, :64].contiguous()
    Ih6, m_pad6, imeta6 = ops.pack_f16(ia64, None, ops.TC_ITEM_MULTIPLE)
    Uh6, b_pad6, umeta6 = ops.pack_f16(ua64, users, ops.TC_USER_MULTIPLE)
    print('score f16 D=64', T(lambda: ops.score_topk_f16(Uh6, umeta6, nu2, b_pad6, Ih6, imeta6, ia64.shape[0], m_pad6, 64, 20, users, None, None)))
