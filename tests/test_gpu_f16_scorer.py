"""GPU parity of the fp16-accumulator tcgen05 scorer (csrc/score_topk_f16.cu).

Contract under test: the tensor core's fp16-accumulated scores are ONLY a filter; the returned top-k is
the exact top-k (score descending, ties by ascending item id) of the fp32 scores of the fp16 operands
fp16(x * 2^s) written by spex_pack_f16.  The oracle here is a float64 matmul of exactly those operand
values (north-star semantics, not reference-pinned: the reference only has the dense matmul of
NGCF_SPEX/code/utility/batch_test.py:158, no fused ranking).
"""
import numpy as np
import pytest
import torch

from helpers import check_topk_against_scores

pytestmark = pytest.mark.gpu


def _operands(x, meta):
    """float64 values of the packed operands: fp16(x * scale) / scale."""
    sc = float(meta[0])
    return (x * sc).half().double() / sc


def _run(ops, Ud, Id, users, k, rp=None, col=None, D=64):
    Ih, m_pad, imeta = ops.pack_f16(Id, None, ops.TC_ITEM_MULTIPLE)
    Uh, b_pad, umeta = ops.pack_f16(Ud, users, ops.TC_USER_MULTIPLE)
    idx, val = ops.score_topk_f16(Uh, umeta, users.numel(), b_pad, Ih, imeta, Id.shape[0], m_pad, D, k, users, rp, col)
    torch.cuda.synchronize()
    return idx, val, umeta.cpu(), imeta.cpu()


@pytest.mark.parametrize("D", [64, 128])
def test_topk_f16_matches_exact_scores(cuda_device, D):
    from spex_b200 import ops
    from spex_b200.graph import build_interaction_csr

    torch.manual_seed(5)
    n_u, m = 300, 5000
    U = torch.randn(n_u, D) * 0.3
    I = torch.randn(m, D) * 0.3 * (1.0 + 3.0 * torch.rand(m, 1))      # row norms spread over 4x
    users = torch.arange(n_u)
    rng = np.random.default_rng(2)
    rp, col = build_interaction_csr(rng.integers(0, n_u, 6000), rng.integers(0, m, 6000), n_u, m)
    masked = [col[rp[u]: rp[u + 1]] for u in range(n_u)]
    rpd, cold = torch.from_numpy(rp).to(cuda_device), torch.from_numpy(col).to(cuda_device)
    Ud, Id = U.to(cuda_device), I.to(cuda_device)
    for k in (1, 20, 50, ops.TC_MAX_K):
        idx, val, umeta, imeta = _run(ops, Ud, Id, users.to(cuda_device), k, rpd, cold, D)
        # power-of-two scales, every scaled row norm below 2^7
        for meta in (umeta, imeta):
            assert float(meta[0]) == 2.0 ** round(np.log2(float(meta[0]))) and float(meta[2]) < 128.0
        scores = torch.matmul(_operands(U, umeta), _operands(I, imeta).t()).numpy()
        # fp32 FMA chain vs float64: 64-128 products of magnitude <= |u||v|
        check_topk_against_scores(idx, val, scores, k, masked, atol=2e-5 * float(np.abs(scores).max()))
    # against the exact fp32 CUDA-core scorer on the same operand values: same items except near-ties
    Uq = _operands(U, umeta).float().to(cuda_device)
    Iq = _operands(I, imeta).float().to(cuda_device)
    i32, v32 = ops.score_topk_f32(Uq, Iq, users.to(cuda_device), 20, rpd, cold)
    i16, v16, _, _ = _run(ops, Ud, Id, users.to(cuda_device), 20, rpd, cold, D)
    assert float((v32 - v16).abs().max()) < 2e-5 * float(v32.abs().max())
    assert float((i32 == i16).float().mean()) > 0.99


def test_topk_f16_ties_and_duplicates(cuda_device):
    """Exact score ties (duplicated and all-zero item rows) are broken by ascending item id."""
    from spex_b200 import ops

    torch.manual_seed(11)
    n_u, m, D = 130, 3000, 64
    U = torch.randn(n_u, D) * 0.3
    base = torch.randn(40, D) * 0.3
    I = base[torch.randint(0, 40, (m,))].clone()      # every item row is one of 40 vectors
    I[::7] = 0.0                                       # and every 7th is all-zero (score 0)
    Ud, Id = U.to(cuda_device), I.to(cuda_device)
    users = torch.arange(n_u, device=cuda_device)
    for k in (5, 20, 50):
        i16, v16, umeta, imeta = _run(ops, Ud, Id, users, k)
        scores = torch.matmul(_operands(U, umeta), _operands(I, imeta).t())
        order = torch.sort(-scores, dim=1, stable=True).indices[:, :k]
        ref_v = torch.gather(scores, 1, order)
        got_v = torch.gather(scores, 1, i16.cpu().long())
        assert torch.allclose(got_v, ref_v, atol=1e-5), k            # same score multiset per rank
        ids = i16.cpu().long()
        # identical item rows give bit-identical exact scores: the whole order must be the stable one
        # wherever distinct scores are clearly apart
        top = -torch.sort(-scores, dim=1, stable=True).values[:, : k + 1]
        gap = top[:, :-1] - top[:, 1:]
        clean = ~((gap > 1e-9) & (gap < 1e-4)).any(dim=1)
        assert int(clean.sum()) > n_u // 2
        assert torch.equal(ids[clean], order[clean]), k


def test_filter_never_drops_a_true_topk_item_on_adversarial_near_ties(cuda_device):
    """Adversarial for an fp16 filter: thousands of items whose exact scores differ by far less than
    one fp16 ulp of the score (so their fp16-accumulated scores collide or even invert), spread over
    many tiles, plus large-norm decoys.  The result must still be the exact fp32 ranking."""
    from spex_b200 import ops

    torch.manual_seed(3)
    n_u, m, D = 128, 128 * 60, 64
    U = torch.randn(n_u, D) * 0.5
    I = torch.randn(m, D) * 0.05
    # 3000 near-duplicates of one strong direction per user block: score differences ~1e-5 relative
    strong = torch.randn(D)
    strong = strong / strong.norm()
    pos = torch.randperm(m)[:3000]
    I[pos] = strong * 0.8 + torch.randn(3000, D) * 2e-4
    # users aligned with it so that those items fill the top of every list, at very close scores
    U = U * 0.05 + strong * (0.5 + 0.5 * torch.rand(n_u, 1))
    # decoys with 10x norm (they set Vmax, i.e. the loosest filter margin) but orthogonal on average
    dec = torch.randperm(m)[:50]
    I[dec] = torch.randn(50, D) * 0.5
    Ud, Id = U.to(cuda_device), I.to(cuda_device)
    users = torch.arange(n_u, device=cuda_device)
    for k in (20, 50):
        idx, val, umeta, imeta = _run(ops, Ud, Id, users, k)
        Uq, Iq = _operands(U, umeta), _operands(I, imeta)
        scores = torch.matmul(Uq, Iq.t())
        # exact reference on the GPU in fp32 with the same operand values
        i32, v32 = ops.score_topk_f32(Uq.float().to(cuda_device), Iq.float().to(cuda_device), users, k)
        # (1) the set of returned items has the same k-th score as the exact ranking (fp64 check):
        kth_ref = -torch.sort(-scores, dim=1).values[:, k - 1]
        got = torch.gather(scores, 1, idx.cpu().long())
        assert bool((got.min(dim=1).values >= kth_ref - 1e-9 - 4e-6 * scores.abs().max()).all())
        # (2) and it is the same ranking as the fp32 CUDA-core scorer except where two fp32 scores are
        # within rounding of each other
        assert float((v32 - val).abs().max()) < 4e-6 * float(v32.abs().max())
        agree = float((i32 == idx).float().mean())
        assert agree > 0.9, agree
    # the near-ties really are below fp16 resolution: many distinct exact scores share one fp16 value
    s_top = -torch.sort(-scores, dim=1).values[:, :50]
    sc = float(umeta[0]) * float(imeta[0])
    collide = ((s_top * sc).half()[:, 1:] == (s_top * sc).half()[:, :-1]).float().mean()
    assert float(collide) > 0.3


def test_model_rank_topk_f16_vs_fp32_and_metrics(cuda_device):
    """Full-rank Recall/NDCG@20 from the f16 scorer, the bf16 scorer and the exact fp32 scorer against
    the oracle's masked top-k (utility1/metrics.py:61-80 semantics on the ranked lists)."""
    from oracle import lightgcn_oracle as O
    from spex_b200 import metrics as M
    from test_gpu_parity import _small_model

    ds, model, (uw, iw), A = _small_model(cuda_device)
    model.eval()
    users = np.arange(ds.n_users)
    i32, v32 = model.rank_topk(users, k=20, precision="fp32")
    i16, v16 = model.rank_topk(users, k=20, precision="f16")
    ib, vb = model.rank_topk(users, k=20, precision="bf16")
    scale = float(v32.abs().max())
    assert float((v32 - v16).abs().max()) <= 2e-3 * scale          # fp16 operands: 2^-11 relative each
    same = [(len(set(a.tolist()) & set(b.tolist())) / 20.0) for a, b in zip(i32.cpu().numpy(), i16.cpu().numpy())]
    assert np.mean(same) > 0.97
    # metrics: hold-out = the dataset's test item of every user (truth as CSR)
    from spex_b200.graph import build_interaction_csr

    tu = np.array([u for u in users if u in ds.testRatings for _ in ds.testRatings[u]], dtype=np.int64)
    ti = np.array([i for u in users if u in ds.testRatings for i in ds.testRatings[u]], dtype=np.int64)
    trp, tcol = build_interaction_csr(tu, ti, ds.n_users, ds.m_items)
    # oracle ranking (torch.matmul + masked_fill + topk on the reference computer() outputs)
    with torch.no_grad():
        ou, oi = O.computer(uw, iw, A, 3)
    rp, col = build_interaction_csr(ds.trainUser, ds.trainItem, ds.n_users + 1, ds.m_items)
    oidx, _ = O.topk_masked(torch.matmul(ou[users], oi.t()), rp, col, users, 20)

    def metric(idx):
        r, n = M.fullrank_recall_ndcg(np.asarray(idx), trp, tcol, 20)
        return {"recall": float(r.mean()), "ndcg": float(n.mean())}

    mo, m32, m16, mb = metric(oidx.numpy()), metric(i32.cpu().numpy()), metric(i16.cpu().numpy()), metric(ib.cpu().numpy())
    # the exact fp32 scorer must return the oracle's lists except where two fp32 scores are within
    # rounding of each other (exact ties are ordered identically: ascending item id)
    agree = float((torch.from_numpy(oidx.numpy()) == i32.cpu().long()).float().mean())
    assert agree > 0.995, agree
    n_users = len(users)
    for key in ("recall", "ndcg"):
        assert abs(m32[key] - mo[key]) <= 1.0 / n_users, (key, m32, mo)
        assert abs(m16[key] - mo[key]) <= 2.0 / n_users, (key, m16, mo)     # fp16 operands: near-tie swaps only
        assert abs(mb[key] - mo[key]) <= 6.0 / n_users, (key, mb, mo)       # bf16 operands: 8x coarser


def test_near_tie_probe(cuda_device):
    """ops.f16_filter_is_selective: random tables keep the tensor-core filter, tables whose items are almost
    parallel (every score within the filter's error band of the k-th best) go to the exact scorer."""
    from spex_b200 import ops

    torch.manual_seed(5)
    U = torch.randn(500, 64, device=cuda_device)
    I = torch.randn(20000, 64, device=cuda_device)
    users = torch.arange(500, device=cuda_device)
    assert ops.f16_filter_is_selective(U, I, users)
    base = torch.randn(1, 64, device=cuda_device)
    I2 = base + 1e-4 * torch.randn(20000, 64, device=cuda_device)       # near-collinear items
    assert not ops.f16_filter_is_selective(U, I2, users)
