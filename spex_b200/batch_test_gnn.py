"""Trust-path evaluation with the semantics of the reference's utility2/batch_test_gnn.py:27-44
(trust_test5): every test path has a candidate list whose LAST entry is the true next user; the
candidates' scores are ranked with torch.topk(50) and Recall/NDCG@{10,20,50} of the hit are summed
over paths and divided by the number of paths.  Vectorised per batch (the reference loops over
paths in Python)."""
from __future__ import annotations

import numpy as np
import torch

Ks = [10, 20, 50]


@torch.no_grad()
def trust_test5(model, test_data):
    l = test_data.length
    sums = np.zeros(6)
    for slice_indices in test_data.generate_batch(model.batch_size):
        scores, cand = model(None, None, None, slice_indices, test_data, 2)
        cs = torch.gather(scores, 1, cand)                     # score[target] per path
        k = min(50, cs.shape[1])
        top = cs.topk(k)[1]
        hit = (top == cand.shape[1] - 1).cpu().numpy()         # the positive is the last candidate
        rank = np.where(hit.any(1), hit.argmax(1), -1)
        for j, K in enumerate(Ks):
            inside = (rank >= 0) & (rank < K)
            sums[j] += inside.sum()                             # recall@K with one positive
            # ndcg@K: dcg / ideal, ideal = 1 whenever the hit is inside the top-50 list
            sums[3 + j] += np.sum(1.0 / np.log2(rank[inside] + 2.0))
    return tuple(float(x) / l for x in sums)
