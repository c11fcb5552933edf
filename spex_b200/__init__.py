"""spex_b200 — B200-native (sm_100a) hot path of XMUDM/SPEX's LightGCN_SPEX.

Host code mirrors the reference's module layout for this path:
    spex_b200.model        <- LightGCN_SPEX/code/utility1/model.py
    spex_b200.dataloader   <- LightGCN_SPEX/code/utility1/dataloader.py
    spex_b200.batch_test   <- LightGCN_SPEX/code/utility1/batch_test.py
    spex_b200.metrics      <- LightGCN_SPEX/code/utility1/metrics.py
Everything numeric runs in libspex_b200.so (include/spex_b200.h); importing spex_b200.ops or
spex_b200.model without the built library raises — there is no CPU fallback.
"""
__version__ = "0.1.0"
