"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's LightGCN_SPEX hot path.

Nothing under spex_b200/ imports this package.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it, and only as the checker.
"""
