"""Test infrastructure only: CPU restatements and reference builds used as checkers.

Nothing in here is imported by the product package
(`gaussian-splatting_deformable_b200/`).  Only tests/, __graft_entry__.smoke() and
bench.py's reference / cpu_baseline legs may use it.
"""
