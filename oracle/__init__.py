"""CPU oracle for the FoundDiff reverse-diffusion sampling hot path.

TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this package.  The product (`founddiff_b200/`) never does.
"""
