"""oracle/ — CPU restatement of the reference's algorithm for the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `wedetect_b200/` may import this package: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs use it, and only as
the checker or the CPU arm.  See DESIGN.md §Oracle for what is pinned against the real reference.
"""
