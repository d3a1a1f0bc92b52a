"""CPU oracle for the `bean run` SVI hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `crispr_bean_b200/` may import this
package; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` do.
"""
