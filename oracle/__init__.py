"""TEST INFRASTRUCTURE ONLY.  `oracle/` holds (1) shims that make the read-only reference importable in
this container, (2) a plain-PyTorch fp32 restatement of the hot-path math, (3) the script that
generated `tests/golden/`.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it; the product (`transfusion_pytorch_b200`) never does."""
