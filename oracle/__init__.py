"""TEST INFRASTRUCTURE — CPU / fp32 restatements of the reference's hot path (one module per model family, each function
citing the reference file:line it follows), the harness that imports the UNMODIFIED reference from /root/reference to
generate golden vectors (build container only), and the synthetic weights / inputs both sides share.

Only `tests/`, `__graft_entry__.smoke()`, `bench.py`'s cpu_baseline / `--impl reference` arm and the `tools/make_golden*.py`
generators may import this package; nothing under `declip_b200/` does (tests/test_abi.py::test_no_oracle_import_in_product)."""
