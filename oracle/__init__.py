"""CPU oracle for the cope-nerf NeuS hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch-CPU restatement of the reference algorithm
(`/root/reference/model/*.py`), written function by function with the reference
file:line each one follows.  It exists so that the CUDA product path can be
checked on a GPU box where `/root/reference` is not mounted.

Rules (see DESIGN.md "oracle"):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
    `--impl reference` legs may import anything from here;
  * nothing under `cope_nerf_b200/` imports it — the product path has no CPU
    fallback and fails loudly when the CUDA library is missing;
  * the restatement is PINNED against the imported reference itself: see
    `tests/golden/make_golden.py` (run in the build container, where the
    reference is mounted) and `tests/test_oracle_golden.py`.
"""
from .neus_oracle import *  # noqa: F401,F403
