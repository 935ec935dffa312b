"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds (1) a plain-C restatement of the reference's CPU algorithm for the hot path
(`vecchia_oracle.c`, `tree_oracle.c`), (2) the recipe that builds the UNMODIFIED reference library from
`/root/reference` into `oracle/_ref/` (`Makefile.ref`), and (3) thin ctypes loaders for both.

Nothing under `gpboost_b200/` imports this package. Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may use it — as the checker / baseline, never as the
product path.
"""
from .build import build_oracle, oracle_lib_path, ref_lib_path, build_ref  # noqa: F401
