"""oracle/ - CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

What it is: a plain-Python / numpy restatement of marl-factory-grid's `Factory.step` and
`OBSBuilder.build_for_agent` semantics (the *actual* behaviour), one
environment at a time.  Every function cites the reference file:line it follows.

Who may import it: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`, and only as the checker or as the timed CPU baseline.  The
product package (`marl_factory_grid_b200/`) never imports it and fails loudly when the CUDA
extension is missing - there is no CPU fallback.

How it is pinned: the reference is pure Python and runs in the build container, so the oracle is
checked step by step against traces of the UNMODIFIED reference (`tests/golden/*.npz`, produced by
the committed `tests/golden/make_golden.py`): integer state, door timers, f64 dirt amounts /
battery levels, done flags and observations bit-for-bit, rewards to <= 1e-12
(`tests/test_oracle_golden.py`).  Two modes, as in SURVEY.md §8c:
  faithful=True   == the untouched reference (uid-equality artefact, oracle-U)
  faithful=False  == the identity-patched reference (oracle-I)
"""
from .env import OracleEnv, load_snapshot  # noqa: F401
from .rays import ray_targets, full_rays  # noqa: F401
