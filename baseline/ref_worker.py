"""CPU arm on the UNMODIFIED reference (illiumst/marl-factory-grid installed into baseline/_ref by baseline/install_ref.sh):
multiprocess vectorised envs on the host cores, SURVEY.md 8d / App. E protocol - one `Factory` per worker,
`random.seed(worker)`, 30 warm-up steps (numba JIT + the floor graph), uniform random actions, in-place `reset()` on done,
stdout suppressed.  Only bench.py's CPU legs import this; nothing here is part of the product."""
import contextlib
import io
import os
import random
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / 'baseline' / '_ref'
STUBS = ROOT / 'tests' / 'golden' / 'stubs'          # import stand-ins for gymnasium / pygame (absent from the image)

_ENV = None
_N_ACT = None
_RNG = None


def available() -> bool:
    return (REF / 'marl_factory_grid' / 'environment' / 'factory.py').exists()


def worker_init(cfg_path: str, warmup: int = 30):
    global _ENV, _N_ACT, _RNG
    sys.dont_write_bytecode = True
    for p in (str(REF), str(STUBS)):
        if p not in sys.path:
            sys.path.insert(0, p)
    wid = os.getpid()
    random.seed(wid)
    with contextlib.redirect_stdout(io.StringIO()):
        from marl_factory_grid.environment.factory import Factory
        _ENV = Factory(str(cfg_path))
        _ENV.reset()
    _N_ACT = [len(a.actions) for a in _ENV.state['Agent']]
    _RNG = random.Random(wid)
    worker_run(warmup)


def worker_run(steps: int):
    """Advance this worker's env by `steps`; returns (agent_steps, seconds)."""
    env, n_act, rng = _ENV, _N_ACT, _RNG
    out = io.StringIO()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(out):
        for _ in range(steps):
            try:
                _, _, _, done, _ = env.step([rng.randrange(n) for n in n_act])
            except Exception:
                # the reference's own defect: Maintainer.get_move_action pops an exhausted route list
                # (modules/maintenance/entities.py:89) every few thousand steps; an unattended run would stop here.  The episode is
                # abandoned (the step still counts as CPU work done) and the env re-spawned.
                done = True
            if done:
                env.reset()
            if out.tell() > 1 << 20:
                out.seek(0)
                out.truncate()
    return steps * len(n_act), time.perf_counter() - t0
