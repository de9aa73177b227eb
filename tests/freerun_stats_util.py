"""Free-running episode statistics of the engine's device code (host build or CUDA), in the same shape as
tests/golden/freerun_stats.json (recorded from the unmodified reference by tests/golden/make_freerun_stats.py)."""
from collections import Counter

import numpy as np

REASON_NAMES = {2: 'DoneAtMaxStepsReached', 3: 'DoneOnAllDirtCleaned', 4: 'DoneAtBatteryDischarge', 5: 'DoneAtDestinationReach',
                6: 'DoneAtMaintainerCollision', 7: 'Collisions'}


def spawn_stats(sim_factory, es, n_envs):
    sim = sim_factory(n_envs)
    sim.reset()
    f = sim.fields if not callable(getattr(sim, 'fields_numpy', None)) else sim.fields_numpy()
    doors = {(int(x) << 8) | int(y) for x, y in es.door_pos}
    apos = f['apos']
    out = {'agent_on_door': int(sum(int(p) in doors for p in apos.reshape(-1))),
           'agent0_tiles': Counter(str((int(p) >> 8, int(p) & 255)) for p in apos[0]),
           'dirt_n0': Counter(str(int(n)) for n in f['dirt_n'][0]) if es.has_dirt else Counter(), 'group_duplicate_tiles': 0}
    for name in ('item_pos', 'pod_pos', 'dest_pos', 'drop_pos', 'mach_pos', 'maint_pos'):
        if name in f:
            col = f[name]
            out['group_duplicate_tiles'] += int(sum(len(set(col[:, e])) != col.shape[0] for e in range(col.shape[1])))
    return out


def episode_stats(es, n_envs, max_steps, rng, step_fn):
    """One complete episode per env (no auto reset; an env is ignored after its first done), like the reference run.
    step_fn(actions [N, A] int32) -> (reward [N, A], done [N], flags [N, A + 1])."""
    A = es.n_agents
    ret = np.zeros((n_envs, A))
    length = np.zeros(n_envs, np.int64)
    alive = np.ones(n_envs, bool)
    reasons = Counter()
    collisions = steps = 0
    n_act = es.n_actions
    for _ in range(max_steps):
        if not alive.any():
            break
        a = np.stack([rng.integers(0, n, n_envs) for n in n_act], 1).astype(np.int32)
        r, d, fl = step_fn(a)
        ret[alive] += r[alive].astype(np.float64)
        length[alive] += 1
        steps += int(alive.sum())
        collisions += int(((fl[alive, :A] & 4) != 0).sum())
        ended = alive & (d != 0)
        for e in np.nonzero(ended)[0]:
            reasons[REASON_NAMES.get(int(fl[e, A]), 'other')] += 1
        alive &= ~ended
    assert not alive.any(), 'some episodes did not end'
    edges = [0, 10, 20, 40, 80, 120, 160, 200, 300, 400, 499, 100000]
    return {'episodes': n_envs, 'length_mean': float(length.mean()), 'length_hist': np.histogram(length, bins=edges)[0].tolist(),
            'done_reasons': dict(reasons), 'return_mean': ret.mean(0).tolist(), 'return_std': ret.std(0).tolist(),
            'collisions_per_step': collisions / max(steps, 1), 'lengths': length}
