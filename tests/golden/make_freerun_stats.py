#!/usr/bin/env python
"""Free-running statistics of the UNMODIFIED reference (test infrastructure, run in the build container only).

The engine's free-running mode (Philox spawn, dirt re-spawn draws, maintainer routing) cannot be replayed against the
reference draw by draw - the reference uses CPython's `random` / PCG64 streams and networkx tie-breaking.  What can be pinned
is the DISTRIBUTION: this script runs the reference with uniform random actions for many episodes and records

  * spawn marginals: how often agents start on door tiles (never), initial dirt count split, distinctness of group tiles,
    per-tile spawn frequency of agent 0 (chi-square against uniform over the empty tiles)
  * episode statistics: length histogram, done-reason shares, mean per-agent return, collisions per step

into tests/golden/freerun_stats.json, which tests/test_freerun_distribution.py compares with the device code.

Protocol: one persistent `Factory` per worker process, `reset()` per episode (constructing a Factory costs ~8 s on level
`large`); `random.seed(worker)`.  Usage: python tests/golden/make_freerun_stats.py [--procs 8]
"""
import argparse
import contextlib
import io
import json
import multiprocessing as mp
import random
import sys
from collections import Counter
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
PLAN = {'cfg1': 2400, 'cfg4': 400}          # episodes per config (the dense stress configs make the reference's maintainer raise IndexError in long free runs)


def worker(args):
    cfg, n_episodes, seed = args
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(HERE / 'stubs'))
    sys.path.insert(1, '/root/reference')
    random.seed(seed)
    rng = random.Random(seed + 1000)
    out = dict(lengths=[], reasons=Counter(), returns=[], dirt_n0=Counter(), agent_on_door=0, agent0_tiles=Counter(),
               group_dup=0, collisions=0, steps=0, episodes=0, maint_moves=Counter())
    with contextlib.redirect_stdout(io.StringIO()):
        from marl_factory_grid.environment.factory import Factory
        f = Factory(str(REPO / 'marl_factory_grid_b200' / 'configs' / f'{cfg}.yaml'))
        for _ in range(n_episodes):
            f.reset()
            st = f.state
            agents = list(st['Agent'])
            doors = {d.pos for d in st['Doors']} if 'Doors' in st.entities.names else set()
            out['agent_on_door'] += sum(a.pos in doors for a in agents)
            out['agent0_tiles'][str(tuple(int(v) for v in agents[0].pos))] += 1
            if 'DirtPiles' in st.entities.names:
                out['dirt_n0'][len(st['DirtPiles'])] += 1
            for g in ('Items', 'ChargePods', 'Destinations', 'DropOffLocations', 'Machines', 'Maintainers'):
                if g in st.entities.names:
                    pos = [e.pos for e in st[g]]
                    out['group_dup'] += len(pos) - len(set(pos))
            n_act = [len(a.actions) for a in agents]
            ret = np.zeros(len(agents))
            t = 0
            while True:
                _, _, r, done, info = f.step([rng.randrange(n) for n in n_act])
                ret += np.asarray(r, float)
                t += 1
                out['collisions'] += sum(1 for k, v in info.items() if k.endswith('_Collisions') and v)
                if done:
                    break
            # done reason: the first valid done result (factory.py:243-248); re-derive it from the rules
            dres = [x for x in st.check_done() if x.validity]
            out['reasons'][dres[0].identifier if dres else 'unknown'] += 1
            out['lengths'].append(t)
            out['returns'].append(ret.tolist())
            out['steps'] += t
            out['episodes'] += 1
    out['reasons'] = dict(out['reasons']); out['dirt_n0'] = {str(k): v for k, v in out['dirt_n0'].items()}
    out['agent0_tiles'] = dict(out['agent0_tiles']); out['maint_moves'] = dict(out['maint_moves'])
    return out


def merge(parts):
    lengths = np.concatenate([np.asarray(p['lengths']) for p in parts])
    returns = np.concatenate([np.asarray(p['returns']) for p in parts], 0)
    reasons, dirt_n0, tiles = Counter(), Counter(), Counter()
    for p in parts:
        reasons.update(p['reasons']); dirt_n0.update(p['dirt_n0']); tiles.update(p['agent0_tiles'])
    n = int(sum(p['episodes'] for p in parts))
    edges = [0, 10, 20, 40, 80, 120, 160, 200, 300, 400, 499, 100000]
    hist, _ = np.histogram(lengths, bins=edges)
    return {'episodes': n, 'env_steps': int(sum(p['steps'] for p in parts)),
            'length_mean': float(lengths.mean()), 'length_hist_edges': edges, 'length_hist': hist.tolist(),
            'done_reasons': dict(reasons), 'return_mean': returns.mean(0).tolist(), 'return_std': returns.std(0).tolist(),
            'dirt_n0': dict(dirt_n0), 'agent_on_door': int(sum(p['agent_on_door'] for p in parts)),
            'group_duplicate_tiles': int(sum(p['group_dup'] for p in parts)),
            'collisions_per_step': float(sum(p['collisions'] for p in parts)) / max(1, sum(p['steps'] for p in parts)),
            'agent0_tiles': dict(tiles)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--procs', type=int, default=8)
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    path = HERE / 'freerun_stats.json'
    result = json.loads(path.read_text()) if path.exists() else {}
    for cfg, n in PLAN.items():
        if args.only and cfg not in args.only.split(','):
            continue
        per = (n + args.procs - 1) // args.procs
        with mp.get_context('spawn').Pool(args.procs) as pool:
            parts = pool.map(worker, [(cfg, per, 100 + i) for i in range(args.procs)])
        result[cfg] = merge(parts)
        print(cfg, {k: v for k, v in result[cfg].items() if k not in ('agent0_tiles',)}, flush=True)
        path.write_text(json.dumps(result, indent=1, sort_keys=True))


if __name__ == '__main__':
    main()
