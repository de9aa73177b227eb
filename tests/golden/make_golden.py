#!/usr/bin/env python
"""Generate golden replay traces by running the UNMODIFIED reference (illiumst/marl-factory-grid)
in this container.  Test infrastructure only: nothing in the product imports this file, and the
GPU box never runs it (there is no /root/reference there) - it only reads the committed *.npz.

Protocol: one fresh `Factory` per episode, `Object._u_idx.clear()` and
`random.seed(s)` before construction, exactly one `reset()`.  Two oracles:
  * mode "U": the untouched reference (uid-equality artefact active)          -> parity target
  * mode "I": harness-level identity patch of Object.__eq__/__hash__          -> debugging target
Recorded per episode: the spawn table, per-step actions, the stochastic *semantic* events
(dirt-respawn tiles, maintainer action per tick), a full state snapshot after every step,
rewards (f64), done and the packed observation tensor (cast to f32).

Usage:  python tests/golden/make_golden.py [--out tests/golden] [--only cfg1,cfg4]
"""
import argparse
import io
import json
import random
import sys
import contextlib
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path('/root/reference')
sys.dont_write_bytecode = True
sys.path.insert(0, str(HERE / 'stubs'))
sys.path.insert(1, str(REFERENCE))

from marl_factory_grid.environment.factory import Factory  # noqa: E402
from marl_factory_grid.environment.entity.object import Object  # noqa: E402
from marl_factory_grid.utils.states import Gamestate  # noqa: E402
from marl_factory_grid.modules.maintenance.entities import Maintainer  # noqa: E402
from marl_factory_grid.modules.clean_up.groups import DirtPiles  # noqa: E402

_ORIG_EQ, _ORIG_HASH = Object.__eq__, Object.__hash__

MAINT_CODE = {'North': 0, 'East': 1, 'South': 2, 'West': 3, 'NorthEast': 4, 'SouthEast': 5, 'SouthWest': 6,
              'NorthWest': 7, 'Noop': 8, 'DoorUse': 9, 'MachineAction': 10}
GONE = -9999
K_DIRT = 64           # padded dirt slots in the snapshot
K_RESPAWN = 8         # padded tiles per respawn event

EV = {'free': [], 'maint': [], 'in_dirt_spawn': False, 'dirt_spawn': []}

_free = Gamestate.get_n_random_free_positions


def _free_hook(self, n):
    r = _free(self, n)
    if EV['in_dirt_spawn']:
        EV['dirt_spawn'].append([(int(x), int(y)) for x, y in r])
    EV['free'].append((n, [(int(x), int(y)) for x, y in r]))
    return r


_tick = Maintainer.tick


def _tick_hook(self, state):
    r = _tick(self, state)
    EV['maint'].append((r.identifier, bool(r.validity)))
    return r


_dirt_spawn = DirtPiles.trigger_spawn


def _dirt_spawn_hook(self, state, *a, **kw):
    EV['in_dirt_spawn'] = True
    try:
        return _dirt_spawn(self, state, *a, **kw)
    finally:
        EV['in_dirt_spawn'] = False


Gamestate.get_n_random_free_positions = _free_hook
Maintainer.tick = _tick_hook
DirtPiles.trigger_spawn = _dirt_spawn_hook


def set_mode(mode):
    if mode == 'I':
        Object.__eq__ = lambda s, o: s is o
        Object.__hash__ = lambda s: id(s)
    else:
        Object.__eq__ = _ORIG_EQ
        Object.__hash__ = _ORIG_HASH


def group(state, name):
    # NB: Entities.__getitem__ on a missing key inserts None -> always test membership first
    if name not in state.entities.names:
        return []
    return list(state[name])


def is_listed(state, e):
    return any(x is e for x in state.entities.pos_dict[e.pos])


def snapshot(f):
    st = f.state
    s = {}
    agents = group(st, 'Agent')
    s['agent_pos'] = np.array([a.pos for a in agents], np.int16).reshape(len(agents), 2)
    doors = group(st, 'Doors')
    s['door_open'] = np.array([d.is_open for d in doors], np.uint8)
    s['door_timer'] = np.array([d.time_to_close for d in doors], np.int16)
    s['door_listed'] = np.array([is_listed(st, d) for d in doors], np.uint8)
    dirt = group(st, 'DirtPiles')
    assert len(dirt) <= K_DIRT
    dp = np.full((K_DIRT, 2), GONE, np.int16)
    da = np.zeros(K_DIRT, np.float64)
    du = np.full(K_DIRT, -1, np.int32)
    dl = np.zeros(K_DIRT, np.uint8)
    for i, d in enumerate(dirt):
        dp[i], da[i], du[i], dl[i] = d.pos, d.amount, d.u_int, is_listed(st, d)
    s['dirt_n'] = np.int32(len(dirt))
    s['dirt_pos'], s['dirt_amt'], s['dirt_uid'], s['dirt_listed'] = dp, da, du, dl
    for key, gname in (('item', 'Items'), ('pod', 'ChargePods'), ('dest', 'Destinations'),
                       ('drop', 'DropOffLocations'), ('machine', 'Machines'), ('maint', 'Maintainers')):
        g = group(st, gname)
        s[f'{key}_pos'] = np.array([e.pos for e in g], np.int16).reshape(len(g), 2)
        s[f'{key}_listed'] = np.array([is_listed(st, e) if e.pos[0] >= 0 else 0 for e in g], np.uint8)
    s['dest_reached'] = np.array([d.was_reached() for d in group(st, 'Destinations')], np.uint8)
    bats = group(st, 'Batteries')
    s['battery'] = np.array([b.charge_level for b in bats], np.float64)
    s['step'] = np.int32(st.curr_step)
    s['dirt_next_uid'] = np.int32(Object._u_idx['DirtPile'] if 'DirtPile' in Object._u_idx else 0)
    respawn = next((r for r in st.rules if r.__class__.__name__ == 'RespawnDirt'), None)
    s['dirt_next_spawn'] = np.int32(respawn._next_dirt_spawn if respawn is not None else -1)
    s['paralysed'] = np.array([bool(a.var_is_paralyzed) for a in agents], np.uint8)
    # arrival order of the agents sharing a tile (Agents.pos_dict list order; tie-break of the destination reward)
    apd = st['Agent'].pos_dict
    s['agent_rank'] = np.array([[x is a for x in apd[a.pos]].index(True) for a in agents], np.int16)
    return s


def pack_obs(obs_by_agent):
    return np.concatenate([np.asarray(o, np.float64) for o in obs_by_agent], axis=0).astype(np.float32)


def run_episode(cfg_path, seed, mode, max_steps, action_seed):
    set_mode(mode)
    Object._u_idx.clear()
    random.seed(seed)
    for k in ('free', 'maint', 'dirt_spawn'):
        EV[k] = []
    with contextlib.redirect_stdout(io.StringIO()):
        f = Factory(str(cfg_path))
        obs0 = f.reset()
    agents = list(f.state['Agent'])
    n_act = [len(a.actions) for a in agents]
    arng = np.random.default_rng(action_seed)

    snaps = [snapshot(f)]
    obs = [pack_obs(list(obs0.values()))]
    # initial dirt spawn event (tiles proposed + amounts are visible in the snapshot)
    init_dirt_spawn = EV['dirt_spawn'][:]
    EV['dirt_spawn'] = []
    EV['maint'] = []
    actions, rewards, dones, maint_act, maint_valid, resp_n, resp_tiles, infos = [], [], [], [], [], [], [], []
    n_maint = len(group(f.state, 'Maintainers'))
    for t in range(max_steps):
        a = [int(arng.integers(0, n)) for n in n_act]
        with contextlib.redirect_stdout(io.StringIO()):
            _, o, r, d, info = f.step(a)
        actions.append(a)
        infos.append({k: float(v) for k, v in info.items()})
        rewards.append(np.asarray(r, np.float64).reshape(-1))
        dones.append(bool(d))
        snaps.append(snapshot(f))
        obs.append(pack_obs(o))
        ma = np.full(n_maint, 8, np.int8)
        mv = np.zeros(n_maint, np.uint8)
        assert len(EV['maint']) == n_maint
        for i, (ident, val) in enumerate(EV['maint']):
            ma[i], mv[i] = MAINT_CODE[ident], val
        EV['maint'] = []
        maint_act.append(ma)
        maint_valid.append(mv)
        rt = np.full((K_RESPAWN, 2), GONE, np.int16)
        assert len(EV['dirt_spawn']) <= 1
        if EV['dirt_spawn']:
            tiles = EV['dirt_spawn'][0]
            resp_n.append(len(tiles))
            rt[:len(tiles)] = tiles
        else:
            resp_n.append(-1)
        EV['dirt_spawn'] = []
        resp_tiles.append(rt)
        if d:
            break
    ep = {
        'actions': np.array(actions, np.int32),
        'reward': np.array(rewards, np.float64),
        'done': np.array(dones, np.uint8),
        'maint_act': np.array(maint_act, np.int8).reshape(len(actions), n_maint),
        'maint_valid': np.array(maint_valid, np.uint8).reshape(len(actions), n_maint),
        'respawn_n': np.array(resp_n, np.int8),
        'respawn_tiles': np.array(resp_tiles, np.int16),
        'obs': np.stack(obs, 0),
        'door_pos': np.array([d.pos for d in group(f.state, 'Doors')], np.int16).reshape(-1, 2),
        'init_dirt_n_proposed': np.int32(len(init_dirt_spawn[0]) if init_dirt_spawn else 0),
        'info_json': np.frombuffer(json.dumps(infos).encode(), np.uint8),      # the reference's per-step `info` dicts
    }
    for key in snaps[0]:
        ep[key] = np.stack([s[key] for s in snaps], 0)
    meta = {
        'seed': seed, 'mode': mode, 'action_seed': action_seed,
        'agent_names': [a.name for a in agents],
        'n_actions': n_act,
        'action_names': [[x.name for x in a.actions] for a in agents],
        'named_action_space': f.named_action_space,
        'obs_layers': [list(f.obs_builder.obs_layers[a.name]) for a in agents],
        'level_shape': [int(x) for x in f.map.level_shape],
    }
    return ep, meta


# (config, [(seed, mode, max_steps)]) -- sizes chosen so that the whole fixture set stays small
PLAN = {
    'cfg1': [(0, 'U', 500), (1, 'U', 300), (2, 'I', 300)],
    'cfg2': [(0, 'U', 500), (1, 'U', 300), (2, 'I', 300)],
    'cfg3': [(0, 'U', 300), (1, 'U', 300), (3, 'I', 200)],
    'cfg4': [(0, 'U', 500), (1, 'U', 500), (2, 'U', 500), (3, 'U', 500), (4, 'I', 300), (5, 'I', 300)],
    'stress': [(s, 'U', 200) for s in range(8)] + [(s, 'I', 200) for s in range(8, 12)],
    'obs_test': [(0, 'U', 3), (1, 'I', 3)],
    'stress2': [(s, 'U', 150) for s in range(4)] + [(s, 'I', 150) for s in range(4, 6)],
    # the reference's own shipped scenarios, unmodified (configs/default_config.yaml, configs/clean_and_bring.yaml)
    'default_config': [(0, 'U', 150), (1, 'U', 150), (2, 'I', 100)],
    'clean_and_bring': [(0, 'U', 120), (1, 'U', 120), (2, 'I', 80)],
    # done rules no other fixture fires: WatchCollisions.done_at_collisions, DoneAtDestinationReach all / simultaneous
    'stress3': [(s, 'U', 120) for s in range(6)] + [(s, 'I', 120) for s in range(6, 8)],
    'dest_all': [(s, 'U', 400) for s in range(3)] + [(3, 'I', 400)],
    'dest_simul': [(s, 'U', 600) for s in range(3)] + [(3, 'I', 600)],
    # the reference's shipped full-observability scenarios: bound destinations, DoRandomInitialSteps, blocking agents
    'eight_puzzle': [(s, 'U', 200) for s in range(4)] + [(4, 'I', 100)],
    'narrow_corridor': [(s, 'U', 200) for s in range(3)] + [(3, 'I', 100)],
    'stress4': [(s, 'U', 120) for s in range(3)] + [(3, 'I', 120)],          # dict per_action_costs
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=str(HERE))
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    only = [x for x in args.only.split(',') if x]
    for cfg, eps in PLAN.items():
        if only and cfg not in only:
            continue
        cfg_path = REPO / 'marl_factory_grid_b200' / 'configs' / f'{cfg}.yaml'
        out, metas = {}, []
        for k, (seed, mode, max_steps) in enumerate(eps):
            ep, meta = run_episode(cfg_path, seed, mode, max_steps, action_seed=1000 + seed)
            for key, val in ep.items():
                out[f'ep{k}/{key}'] = val
            metas.append(meta)
            print(f'{cfg} ep{k} seed={seed} mode={mode}: T={len(ep["actions"])} done={bool(ep["done"][-1])} '
                  f'ret={ep["reward"].sum(0)}', flush=True)
        out['meta'] = np.frombuffer(json.dumps(metas).encode(), np.uint8)
        np.savez_compressed(Path(args.out) / f'{cfg}.npz', **out)


if __name__ == '__main__':
    main()
