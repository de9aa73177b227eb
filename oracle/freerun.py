"""Free-running driver for the oracle: supplies the stochastic events (spawn table, dirt respawn tiles,
maintainer routing) that replay mode takes from a tape.  TEST INFRASTRUCTURE / CPU BASELINE ONLY
(see oracle/__init__.py): `bench.py` times this on the host cores as the `cpu_baseline` and as the
`--impl reference` arm ("port" of the reference's CPU path).

Distribution-equivalent restatement of:
  * SpawnAgents.on_reset / SpawnEntity.on_reset          environment/rules.py:163-199
  * Collection.trigger_spawn / DirtPiles.trigger_spawn   groups/collection.py:102-130, clean_up/groups.py:70-95
  * Maintainer.tick / get_move_action / _predict_move    modules/maintenance/entities.py:37-136
Random free tiles are drawn as in the reference: shuffle the floor list, keep the first n free ones.
"""
from __future__ import annotations

import numpy as np

from marl_factory_grid_b200 import spec as S
from marl_factory_grid_b200.abi import build_nexthop
from .env import Ent, OracleEnv

_NEXTHOP_CACHE = {}


def _nexthop(spec):
    key = id(spec)
    if key not in _NEXTHOP_CACHE:
        index = {tuple(p): i for i, p in enumerate(spec.floor.tolist())}
        _NEXTHOP_CACHE[key] = (build_nexthop(spec), index)
    return _NEXTHOP_CACHE[key]


class FreeRunEnv(OracleEnv):
    def __init__(self, spec: S.EnvSpec, faithful: bool = True, seed: int = 0):
        super().__init__(spec, faithful)
        self.rng = np.random.default_rng(seed)
        self.floor = [tuple(p) for p in spec.floor.tolist()]
        self._maint_state = []
        if spec.n_maint:
            self.nexthop, self.floor_index = _nexthop(spec)

    # ------------------------------------------------------------------ random tiles
    def _shuffled_floor(self):
        order = self.rng.permutation(len(self.floor))
        return (self.floor[i] for i in order)

    def n_free(self, n):
        out = []
        for p in self._shuffled_floor():
            if len(out) == n:
                break
            if self.is_free(p):
                out.append(p)
        return out

    # ------------------------------------------------------------------ reset
    def reset(self):
        sp = self.spec
        respawn = sp.rule(S.R_RESPAWN_DIRT)
        self.clear()
        self.dirt_next_spawn = int(respawn.params[0]) if respawn else -1
        self.apos = [None] * self.A
        for k, p in enumerate(sp.door_pos.tolist()):
            d = Ent('door', k, tuple(p))
            d.timer = S.DOOR_AUTO_CLOSE_INTERVAL
            self.doors.append(d)
            self.l_add(d)
        door_tiles = {d.pos for d in self.doors}
        for i, ag in enumerate(sp.agents):
            taken = set(q for q in self.apos if q is not None) | door_tiles
            if ag.positions:
                p = next((tuple(q) for q in ag.positions if tuple(q) not in taken), None)
                if p is None:
                    raise ValueError(f'It was not possible to spawn an Agent on the available position: {ag.positions}')
            else:
                p = next(q for q in self._shuffled_floor() if q not in taken)
            self.apos[i] = p
        lists = {'ChargePods': ('pod', self.pods), 'Destinations': ('dest', self.dests), 'Items': ('item', self.items),
                 'DropOffLocations': ('drop', self.drops), 'Machines': ('machine', self.machines),
                 'Maintainers': ('maint', self.maints)}
        for g in sp.groups:
            if g.name == 'DirtPiles':
                q = sp.dirt_quantity
                n_new = int(abs(q + self.rng.uniform(-sp.dirt_n_var, sp.dirt_n_var)))
                tiles = self.n_free(n_new)
                amounts = [sp.dirt_initial_amount + self.rng.uniform(-sp.dirt_amount_var, sp.dirt_amount_var)
                           for _ in range(q)]
                self.dirt_spawn(tiles, amounts)
            elif g.name == 'Destinations' and sp.dest_mode != S.DEST_FREE:
                # bound destinations (modules/destinations/rules.py:95-162)
                for k in range(sp.n_dest):
                    ag = sp.dest_bound[k]
                    if sp.dest_mode == S.DEST_ON_AGENT:
                        p = self.apos[ag]
                    else:
                        cands = [tuple(c) for c in sp.dest_cands[k]] or list(self.floor)
                        cands = [c for c in cands if c != self.apos[ag] and all(d.pos != c for d in self.dests)]
                        if not cands:
                            raise SystemExit(f'Could not spawn Destinations at: {sp.dest_cands[k]}')
                        p = cands[int(self.rng.integers(len(cands)))]
                    e = Ent('dest', k, p)
                    self.dests.append(e)
                    self.l_add(e)
            elif g.name in lists:
                cls, lst = lists[g.name]
                tiles = [tuple(c) for c in g.coords] if g.coords else self.n_free(g.quantity)
                for k, p in enumerate(tiles):
                    e = Ent(cls, k, p)
                    lst.append(e)
                    self.l_add(e)
        self._maint_state = [dict(path_target=None, nxt=[], last_serviced=None) for _ in self.maints]
        # DoRandomInitialSteps.on_reset_post_spawn (environment/rules.py:341-355)
        for _ in range(sp.random_initial_steps):
            free = self.n_free(1)
            if not free:
                break
            fp = free[0]
            nbs = [(fp[0] + dx, fp[1] + dy) for dx, dy in S.DIR_DELTA[:4]]
            nbs = [q for q in nbs if self.in_grid(q) and not self.wall[q]]
            if not nbs:
                continue
            frm = nbs[int(self.rng.integers(len(nbs)))]
            here = [i for i in range(self.A) if self.apos[i] == frm]
            if not here:
                continue                        # (the reference asserts an agent stands there)
            who = min(here, key=lambda i: self.stamp[i])
            if self.blocked(fp) or (sp.agents[who].is_blocking_pos and self.n_coll(fp) >= 1):
                continue
            self.apos[who] = fp
            self.stamp[who] = self.clock
            self.clock += 1
        return self.observe()

    # ------------------------------------------------------------------ maintainer policy
    def _maint_action(self, k):
        m, stt = self.maints[k], self._maint_state[k]
        here = next((i for i, x in enumerate(self.machines) if x.pos == m.pos), None)
        if here is not None and here != stt['last_serviced']:
            stt['last_serviced'] = here
            return S.MAINT_MACHINE
        if stt['path_target'] is None or stt['path_target'] == m.pos:
            for attempt in range(2):
                if not stt['nxt']:
                    if attempt == 1:
                        break
                    free = self.n_free(1)
                    cand = [x.pos for x in self.machines] + ([free[0]] if free else [])
                    self.rng.shuffle(cand)
                    stt['nxt'] = [tuple(c) for c in cand]
                stt['path_target'] = stt['nxt'].pop()
                if stt['path_target'] != m.pos:
                    break
            if stt['path_target'] is None or stt['path_target'] == m.pos:
                return S.MAINT_NOOP
        d = int(self.nexthop[self.floor_index[m.pos], self.floor_index[stt['path_target']]])
        if d > 7:
            return S.MAINT_NOOP
        nxt = (m.pos[0] + S.DIR_DELTA[d][0], m.pos[1] + S.DIR_DELTA[d][1])
        door = self._door_at(nxt)
        if door is not None and not door.open:
            return S.MAINT_DOORUSE
        if self.n_coll(nxt) > 0:
            return S.MAINT_NOOP
        return d

    # ------------------------------------------------------------------ step
    def step_free(self, actions):
        sp = self.spec
        # events are generated lazily against the state they are consumed in: hand the base class callables
        maint = _LazyMaint(self) if sp.n_maint else None
        respawn = None
        rule = sp.rule(S.R_RESPAWN_DIRT)
        if rule is not None and self.dirt_next_spawn == 0:
            respawn = _LazyRespawn(self, int(rule.params[1]))
        return self.step(actions, maint, respawn)


class _LazyMaint:
    """Indexable stand-in for the maintainer tape: the action is decided when MoveMaintainers runs."""

    def __init__(self, env):
        self.env = env

    def __getitem__(self, k):
        return self.env._maint_action(k)


class _LazyRespawn:
    """Iterable stand-in for the respawn tile list: tiles are drawn when RespawnDirt fires."""

    def __init__(self, env, n):
        self.env, self.n = env, n

    def __iter__(self):
        sp = self.env.spec
        n_new = int(abs(self.n + self.env.rng.uniform(-sp.dirt_n_var, sp.dirt_n_var)))
        return iter(self.env.n_free(n_new))


_WORKER = {}


def worker_init(cfg_path, faithful, seed_base):
    """Pool initializer: one persistent env per CPU worker (spec compile and next-hop table built once)."""
    import os
    from marl_factory_grid_b200 import FactoryConfigParser
    spec = FactoryConfigParser(cfg_path).compile()
    seed = seed_base + os.getpid()
    env = FreeRunEnv(spec, faithful=faithful, seed=seed)
    env.reset()
    _WORKER.update(env=env, spec=spec, arng=np.random.default_rng(seed + 1))


def worker_run(n_steps):
    """Advance this worker's env by n_steps (uniform random actions, observation built every step, in-place
    reset on done).  Returns (agent_steps, seconds)."""
    import time
    env, spec, arng = _WORKER['env'], _WORKER['spec'], _WORKER['arng']
    n_act = spec.n_actions
    t0 = time.perf_counter()
    for _ in range(n_steps):
        _, done = env.step_free([int(arng.integers(0, k)) for k in n_act])
        env.observe()
        if done:
            env.reset()
    return n_steps * spec.n_agents, time.perf_counter() - t0
