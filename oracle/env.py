"""Single-environment CPU restatement of the reference step + observation semantics.

TEST INFRASTRUCTURE (see oracle/__init__.py).  One `OracleEnv` == one reference `Factory` that is
used for exactly one episode (fresh-Factory protocol).  The code is organised like
SURVEY.md App. F (F.1 step, F.2 uid listing, F.3 faithful observation); each block cites the
reference lines it restates.  Stochastic events (spawn table, dirt respawn tiles, maintainer
action) are INPUTS in replay mode; `oracle/freerun.py` supplies them from Philox for free-running.
"""
from __future__ import annotations

import numpy as np

from marl_factory_grid_b200 import spec as S
from .rays import full_rays

GONE = None
MOVES = S.DIR_DELTA


class Ent:
    """A positional entity with an integer uid (object.py:103-113: uid drives __eq__/__hash__)."""
    __slots__ = ('cls', 'uid', 'pos', 'listed', 'open', 'timer', 'amount', 'reached')

    def __init__(self, cls, uid, pos):
        self.cls, self.uid, self.pos, self.listed = cls, uid, pos, False
        self.open, self.timer, self.amount, self.reached = False, 0, 0.0, False

    def __repr__(self):
        return f'{self.cls}[{self.uid}]@{self.pos}{"" if self.listed else "(unlisted)"}'


class OracleEnv:
    def __init__(self, spec: S.EnvSpec, faithful: bool = True):
        self.spec = spec
        self.faithful = faithful
        self.A = spec.n_agents
        self.wall = spec.walls
        self.H, self.W = spec.H, spec.W
        # radius handed to RayCaster = min(obs_shape): the window DIAMETER for POMDP (defect B14), min(H, W) when
        # pomdp_r == 0 (full observability, observation_builder.py:51, 244)
        self.full_obs = spec.pomdp_r == 0
        self.rays = full_rays(min(spec.H, spec.W) if self.full_obs else spec.obs_d)
        # wall uid = row-major wall index (walls are created first, level_parser.py:77-78)
        self.wall_uid = {tuple(p): k for k, p in enumerate(np.argwhere(spec.walls).tolist())}
        self._rules = spec.rules
        respawn = spec.rule(S.R_RESPAWN_DIRT)
        self.dirt_next_spawn = int(respawn.params[0]) if respawn else -1   # clean_up/rules.py:47
        self.clear()

    # ------------------------------------------------------------------------------- state
    def clear(self):
        sp = self.spec
        self.step_no = 0
        self.apos = [(0, 0)] * self.A
        self.stamp = list(range(self.A))
        self.clock = self.A
        self.paralysed = [False] * self.A
        self.bat = [sp.battery_initial] * self.A
        self.doors = []
        self.dirt = []                 # creation order (dict insertion order of the DirtPiles collection)
        self.dirt_next_uid = 0
        self.items, self.pods, self.dests, self.drops, self.machines, self.maints = [], [], [], [], [], []
        self.slot = {}

    def _key(self, e):
        return (e.uid, e.pos) if self.faithful else (e.cls, e.uid, e.pos)

    def l_add(self, e):
        """objects.py:203-214 notify_add_entity on the global pos_dict."""
        k = self._key(e)
        if k not in self.slot:
            self.slot[k] = e
            e.listed = True
        else:
            e.listed = False

    def l_del(self, e):
        """objects.py:193-201 notify_del_entity: list.remove() drops the first EQUAL element."""
        other = self.slot.pop(self._key(e), None)
        if other is not None:
            other.listed = False

    def positional(self):
        return self.doors + self.dirt + self.items + self.pods + self.dests + self.drops + self.machines + self.maints

    # ------------------------------------------------------------------------------- queries (F.1 / F.2)
    def _door_at(self, p):
        for d in self.doors:
            if d.pos == p:
                return d
        return None

    def closed_listed_door(self, p):
        d = self._door_at(p)
        return d is not None and not d.open and d.listed

    def in_grid(self, p):
        return 0 <= p[0] < self.H and 0 <= p[1] < self.W

    def blocked(self, p):
        """states.py:259-270 check_pos_validity (negated)."""
        if not self.in_grid(p) or self.wall[p]:
            return True
        if self.closed_listed_door(p):
            return True
        return any(self.spec.agents[i].is_blocking_pos and self.apos[i] == p for i in range(self.A))

    def n_coll(self, p):
        """Collidable LISTED entities on p: agents, maintainers, closed doors, walls."""
        n = sum(1 for q in self.apos if q == p)
        n += sum(1 for m in self.maints if m.pos == p and m.listed)
        n += 1 if self.closed_listed_door(p) else 0
        n += 1 if self.in_grid(p) and self.wall[p] else 0
        return n

    def occupied(self, p):
        """global_entities.py:187-194 is_occupied."""
        return self.n_coll(p) >= 1 or self.blocked(p)

    def n_entities(self, p):
        """len(global pos_dict[p]): agents + every listed entity (doors/entitites.py:109)."""
        return sum(1 for q in self.apos if q == p) + sum(1 for e in self.positional() if e.pos == p and e.listed)

    def is_free(self, p):
        """global_entities.py:111-121 free_positions_generator predicate (floor tile, nothing collidable/blocking)."""
        return self.in_grid(p) and not self.wall[p] and self.n_coll(p) == 0 and not self.blocked(p)

    def toggle_near(self, p):
        """doors/actions.py:18-34 + global_entities.py:20-38: every LISTED door in the 3x3 block (floor tiles)."""
        valid = False
        for d in self.doors:
            if d.listed and abs(d.pos[0] - p[0]) <= 1 and abs(d.pos[1] - p[1]) <= 1:
                if d.open:
                    d.open = False
                else:
                    d.open, d.timer = True, S.DOOR_AUTO_CLOSE_INTERVAL
                valid = True
        return valid

    def try_move(self, p, d, mover_blocks):
        """actions.py:77-100 Move.do -> states.py:240-257 check_move_validity."""
        t = (p[0] + MOVES[d][0], p[1] + MOVES[d][1])
        ok = t != p and not self.blocked(t) and not (mover_blocks and self.occupied(t))
        return ok, t

    # ------------------------------------------------------------------------------- dirt helpers
    def _dirt_at(self, p):
        for d in self.dirt:
            if d.pos == p:
                return d
        return None

    def _set_dirt_amount(self, d, new):
        d.amount = min(new, S.DIRT_PILE_MAX)           # clean_up/entitites.py:34-38

    def dirt_spawn(self, tiles, amounts):
        """clean_up/groups.py:70-95 trigger_spawn body after the random draws."""
        for pos, a in zip(tiles, amounts):
            if not (sum(d.amount for d in self.dirt) > self.spec.dirt_max_global):
                d = self._dirt_at(pos)
                if d is not None:
                    self._set_dirt_amount(d, d.amount + a)
                else:
                    e = Ent('dirt', self.dirt_next_uid, pos)
                    self.dirt_next_uid += 1
                    e.amount = a
                    self.dirt.append(e)
                    self.l_add(e)
            else:
                return

    # ------------------------------------------------------------------------------- step (F.1)
    def step(self, actions, maint_actions=None, respawn_tiles=None):
        """factory.py:189-259 + states.py:170-226.  Returns (reward[A] f64, done bool).

        maint_actions: one tape code per maintainer (0..7 move dir, 8 noop, 9 door use, 10 machine action).
        respawn_tiles: proposed free tiles if RespawnDirt fires in this step (else ignored)."""
        sp = self.spec
        A = self.A
        self.step_no += 1
        rew = [0.0] * A
        glob = 0.0

        # ---- agents act sequentially against the live state (states.py:189-198)
        acted = [None] * A                      # index of the action each agent took this tick (None: paralysed, skipped)
        for i in range(A):
            if self.paralysed[i]:
                continue
            acted[i] = int(actions[i])
            act = sp.agents[i].actions[int(actions[i])]
            p = self.apos[i]
            op = act.opcode
            if op == S.OP_MOVE:
                ok, t = self.try_move(p, act.direction, sp.agents[i].is_blocking_pos)
                if ok:
                    self.apos[i] = t
                    self.stamp[i] = self.clock
                    self.clock += 1
            elif op == S.OP_NOOP:
                ok = True
            elif op == S.OP_DOORUSE:
                ok = self.toggle_near(p)
            elif op == S.OP_CLEAN:                      # clean_up/actions.py:19-36 (global index -> listed pile)
                d = self._dirt_at(p)
                ok = d is not None and d.listed
                if ok:
                    new = d.amount - sp.dirt_clean_amount
                    if new <= 0:
                        self.l_del(d)
                        self._remove_identity(self.dirt, d)
                    else:
                        self._set_dirt_amount(d, max(new, 0))
            elif op == S.OP_ITEM:                       # items/actions.py:41-63
                if any(x.pos == p for x in self.drops):
                    rew[i] += act.aux_reward            # inventory is always empty -> failed drop-off
                    continue
                it = next((x for x in self.items if x.pos == p), None)
                ok = it is not None
                if ok:
                    self.l_del(it)
                    it.pos, it.listed = GONE, False
            elif op == S.OP_CHARGE:                     # batteries/actions.py:20-31, entitites.py:98-111
                ok = False
                if any(x.pos == p for x in self.pods):
                    if not self.bat[i] >= 1.0 and not sum(1 for q in self.apos if q == p) > 1:
                        self.bat[i] = min(1, S.CHARGE_RATE + self.bat[i])
                        ok = True
            elif op == S.OP_DEST:                       # destinations/actions.py:17-24
                if any(x.pos == p for x in self.dests):
                    raise AttributeError("'list' object has no attribute 'do_wait_action'")   # reference raises
                ok = False
            elif op == S.OP_MACHINE:                    # machines/actions.py:19-25: maintain() on a healthy machine
                ok = any(x.pos == p for x in self.machines)
            else:
                raise ValueError(op)
            rew[i] += act.valid_reward if ok else act.fail_reward

        # ---- tick_step hooks in yaml order (states.py:56-61)
        mi = 0
        for r in self._rules:
            op, P = r.opcode, r.params
            if op == S.R_DOOR_AUTO_CLOSE:               # doors/entitites.py:108-122
                for d in self.doors:
                    if self.n_entities(d.pos) <= 2:
                        if d.open and d.timer:
                            d.timer -= 1
                        elif d.open and not d.timer:
                            d.open = False
                    else:
                        d.timer = S.DOOR_AUTO_CLOSE_INTERVAL
            elif op == S.R_MOVE_MAINTAINERS:            # maintenance/entities.py:37-136 via recorded action
                for k, m in enumerate(self.maints):
                    code = int(maint_actions[k])
                    if code < 8:
                        ok, t = self.try_move(m.pos, code, False)
                        if ok:
                            self.l_del(m)
                            m.pos = t
                            self.l_add(m)
                    elif code == S.MAINT_DOORUSE:
                        self.toggle_near(m.pos)
            elif op == S.R_RESPAWN_DIRT:                # clean_up/rules.py:49-59
                if self.dirt_next_spawn < 0:
                    pass
                elif self.dirt_next_spawn == 0:
                    n = int(P[1])
                    self.dirt_spawn(list(respawn_tiles), [P[2]] * n)
                    self.dirt_next_spawn = int(P[0])
                else:
                    self.dirt_next_spawn -= 1
            elif op in (S.R_BATTERY_DECHARGE, S.R_DONE_BATTERY):    # batteries/rules.py:50-63, entitites.py:60-69
                for i in range(A):
                    cost = P[0]
                    if P[5]:        # per_action_costs as a dict: class name of the action taken; a skipped (paralysed) agent: 'Noop'
                        costs = sp.act_costs[i]
                        cost = costs[-1] if acted[i] is None else costs[acted[i]]
                    if self.bat[i] != 0:
                        self.bat[i] = max(0, cost + self.bat[i])
            elif op in (S.R_DEST_REACH_REWARD, S.R_DONE_DEST):      # destinations/rules.py:34-54
                for k, d in enumerate(self.dests):
                    if not d.reached:
                        here = [i for i in range(A) if self.apos[i] == d.pos]
                        bound = sp.dest_bound[k] if k < len(sp.dest_bound) else -1
                        if here and bound >= 0 and bound not in here:
                            here = []           # a bound destination is only reached by its own agent (rules.py:40-47)
                        if here:
                            d.reached = True
                            last = max(here, key=lambda i: self.stamp[i])   # last in Agents.pos_dict list order
                            rew[last] += P[0]

        # ---- tick_post_step hooks (states.py:70-75)
        for r in self._rules:
            op, P = r.opcode, r.params
            if op == S.R_WATCH_COLLISIONS:              # environment/rules.py:276-306, states.py:228-238
                for i in range(A):
                    if self.n_coll(self.apos[i]) >= 2:
                        rew[i] += P[0]
            elif op in (S.R_BATTERY_DECHARGE, S.R_DONE_BATTERY):    # batteries/rules.py:66-87
                for i in range(A):
                    if self.bat[i] == 0:
                        rew[i] += P[1]
                        if P[2]:
                            self.paralysed[i] = True
                    if self.paralysed[i] and not self.bat[i] == 0:
                        self.paralysed[i] = False

        # ---- on_check_done hooks (states.py:216-226)
        done = False
        for r in self._rules:
            op, P = r.opcode, r.params
            if op == S.R_DONE_MAX_STEPS:
                done |= int(P[0]) <= self.step_no
            elif op == S.R_DONE_ALL_DIRT:
                if len(self.dirt) == 0 and self.step_no:
                    done = True
                    glob += P[0]
            elif op == S.R_DONE_BATTERY:
                if P[4] and any(b == 0 for b in self.bat):
                    done = True
                    glob += P[3]
            elif op == S.R_DONE_DEST:
                cond = int(P[1])
                reached = [d.reached for d in self.dests]
                if (cond == 0 and any(reached)) or (cond in (1, 2) and all(reached)):
                    done = True
                    glob += P[2]
                elif cond == 2:
                    for d in self.dests:
                        d.reached = False
            elif op == S.R_DONE_MAINT_COLLISION:        # maintenance/rules.py:32-40
                mpos = [m.pos for m in self.maints]
                for i in range(A):
                    if self.apos[i] in mpos:
                        done = True
                        rew[i] += -5
            elif op == S.R_WATCH_COLLISIONS and P[1]:   # environment/rules.py:308-325
                if any(self.n_coll(self.apos[i]) >= 2 for i in range(A)) or self._any_listed_collision():
                    done = True
                    glob += P[2]
        # ---- reward fold (factory.py:222-259)
        # individual rewards only: with `individual_rewards: false` the reference raises TypeError at factory.py:217
        # (`sum(reward)` of a float) on the first step, so there is no scalar fold to restate
        if not sp.individual_rewards:
            raise TypeError("'float' object is not iterable")
        return np.array([x + glob for x in rew], np.float64), bool(done)

    @staticmethod
    def _remove_identity(lst, e):
        for k, x in enumerate(lst):
            if x is e:
                del lst[k]
                return

    def _any_listed_collision(self):
        tiles = [m.pos for m in self.maints if m.listed]
        return any(self.n_coll(p) >= 2 for p in tiles)

    # ------------------------------------------------------------------------------- observation (F.3)
    def _blocks_light(self, p):
        return (self.in_grid(p) and bool(self.wall[p])) or self.closed_listed_door(p)

    def _entities_on(self, p, index):
        return index.get(p, ())

    def visible(self, a):
        """ray_caster.py:66-104 visible_entities + observation_builder.py:155 `set(visible)`.

        Returns the de-duplicated list [(key, tag, pos, encoding)] in first-visit order."""
        ax, ay = self.apos[a]
        # global pos_dict restricted to what can be seen: listed entities, agents, walls
        index = {}
        for e in self.positional():
            if e.listed and e.pos is not GONE:
                index.setdefault(e.pos, []).append(e)
        agents_at = {}
        for j, q in enumerate(self.apos):
            agents_at.setdefault(q, []).append(j)
        seen = {}
        for ray in self.rays:
            px, py = ax + ray[0][0], ay + ray[0][1]
            for ox, oy in ray:
                x, y = ax + ox, ay + oy
                cx, cy = x - px, y - py
                hits = self._blocks_light((x, y))
                diag = (cx != 0 and cy != 0) and self._blocks_light((x, y - cy)) and self._blocks_light((x - cx, y))
                if not diag:
                    p = (x, y)
                    if self.in_grid(p) and self.wall[p]:
                        key = self.wall_uid[p] if self.faithful else ('wall', p)
                        seen.setdefault(key, ('Walls', p, 1.0))
                    for j in agents_at.get(p, ()):
                        seen.setdefault(self.spec.agents[j].name, (S.G_AGENT0 + j, p, 1.0))
                    for e in index.get(p, ()):
                        key = e.uid if self.faithful else (e.cls, e.uid)
                        seen.setdefault(key, (e.cls, p, self._encoding(e)))
                if hits or diag:
                    break
                px, py = x, y
        return list(seen.values())

    @staticmethod
    def _encoding(e):
        if e.cls == 'door':
            return S.ENC_DOOR_OPEN if e.open else S.ENC_DOOR_CLOSED
        if e.cls == 'dirt':
            return e.amount
        if e.cls == 'machine':
            return S.ENC_MACHINE
        if e.cls == 'dest':
            return 0.0 if e.reached else 1.0
        return 1.0

    _CLS_GROUP = {'Walls': S.G_WALLS, 'door': S.G_DOORS, 'dirt': S.G_DIRT, 'item': S.G_ITEMS, 'drop': S.G_DROPOFF,
                  'pod': S.G_PODS, 'dest': S.G_DEST, 'machine': S.G_MACHINES, 'maint': S.G_MAINT}

    def observe_agent(self, a):
        """observation_builder.py:138-220.  Returns f64 [C_a, D, D]."""
        sp = self.spec
        r = sp.pomdp_r
        shape = (self.H, self.W) if self.full_obs else (sp.obs_d, sp.obs_d)
        ax, ay = self.apos[a]
        planes = {}
        for tag, (ex, ey), enc in self.visible(a):
            term = tag if isinstance(tag, int) else self._CLS_GROUP[tag]
            pl = planes.setdefault(term, np.zeros(shape))
            x, y = (ex, ey) if self.full_obs else (ex - ax + r, ey - ay + r)      # observation_builder.py:156-160
            if 0 <= x < shape[0] and 0 <= y < shape[1]:
                pl[x, y] += enc
        chans = sp.agents[a].channels
        obs = np.zeros((len(chans),) + shape)
        for c, ch in enumerate(chans):
            if ch.kind == S.CH_TERMS:
                if len(ch.terms) == 1 and not ch.name.startswith('Combined('):
                    if ch.terms[0] in planes:
                        obs[c] = planes[ch.terms[0]]
                else:
                    comb = [planes[t] for t in ch.terms if t in planes]
                    if comb:
                        obs[c] = np.sum(comb, axis=0)
            elif ch.kind == S.CH_BATTERY:
                obs[c].flat[0] = self.bat[a]
            elif ch.kind == S.CH_GLOBALPOS:         # environment/entity/util.py:56-66
                obs[c].flat[0] = ax / self.H
                obs[c].flat[1] = ay / self.W
        return obs

    def observe(self):
        """Packed [sum(C_a), D, D] f32, agents in index order (the engine's output layout)."""
        return np.concatenate([self.observe_agent(a) for a in range(self.A)], 0).astype(np.float32)

    # ------------------------------------------------------------------------------- snapshots
    def snapshot(self):
        """Same fields / layout as tests/golden/make_golden.py::snapshot (for direct comparison)."""
        K = 64
        s = {'agent_pos': np.array(self.apos, np.int16).reshape(self.A, 2),
             'door_open': np.array([d.open for d in self.doors], np.uint8),
             'door_timer': np.array([d.timer for d in self.doors], np.int16),
             'door_listed': np.array([d.listed for d in self.doors], np.uint8)}
        dp = np.full((K, 2), -9999, np.int16)
        da, du, dl = np.zeros(K), np.full(K, -1, np.int32), np.zeros(K, np.uint8)
        for i, d in enumerate(self.dirt):
            dp[i], da[i], du[i], dl[i] = d.pos, d.amount, d.uid, d.listed
        s.update(dirt_n=np.int32(len(self.dirt)), dirt_pos=dp, dirt_amt=da, dirt_uid=du, dirt_listed=dl)
        for key, lst in (('item', self.items), ('pod', self.pods), ('dest', self.dests), ('drop', self.drops),
                         ('machine', self.machines), ('maint', self.maints)):
            s[f'{key}_pos'] = np.array([e.pos if e.pos is not GONE else (-9999, -9999) for e in lst],
                                       np.int16).reshape(len(lst), 2)
            s[f'{key}_listed'] = np.array([e.listed for e in lst], np.uint8)
        s['dest_reached'] = np.array([d.reached for d in self.dests], np.uint8)
        s['battery'] = np.array(self.bat if self.spec.has_batteries else [], np.float64)
        s['step'] = np.int32(self.step_no)
        s['dirt_next_uid'] = np.int32(self.dirt_next_uid)
        s['dirt_next_spawn'] = np.int32(self.dirt_next_spawn)
        s['paralysed'] = np.array(self.paralysed, np.uint8)
        return s


def load_snapshot(env: OracleEnv, snap: dict, door_pos) -> OracleEnv:
    """Fill an OracleEnv from one golden snapshot (dict of arrays for a single time index)."""
    env.clear()
    sp = env.spec
    env.step_no = int(snap['step'])
    env.apos = [tuple(int(v) for v in p) for p in snap['agent_pos']]
    rank = snap.get('agent_rank', np.zeros(env.A))
    order = sorted(range(env.A), key=lambda i: (int(rank[i]), i))
    for k, i in enumerate(order):
        env.stamp[i] = k
    env.clock = env.A
    env.paralysed = [bool(x) for x in snap['paralysed']] if 'paralysed' in snap else [False] * env.A
    if sp.has_batteries:
        env.bat = [float(x) for x in snap['battery']]

    def mk(cls, uid, pos, listed):
        e = Ent(cls, uid, tuple(int(v) for v in pos))
        e.listed = bool(listed)
        if e.listed:
            env.slot[env._key(e)] = e
        return e

    for k, p in enumerate(door_pos):
        d = mk('door', k, p, snap['door_listed'][k])
        d.open, d.timer = bool(snap['door_open'][k]), int(snap['door_timer'][k])
        env.doors.append(d)
    for k in range(int(snap['dirt_n'])):
        d = mk('dirt', int(snap['dirt_uid'][k]), snap['dirt_pos'][k], snap['dirt_listed'][k])
        d.amount = float(snap['dirt_amt'][k])
        env.dirt.append(d)
    env.dirt_next_uid = int(snap['dirt_next_uid'])
    env.dirt_next_spawn = int(snap['dirt_next_spawn'])
    for key, lst in (('item', env.items), ('pod', env.pods), ('dest', env.dests), ('drop', env.drops),
                     ('machine', env.machines), ('maint', env.maints)):
        for k, p in enumerate(snap[f'{key}_pos']):
            if int(p[0]) < 0:
                e = Ent(key, k, GONE)
            else:
                e = mk(key, k, p, snap[f'{key}_listed'][k])
            lst.append(e)
    for k, d in enumerate(env.dests):
        d.reached = bool(snap['dest_reached'][k])
    return env
