"""ctypes mirror of include/mfg_b200.h and the EnvSpec -> MfgSpec packer (host side only)."""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from . import spec as S
from .rays import full_ray_table
from .spec import EnvSpec

MAX_AGENTS, MAX_ACTIONS, MAX_DOORS, MAX_DIRT = 16, 32, 64, 64
MAX_RULES, MAX_CHANNELS, MAX_SMALL, MAX_GROUPS, MAX_FIXED = 32, 32, 32, 16, 32
N_TERMS = 9 + MAX_AGENTS
RULE_NPARAM = 6
MAX_RAYS, MAX_RAY_LEN = 128, 16
NO_POS = 0xFFFF
N_STATS = 32
ENV_BLOCK = 128          # envs per state block (blocked struct-of-arrays layout, include/mfg_b200.h)
RESPAWN_TAPE_W = 8

# spawn-program ids (MFG_SP_*)
SP = {'Doors': 0, 'DirtPiles': 1, 'Batteries': 2, 'ChargePods': 3, 'Destinations': 4, 'Items': 5, 'Inventories': 6,
      'DropOffLocations': 7, 'Machines': 8, 'Maintainers': 9, 'GlobalPositions': 10}

# statistics vector indices (MFG_ST_*)
ST_EPISODES, ST_STEPS, ST_DONE_MAX_STEPS, ST_DONE_ALL_DIRT, ST_DONE_BATTERY, ST_DONE_DEST, ST_DONE_MAINT, \
    ST_DONE_COLLISION, ST_COLLISIONS, ST_DIRT_OVERFLOW, ST_SPAWN_FAIL, ST_RETURN_SUM, ST_RETURN_AGENT0 = range(13)


FLAG_VALID, FLAG_SKIPPED, FLAG_COLLISION, FLAG_MOVE_COLLISION, FLAG_AUX_REWARD = 1, 2, 4, 8, 16        # MFG_FLAG_* (mfg_bind_step_flags)


class MfgSpec(C.Structure):
    _fields_ = [
        ('H', C.c_int32), ('W', C.c_int32), ('pomdp_r', C.c_int32), ('n_agents', C.c_int32),
        ('individual_rewards', C.c_int32), ('faithful', C.c_int32),
        ('n_floor', C.c_int32), ('n_doors', C.c_int32), ('n_walls', C.c_int32),
        ('has_dirt', C.c_int32), ('dirt_slots', C.c_int32), ('dirt_quantity', C.c_int32),
        ('has_batteries', C.c_int32), ('has_globalpos', C.c_int32),
        ('n_items', C.c_int32), ('n_dropoff', C.c_int32), ('n_pods', C.c_int32), ('n_dest', C.c_int32),
        ('n_machines', C.c_int32), ('n_maint', C.c_int32),
        ('n_rules', C.c_int32), ('n_groups', C.c_int32), ('n_rays', C.c_int32), ('reserved0', C.c_int32),
        ('dirt_initial_amount', C.c_double), ('dirt_clean_amount', C.c_double), ('dirt_max_global', C.c_double),
        ('dirt_n_var', C.c_double), ('dirt_amount_var', C.c_double), ('battery_initial', C.c_double),
        ('seed', C.c_uint64),
        ('n_actions', C.c_int32 * MAX_AGENTS), ('agent_blocking', C.c_int32 * MAX_AGENTS),
        ('agent_n_fixed', C.c_int32 * MAX_AGENTS), ('agent_fixed_pos', (C.c_uint16 * MAX_FIXED) * MAX_AGENTS),
        ('act_opcode', (C.c_int32 * MAX_ACTIONS) * MAX_AGENTS), ('act_dir', (C.c_int32 * MAX_ACTIONS) * MAX_AGENTS),
        ('act_valid', (C.c_double * MAX_ACTIONS) * MAX_AGENTS), ('act_fail', (C.c_double * MAX_ACTIONS) * MAX_AGENTS),
        ('act_aux', (C.c_double * MAX_ACTIONS) * MAX_AGENTS),
        ('n_channels', C.c_int32 * MAX_AGENTS), ('ch_offset', C.c_int32 * MAX_AGENTS),
        ('ch_kind', (C.c_int32 * MAX_CHANNELS) * MAX_AGENTS), ('term_chmask', (C.c_uint32 * N_TERMS) * MAX_AGENTS),
        ('rule_op', C.c_int32 * MAX_RULES), ('rule_param', (C.c_double * RULE_NPARAM) * MAX_RULES),
        ('group_id', C.c_int32 * MAX_GROUPS), ('group_quantity', C.c_int32 * MAX_GROUPS),
        ('group_n_fixed', C.c_int32 * MAX_GROUPS), ('group_fixed_pos', (C.c_uint16 * MAX_FIXED) * MAX_GROUPS),
        ('ray_len', C.c_int32 * MAX_RAYS), ('ray_dx', (C.c_int8 * MAX_RAY_LEN) * MAX_RAYS),
        ('ray_dy', (C.c_int8 * MAX_RAY_LEN) * MAX_RAYS),
        ('dest_mode', C.c_int32), ('random_initial_steps', C.c_int32),
        ('dest_bound', C.c_int32 * MAX_SMALL), ('dest_n_cand', C.c_int32 * MAX_SMALL),
        ('dest_cand', (C.c_uint16 * MAX_FIXED) * MAX_SMALL),
        ('act_cost', (C.c_double * (MAX_ACTIONS + 1)) * MAX_AGENTS),
        ('walls', C.c_void_p), ('floor_pos', C.c_void_p), ('door_pos', C.c_void_p), ('nexthop', C.c_void_p),
    ]


class MfgTape(C.Structure):
    _fields_ = [('d_maint_action', C.c_void_p), ('d_respawn_n', C.c_void_p), ('d_respawn_pos', C.c_void_p)]


class MfgField(C.Structure):
    _fields_ = [('offset', C.c_size_t), ('rows', C.c_int32), ('elem_size', C.c_int32), ('block_bytes', C.c_size_t)]


def pos16(xy) -> int:
    return (int(xy[0]) << 8) | int(xy[1])


def pos16_array(xy: np.ndarray) -> np.ndarray:
    xy = np.asarray(xy, np.int64).reshape(-1, 2)
    return ((xy[:, 0] << 8) | xy[:, 1]).astype(np.uint16)


def build_nexthop(es: EnvSpec) -> np.ndarray:
    """[F, F] uint8: direction (Move8 order) of the first step of a shortest path i -> j on the 8-connected
    non-wall graph (algorithms/static/utils.py:7-41 builds the same graph for the maintainers); 255 on the
    diagonal / when unreachable.  Tie-breaking between equally short routes is arbitrary in the reference
    (networkx over a shuffled adjacency), here it is the fixed BFS order of scipy."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import shortest_path
    F = es.n_floor
    index = -np.ones((es.H, es.W), np.int64)
    index[es.floor[:, 0], es.floor[:, 1]] = np.arange(F)
    rows, cols = [], []
    for d, (dx, dy) in enumerate(S.DIR_DELTA):
        nx, ny = es.floor[:, 0] + dx, es.floor[:, 1] + dy
        ok = (nx >= 0) & (ny >= 0) & (nx < es.H) & (ny < es.W)
        j = np.where(ok, index[np.clip(nx, 0, es.H - 1), np.clip(ny, 0, es.W - 1)], -1)
        m = j >= 0
        rows.append(np.arange(F)[m])
        cols.append(j[m])
    g = coo_matrix((np.ones(sum(len(r) for r in rows)), (np.concatenate(rows), np.concatenate(cols))), shape=(F, F))
    _, pred = shortest_path(g.tocsr(), method='D', unweighted=True, return_predecessors=True)
    # pred[j, i] = node before i on the path j -> i  ==  next node after i on the (reversed) path i -> j
    nxt = pred.T
    out = np.full((F, F), 255, np.uint8)
    ii, jj = np.nonzero(nxt >= 0)
    step = es.floor[nxt[ii, jj]] - es.floor[ii]
    lut = {delta: d for d, delta in enumerate(S.DIR_DELTA)}
    codes = np.array([lut[(int(a), int(b))] for a, b in step], np.uint8) if len(ii) else np.zeros(0, np.uint8)
    out[ii, jj] = codes
    return out


class PackedSpec:
    """MfgSpec + the numpy arrays it points to (kept alive for the lifetime of the object)."""

    def __init__(self, es: EnvSpec, faithful: bool = True, seed: int = None, with_nexthop: bool = None):
        s = MfgSpec()
        self.es = es
        s.H, s.W, s.pomdp_r, s.n_agents = es.H, es.W, es.pomdp_r, es.n_agents
        s.individual_rewards = int(es.individual_rewards)
        s.faithful = int(bool(faithful))
        s.n_floor, s.n_doors, s.n_walls = es.n_floor, es.n_doors, int(es.walls.sum())
        s.has_dirt, s.dirt_slots, s.dirt_quantity = int(es.has_dirt), es.dirt_slots, es.dirt_quantity
        s.has_batteries, s.has_globalpos = int(es.has_batteries), int(es.has_globalpos)
        s.n_items, s.n_dropoff, s.n_pods, s.n_dest = es.n_items, es.n_dropoff, es.n_pods, es.n_dest
        s.n_machines, s.n_maint = es.n_machines, es.n_maint
        s.dirt_initial_amount, s.dirt_clean_amount = es.dirt_initial_amount, es.dirt_clean_amount
        s.dirt_max_global, s.dirt_n_var, s.dirt_amount_var = es.dirt_max_global, es.dirt_n_var, es.dirt_amount_var
        s.battery_initial = es.battery_initial
        s.seed = int(es.env_seed if seed is None else seed) & 0xFFFFFFFFFFFFFFFF

        off = 0
        for i, ag in enumerate(es.agents):
            s.n_actions[i] = len(ag.actions)
            s.agent_blocking[i] = int(ag.is_blocking_pos)
            if len(ag.positions) > MAX_FIXED:
                raise ValueError(f'{ag.name}: more than {MAX_FIXED} fixed positions.')
            s.agent_n_fixed[i] = len(ag.positions)
            for j, p in enumerate(ag.positions):
                s.agent_fixed_pos[i][j] = pos16(p)
            for j, act in enumerate(ag.actions):
                s.act_opcode[i][j], s.act_dir[i][j] = act.opcode, act.direction
                s.act_valid[i][j], s.act_fail[i][j], s.act_aux[i][j] = act.valid_reward, act.fail_reward, act.aux_reward
            s.n_channels[i], s.ch_offset[i] = len(ag.channels), off
            off += len(ag.channels)
            for c, ch in enumerate(ag.channels):
                s.ch_kind[i][c] = ch.kind
                if ch.kind == S.CH_TERMS:
                    for t in ch.terms:
                        s.term_chmask[i][t] |= (1 << c)
        s.n_rules = len(es.rules)
        for r, rule in enumerate(es.rules):
            s.rule_op[r] = rule.opcode
            for k, v in enumerate(rule.params):
                s.rule_param[r][k] = float(v)
        s.n_groups = len(es.groups)
        for g, grp in enumerate(es.groups):
            s.group_id[g] = SP[grp.name]
            s.group_quantity[g] = grp.quantity
            if grp.coords:
                if len(grp.coords) > MAX_FIXED:
                    raise ValueError(f'{grp.name}: more than {MAX_FIXED} fixed coordinates.')
                s.group_n_fixed[g] = len(grp.coords)
                for j, p in enumerate(grp.coords):
                    s.group_fixed_pos[g][j] = pos16(p)
        s.dest_mode, s.random_initial_steps = es.dest_mode, es.random_initial_steps
        for k in range(MAX_SMALL):
            s.dest_bound[k] = -1
        for k, b in enumerate(es.dest_bound):
            s.dest_bound[k] = b
        for k, cands in enumerate(es.dest_cands):
            if len(cands) > MAX_FIXED:
                raise ValueError(f'Destinations: more than {MAX_FIXED} candidate positions for one agent.')
            s.dest_n_cand[k] = len(cands)
            for j, pxy in enumerate(cands):
                s.dest_cand[k][j] = pos16(pxy)
        for i, costs in enumerate(es.act_costs):
            for j, v in enumerate(costs):
                s.act_cost[i][j] = float(v)
        # ray radius = min(observation shape): the window diameter for POMDP, min(H, W) for full observability
        # (observation_builder.py:244)
        rays = full_ray_table(min(es.obs_shape))
        if len(rays) > MAX_RAYS or max(len(r) for r in rays) > MAX_RAY_LEN:
            raise ValueError('ray table exceeds the engine limits')
        s.n_rays = len(rays)
        for k, ray in enumerate(rays):
            s.ray_len[k] = len(ray)
            for j, (dx, dy) in enumerate(ray):
                s.ray_dx[k][j], s.ray_dy[k][j] = dx, dy

        self.walls = np.ascontiguousarray(es.walls.astype(np.uint8).reshape(-1))
        self.floor_pos = pos16_array(es.floor)
        self.door_pos = pos16_array(es.door_pos) if es.n_doors else np.zeros(1, np.uint16)
        if with_nexthop is None:
            with_nexthop = es.n_maint > 0
        self.nexthop = build_nexthop(es) if with_nexthop else None
        s.walls = self.walls.ctypes.data
        s.floor_pos = self.floor_pos.ctypes.data
        s.door_pos = self.door_pos.ctypes.data
        s.nexthop = self.nexthop.ctypes.data if self.nexthop is not None else None
        self.struct = s

    @property
    def ptr(self):
        return C.byref(self.struct)


FIELD_DTYPES = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}
# fields whose natural dtype is not the unsigned integer of their width
FIELD_VIEW = {'bat': np.float64, 'ep_ret': np.float64, 'dirt_amt': np.float64, 'dirt_next_spawn': np.int16}

STATE_FIELD_NAMES: List[str] = [
    'step', 'episode', 'clock', 'apos', 'astamp', 'aflag', 'finished', 'bat', 'ep_ret', 'door_open', 'door_listed', 'door_timer',
    'dirt_pos', 'dirt_amt', 'dirt_uid', 'dirt_listed', 'dirt_end', 'dirt_n', 'dirt_next_uid', 'dirt_next_spawn',
    'item_pos', 'pod_pos', 'dest_pos', 'drop_pos', 'mach_pos', 'maint_pos', 'item_listed', 'pod_listed', 'dest_listed',
    'drop_listed', 'mach_listed', 'maint_listed', 'dest_reached', 'maint_target', 'maint_rand', 'maint_remaining',
    'maint_last']
