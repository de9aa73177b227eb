"""Helpers shared by the parity tests: load the committed reference traces (tests/golden/*.npz)."""
import json
from functools import lru_cache
from pathlib import Path

import numpy as np

from marl_factory_grid_b200 import FactoryConfigParser

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / 'tests' / 'golden'
CONFIGS = ROOT / 'marl_factory_grid_b200' / 'configs'
ALL_CFGS = ['cfg1', 'cfg2', 'cfg3', 'cfg4', 'stress', 'stress2', 'default_config', 'clean_and_bring', 'stress3', 'dest_all',
            'dest_simul', 'stress4']      # POMDP configs: oracle, direct AND tiled observation kernels
FULL_OBS_CFGS = ['obs_test', 'eight_puzzle', 'narrow_corridor']                                          # pomdp_r = 0 (full observability): oracle and the direct kernel
ORACLE_ONLY_CFGS = []

SNAP_KEYS = ['agent_pos', 'door_open', 'door_timer', 'door_listed', 'dirt_n', 'dirt_pos', 'dirt_amt', 'dirt_uid',
             'dirt_listed', 'item_pos', 'item_listed', 'pod_pos', 'pod_listed', 'dest_pos', 'dest_listed', 'drop_pos',
             'drop_listed', 'machine_pos', 'machine_listed', 'maint_pos', 'maint_listed', 'dest_reached', 'battery',
             'step', 'dirt_next_uid', 'dirt_next_spawn', 'paralysed', 'agent_rank']


@lru_cache(maxsize=None)
def spec_for(cfg, dirt_slots=40):
    return FactoryConfigParser(CONFIGS / f'{cfg}.yaml').compile(dirt_slots=dirt_slots)


@lru_cache(maxsize=None)
def _load(cfg):
    z = np.load(GOLDEN / f'{cfg}.npz')
    metas = json.loads(bytes(z['meta']).decode())
    eps = []
    for k, meta in enumerate(metas):
        ep = {key.split('/', 1)[1]: z[key] for key in z.files if key.startswith(f'ep{k}/')}
        ep['meta'] = meta
        eps.append(ep)
    return eps


def episodes(cfg):
    return _load(cfg)


def episode_ids(include_oracle_only=False):
    cfgs = ALL_CFGS + FULL_OBS_CFGS + (ORACLE_ONLY_CFGS if include_oracle_only else [])
    return [(cfg, k) for cfg in cfgs for k in range(len(_load(cfg)))]


def snap_at(ep, t):
    return {k: ep[k][t] for k in SNAP_KEYS if k in ep}


def respawn_tiles_at(ep, t):
    n = int(ep['respawn_n'][t])
    if n < 0:
        return None
    return [tuple(int(v) for v in p) for p in ep['respawn_tiles'][t][:n]]
