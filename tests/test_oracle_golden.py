"""Pins the oracle (oracle/env.py) against traces of the UNMODIFIED reference (tests/golden/*.npz).

Every episode is replayed from its spawn table with the recorded actions and stochastic events; after
every step the complete state, reward, done flag and packed observation must equal the reference's:
integers / f64 state bit-exact, observations bit-exact after the f32 cast, rewards to 1e-12.
mode "U" episodes check faithful=True (untouched reference), mode "I" episodes faithful=False.
"""
import numpy as np
import pytest

from oracle import OracleEnv, load_snapshot
from golden_util import episode_ids, episodes, respawn_tiles_at, snap_at, spec_for

CMP_KEYS = ['agent_pos', 'door_open', 'door_timer', 'door_listed', 'dirt_n', 'dirt_pos', 'dirt_amt', 'dirt_uid',
            'dirt_listed', 'item_pos', 'item_listed', 'pod_listed', 'dest_listed', 'drop_listed', 'machine_listed',
            'maint_pos', 'maint_listed', 'dest_reached', 'battery', 'step', 'dirt_next_uid', 'dirt_next_spawn',
            'paralysed']


@pytest.mark.parametrize('cfg,k', episode_ids(include_oracle_only=True))
def test_oracle_replays_reference_episode(cfg, k):
    ep = episodes(cfg)[k]
    spec = spec_for(cfg)
    env = load_snapshot(OracleEnv(spec, faithful=ep['meta']['mode'] == 'U'), snap_at(ep, 0), ep['door_pos'])
    np.testing.assert_array_equal(env.observe(), ep['obs'][0], err_msg=f'{cfg} ep{k} obs after reset')
    T = len(ep['actions'])
    for t in range(T):
        rew, done = env.step(ep['actions'][t], ep['maint_act'][t], respawn_tiles_at(ep, t))
        got, want = env.snapshot(), snap_at(ep, t + 1)
        for key in CMP_KEYS:
            np.testing.assert_array_equal(got[key], want[key], err_msg=f'{cfg} ep{k} step {t + 1}: {key}')
        np.testing.assert_allclose(rew, ep['reward'][t], rtol=1e-12, atol=1e-12, err_msg=f'{cfg} ep{k} step {t + 1} reward')
        assert done == bool(ep['done'][t]), f'{cfg} ep{k} step {t + 1} done'
        np.testing.assert_array_equal(env.observe(), ep['obs'][t + 1], err_msg=f'{cfg} ep{k} step {t + 1} obs')
