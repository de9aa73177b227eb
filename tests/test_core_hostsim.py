"""Replays the reference traces through the engine's per-environment device code, compiled for the host
(tests/hostsim).  Same bar as the oracle test: state / done bit-exact, observations bit-exact after the
f32 cast, rewards within f32 rounding of the reference's f64 value (the engine emits f32 rewards)."""
import numpy as np
import pytest

from golden_util import episode_ids, episodes, snap_at, spec_for
from hostsim_util import HostSim, tape_respawn

CMP_KEYS = ['agent_pos', 'door_open', 'door_timer', 'door_listed', 'dirt_n', 'dirt_pos', 'dirt_amt', 'dirt_uid',
            'dirt_listed', 'item_pos', 'item_listed', 'pod_listed', 'dest_listed', 'drop_listed', 'machine_listed',
            'maint_pos', 'maint_listed', 'dest_reached', 'battery', 'step', 'dirt_next_uid', 'dirt_next_spawn',
            'paralysed']


@pytest.mark.parametrize('cfg,k', episode_ids())
def test_device_code_replays_reference_episode(cfg, k):
    ep = episodes(cfg)[k]
    es = spec_for(cfg)
    sim = HostSim(es, n_envs=1, faithful=ep['meta']['mode'] == 'U')
    sim.load_snapshot(0, snap_at(ep, 0))
    np.testing.assert_array_equal(sim.observe()[0], ep['obs'][0], err_msg=f'{cfg} ep{k} obs after reset')
    for t in range(len(ep['actions'])):
        rn, rp = tape_respawn(ep, t)
        rew, done = sim.step(ep['actions'][t][None], ep['maint_act'][t][None], ([rn], [rp]))
        got, want = sim.snapshot(0), snap_at(ep, t + 1)
        for key in CMP_KEYS:
            np.testing.assert_array_equal(got[key], want[key], err_msg=f'{cfg} ep{k} step {t + 1}: {key}')
        np.testing.assert_allclose(rew[0], ep['reward'][t], rtol=1e-6, atol=1e-7, err_msg=f'{cfg} ep{k} step {t + 1} reward')
        assert bool(done[0]) == bool(ep['done'][t]), f'{cfg} ep{k} step {t + 1} done'
        np.testing.assert_array_equal(sim.observe()[0], ep['obs'][t + 1], err_msg=f'{cfg} ep{k} step {t + 1} obs')
