"""The per-step `info` dict (reference: utils/results.py:42-52 folded at environment/factory.py:222-239).

The engine writes per-agent result flags (`mfg_bind_step_flags`); `factory.step_info` turns one env's row into the reference's
`"{agent}_{ActionClass}"` / `"{agent}_Collisions"` entries.  The goldens that carry `info_json` hold the UNMODIFIED reference's
own info dict of every step; the per-agent action and collision keys must agree exactly (values are config constants)."""
import json

import numpy as np
import pytest

from golden_util import episodes, snap_at, spec_for
from hostsim_util import HostSim, tape_respawn
from marl_factory_grid_b200.factory import step_info

INFO_CFGS = ['stress', 'stress2', 'stress3', 'cfg4']


def _info_ids():
    return [(cfg, k) for cfg in INFO_CFGS for k in range(len(episodes(cfg))) if 'info_json' in episodes(cfg)[k]]


def _ref_subset(es, info):
    """the keys of a reference info dict that step_info claims: every agent's action result and its Collisions entry"""
    act_names = {f'{ag.name}_{a.class_name}' for ag in es.agents for a in ag.actions}
    coll_names = {f'{ag.name}_Collisions' for ag in es.agents}        # (maintainers / the done rule report under other names)
    return {k: v for k, v in info.items() if k in act_names or k in coll_names}


def _compare(cfg, k, t, es, flags_row, actions, ref_info):
    got, want = step_info(es, flags_row, actions), _ref_subset(es, ref_info)
    assert set(got) == set(want), f'{cfg} ep{k} step {t + 1}: keys {sorted(set(got) ^ set(want))}'
    for key in want:
        assert got[key] == pytest.approx(want[key], abs=1e-12), f'{cfg} ep{k} step {t + 1}: {key}'


def test_fixtures_carry_reference_info():
    assert len(_info_ids()) >= 20


@pytest.mark.parametrize('cfg,k', _info_ids())
def test_step_info_matches_reference_host_build(cfg, k):
    ep, es = episodes(cfg)[k], spec_for(cfg)
    infos = json.loads(bytes(ep['info_json']).decode())
    sim = HostSim(es, n_envs=1, faithful=ep['meta']['mode'] == 'U')
    flags = sim.enable_step_flags()
    sim.load_snapshot(0, snap_at(ep, 0))
    for t in range(len(ep['actions'])):
        rn, rp = tape_respawn(ep, t)
        sim.step(ep['actions'][t][None], ep['maint_act'][t][None], ([rn], [rp]))
        _compare(cfg, k, t, es, flags[0], ep['actions'][t], infos[t])


@pytest.mark.gpu
@pytest.mark.parametrize('cfg', INFO_CFGS)
def test_step_info_matches_reference_cuda(cfg):
    """All episodes of a config side by side in one batch, through the C ABI (mfg_bind_step_flags + mfg_step)."""
    import torch
    from marl_factory_grid_b200.engine import Engine
    es = spec_for(cfg)
    for faithful in (True, False):
        eps = [(k, ep) for k, ep in enumerate(episodes(cfg)) if (ep['meta']['mode'] == 'U') == faithful and 'info_json' in ep]
        if not eps:
            continue
        n = len(eps)
        eng = Engine(es, n, device='cuda:0', faithful=faithful)
        flags = eng.enable_step_flags()
        for e, (_, ep) in enumerate(eps):
            eng.load_snapshot(e, snap_at(ep, 0))
        infos = [json.loads(bytes(ep['info_json']).decode()) for _, ep in eps]
        T = max(len(ep['actions']) for _, ep in eps)
        nm = eps[0][1]['maint_act'].shape[1]
        for t in range(T):
            live = [t < len(ep['actions']) for _, ep in eps]
            acts = np.stack([ep['actions'][t] if l else np.zeros(es.n_agents, np.int64) for (_, ep), l in zip(eps, live)])
            ma = np.stack([ep['maint_act'][t] if l else np.full(nm, 8, np.uint8) for (_, ep), l in zip(eps, live)])
            rs = [tape_respawn(ep, t) if l else (np.int8(0), np.zeros(8, np.uint16)) for (_, ep), l in zip(eps, live)]
            tape = dict(maint_action=torch.as_tensor(ma), respawn_n=torch.as_tensor(np.array([r[0] for r in rs], np.int8)),
                        respawn_pos=torch.as_tensor(np.stack([r[1] for r in rs])))
            eng.step(torch.as_tensor(acts.astype(np.int32)), tape=tape)
            fl = flags.cpu().numpy()
            for e, ((k, ep), l) in enumerate(zip(eps, live)):
                if l:
                    _compare(cfg, k, t, es, fl[e], ep['actions'][t], infos[e][t])
        eng.close()


@pytest.mark.gpu
def test_factory_info_recorder_and_getitem():
    """Un-batched drop-in shape: `info` carries the reference's keys, env['Agent'] / summarize_state / EnvRecorder work."""
    from golden_util import CONFIGS
    from marl_factory_grid_b200 import EnvRecorder, Factory
    env = EnvRecorder(Factory(CONFIGS / 'cfg4.yaml', seed=3))
    env.reset()
    names = [ag.name for ag in env.spec.agents]
    rng = np.random.default_rng(0)
    for _ in range(12):
        a = [int(rng.integers(0, n)) for n in env.spec.n_actions]
        _, obs, rew, done, info = env.step(a)
        assert {'step_reward', 'step'} <= set(info)
        assert sum(k.split('_')[0] in names for k in info) >= len(names) - 0      # one action entry per (non-paralysed) agent
        assert info['step_reward'] == pytest.approx(sum(rew))
    st = env.summarize_state(0)
    assert st['step'] == 12 and len(st['agents']) == len(names) and {'name', 'x', 'y', 'can_collide', 'valid', 'action'} <= set(st['agents'][0])
    pos = env['Agent']['pos'][0].cpu().numpy()
    assert [(a['x'], a['y']) for a in st['agents']] == [tuple(int(v) for v in p) for p in pos]
    assert env['DirtPiles']['amount'].shape[0] == 1 and env['Doors']['open'].dtype == __import__('torch').bool
    with pytest.raises(KeyError):
        env['Nonsense']
    assert len(env._curr_ep_recorder) == 12 and env._curr_ep_recorder[0]['episode'] == 1
