"""Free-running mode against the UNMODIFIED reference, as distributions.

The engine's Philox spawn, dirt re-spawn draws and maintainer routing cannot be replayed draw by draw against the reference
(CPython `random`, PCG64, networkx tie-breaking), so they are pinned statistically: tests/golden/make_freerun_stats.py ran the
reference free (uniform random actions; 2 400 episodes of cfg1, 400 of cfg4) and committed tests/golden/freerun_stats.json; here
the device code (host build on CPU, CUDA kernels under `-m gpu`) runs one complete episode per env and must agree on

  * spawn marginals: agents never start on door tiles, group members on distinct tiles, initial dirt count 9 / 10 at p = 1/2,
    agent start tiles uniform over the empty tiles (chi-square)
  * episode statistics: mean length, length histogram, done-reason shares, mean per-agent return

within 4 standard errors of the two samples (plus a small absolute slack, written next to each check)."""
import json
from collections import Counter

import numpy as np
import pytest

from freerun_stats_util import episode_stats, spawn_stats
from golden_util import GOLDEN, spec_for
from hostsim_util import HostSim

REF = json.loads((GOLDEN / 'freerun_stats.json').read_text())


def _check_spawn(cfg, es, sp, n):
    ref = REF[cfg]
    assert sp['agent_on_door'] == 0 and ref['agent_on_door'] == 0                    # SpawnAgents: EMPTY tiles only (rules.py:182-199)
    assert sp['group_duplicate_tiles'] == 0 and ref['group_duplicate_tiles'] == 0    # trigger_spawn: n distinct free tiles
    if es.has_dirt:
        q = es.dirt_quantity
        assert set(sp['dirt_n0']) <= {str(q - 1), str(q)} and set(ref['dirt_n0']) <= {str(q - 1), str(q)}
        p_eng = sp['dirt_n0'][str(q)] / n
        assert abs(p_eng - 0.5) < 4 * np.sqrt(0.25 / n)                              # int(abs(q + U(-.2, .2))): q - 1 or q, p = 1/2
        # the reference's workers share one PCG64 stream (env_seed), so its effective sample is episodes / 8 workers
        n_eff = ref['episodes'] / 8
        assert abs(ref['dirt_n0'][str(q)] / ref['episodes'] - 0.5) < 4 * np.sqrt(0.25 / n_eff)
    # agent 0 start tile: uniform over the empty tiles = floor minus door tiles (both samples)
    doors = {tuple(int(v) for v in d) for d in es.door_pos}
    empty = [tuple(int(v) for v in p) for p in es.floor if tuple(int(v) for v in p) not in doors]
    for name, counts, total in (('engine', sp['agent0_tiles'], n), ('reference', Counter(ref['agent0_tiles']), ref['episodes'])):
        assert set(counts) <= {str(t) for t in empty}, name
        expected = total / len(empty)
        chi2 = sum((counts.get(str(t), 0) - expected) ** 2 / expected for t in empty)
        dof = len(empty) - 1
        assert chi2 < dof + 5 * np.sqrt(2 * dof), f'{name}: start tiles are not uniform (chi2 {chi2:.0f}, dof {dof})'


def _check_episodes(cfg, st):
    ref = REF[cfg]
    n_e, n_r = st['episodes'], ref['episodes']
    assert abs(st['length_mean'] - ref['length_mean']) <= 30.0          # ~4 SE of the reference sample (std ~150 steps, 400 episodes)
    h_e, h_r = np.array(st['length_hist']) / n_e, np.array(ref['length_hist']) / n_r
    for b, (pe, pr) in enumerate(zip(h_e, h_r)):
        p = (pe * n_e + pr * n_r) / (n_e + n_r)
        assert abs(pe - pr) <= 4 * np.sqrt(p * (1 - p) * (1 / n_e + 1 / n_r)) + 0.01, f'length histogram bin {b}: {pe:.3f} vs {pr:.3f}'
    for reason in set(st['done_reasons']) | set(ref['done_reasons']):
        pe, pr = st['done_reasons'].get(reason, 0) / n_e, ref['done_reasons'].get(reason, 0) / n_r
        p = (pe * n_e + pr * n_r) / (n_e + n_r)
        assert abs(pe - pr) <= 4 * np.sqrt(p * (1 - p) * (1 / n_e + 1 / n_r)) + 0.01, f'done reason {reason}: {pe:.3f} vs {pr:.3f}'
    for a, (me, mr) in enumerate(zip(st['return_mean'], ref['return_mean'])):
        se = np.sqrt(st['return_std'][a] ** 2 / n_e + ref['return_std'][a] ** 2 / n_r)
        assert abs(me - mr) <= 4 * se + 0.05, f'mean return of agent {a}: {me:.3f} vs {mr:.3f} (se {se:.3f})'


@pytest.mark.parametrize('faithful', [True, False])
@pytest.mark.parametrize('cfg,n_envs', [('cfg1', 384), ('cfg4', 1024)])
def test_freerun_statistics_match_the_reference_host_build(cfg, n_envs, faithful):
    es = spec_for(cfg)
    sim = HostSim(es, n_envs, faithful=faithful, seed=123)
    flags = sim.enable_step_flags()
    sim.reset()

    def step_fn(a):
        r, d = sim.step(a, auto_reset=False)
        return r.copy(), d.copy(), flags.copy()

    _check_episodes(cfg, episode_stats(es, n_envs, 501, np.random.default_rng(1), step_fn))


@pytest.mark.parametrize('cfg', ['cfg1', 'cfg4'])
def test_spawn_marginals_match_the_reference_host_build(cfg):
    es = spec_for(cfg)
    n = 8192
    _check_spawn(cfg, es, spawn_stats(lambda k: HostSim(es, k, faithful=True, seed=5), es, n), n)


@pytest.mark.gpu
@pytest.mark.parametrize('faithful', [True, False])
@pytest.mark.parametrize('cfg,n_envs', [('cfg1', 4096), ('cfg4', 16384)])
def test_freerun_statistics_match_the_reference_cuda(cfg, n_envs, faithful):
    """The same comparison through the C ABI on the GPU, with a sample large enough that the reference's standard error dominates."""
    import torch
    from marl_factory_grid_b200.engine import Engine
    es = spec_for(cfg)
    eng = Engine(es, n_envs, device='cuda:0', faithful=faithful, seed=321)
    flags = eng.enable_step_flags()
    eng.reset()
    n = n_envs
    if True:
        f = eng.fields_numpy()
        doors = {(int(x) << 8) | int(y) for x, y in es.door_pos}
        sp = {'agent_on_door': int(sum(int(p) in doors for p in f['apos'].reshape(-1))),
              'agent0_tiles': Counter(str((int(p) >> 8, int(p) & 255)) for p in f['apos'][0]),
              'dirt_n0': Counter(str(int(v)) for v in f['dirt_n'][0]), 'group_duplicate_tiles': 0}
        for name in ('item_pos', 'pod_pos', 'dest_pos', 'drop_pos', 'mach_pos', 'maint_pos'):
            if name in f:
                col = f[name]
                sp['group_duplicate_tiles'] += int(sum(len(set(col[:, e])) != col.shape[0] for e in range(0, col.shape[1], 16)))
        _check_spawn(cfg, es, sp, n)

    def step_fn(a):
        r, d = eng.step(torch.as_tensor(a), auto_reset=False)
        return r.cpu().numpy(), d.cpu().numpy(), flags.cpu().numpy()

    _check_episodes(cfg, episode_stats(es, n_envs, 501, np.random.default_rng(2), step_fn))
    eng.close()
