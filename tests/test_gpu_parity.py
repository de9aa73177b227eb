"""GPU parity tests (run with `-m gpu` on the B200 box).  Everything goes through the C ABI
(libmfg_b200.so via marl_factory_grid_b200.engine); the checkers are the committed reference traces
(tests/golden), the oracle and the test-only host build of the device code.

Bars: integer state, door timers, f64 dirt amounts / battery levels, done flags and observations are
compared bit-exactly; rewards (emitted as f32) within rtol 1e-6 of the reference's f64 value.
"""
import numpy as np
import pytest

from golden_util import ALL_CFGS, FULL_OBS_CFGS, episodes, snap_at, spec_for
from hostsim_util import HostSim, tape_respawn

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

CMP_KEYS = ['agent_pos', 'door_open', 'door_timer', 'door_listed', 'dirt_n', 'dirt_pos', 'dirt_amt', 'dirt_uid',
            'dirt_listed', 'item_pos', 'item_listed', 'pod_listed', 'dest_listed', 'drop_listed', 'machine_listed',
            'maint_pos', 'maint_listed', 'dest_reached', 'battery', 'step', 'dirt_next_uid', 'dirt_next_spawn',
            'paralysed']


def _engine(es, n, **kw):
    from marl_factory_grid_b200.engine import Engine
    return Engine(es, n, device='cuda:0', **kw)


def _replay_batch(cfg, mode, obs_kernels, fused=False):
    """All episodes of (cfg, mode) as ONE batch: env i replays episode i; finished episodes idle.

    fused=True drives the ONE-call form `mfg_step_observe` with the tape and auto_reset on (k_step, packed re-spawn +
    list-mode observation on the side stream, tiled observation kernel on the caller's stream): an env is compared until
    the step that ends its episode (reward / done included); after that it is re-spawned from Philox and ignored."""
    es = spec_for(cfg)
    eps = [ep for ep in episodes(cfg) if ep['meta']['mode'] == mode]
    if not eps:
        pytest.skip(f'no {mode} episodes for {cfg}')
    N, A, NM = len(eps), es.n_agents, es.n_maint
    eng = _engine(es, N, faithful=mode == 'U')
    for i, ep in enumerate(eps):
        eng.load_snapshot(i, snap_at(ep, 0))
    T = [len(ep['actions']) for ep in eps]

    def check_obs(t):
        for kern in obs_kernels:
            eng.set_option('obs_kernel', kern)
            obs = eng.observe().cpu().numpy()
            for i, ep in enumerate(eps):
                if t <= T[i]:
                    np.testing.assert_array_equal(obs[i], ep['obs'][t], err_msg=f'{cfg}/{mode} env{i} t={t} obs kernel {kern}')

    check_obs(0)
    for t in range(max(T)):
        acts = np.zeros((N, A), np.int32)
        ma = np.full((N, max(NM, 1)), 8, np.uint8)
        rn = np.zeros(N, np.int8)
        rp = np.zeros((N, 8), np.uint16)
        for i, ep in enumerate(eps):
            if t < T[i]:
                acts[i] = ep['actions'][t]
                if NM:
                    ma[i, :NM] = ep['maint_act'][t]
                rn[i], rp[i] = tape_respawn(ep, t)
        tape = dict(maint_action=ma[:, :NM] if NM else None, respawn_n=rn, respawn_pos=rp)
        if fused:
            obs, rew, done = eng.step_observe(acts, tape=tape, auto_reset=True)
            obs = obs.cpu().numpy()
        else:
            rew, done = eng.step(acts, tape=tape)
        rew, done = rew.cpu().numpy(), done.cpu().numpy()
        fields = eng.fields_numpy()
        for i, ep in enumerate(eps):
            if t >= T[i]:
                continue
            np.testing.assert_allclose(rew[i], ep['reward'][t], rtol=1e-6, atol=1e-7, err_msg=f'{cfg}/{mode} env{i} step {t + 1}')
            assert bool(done[i]) == bool(ep['done'][t]), f'{cfg}/{mode} env{i} step {t + 1} done'
            if fused and done[i]:
                continue                      # re-spawned inside the call: the reference episode is over
            got, want = eng.snapshot(i, fields), snap_at(ep, t + 1)
            for key in CMP_KEYS:
                np.testing.assert_array_equal(got[key], want[key], err_msg=f'{cfg}/{mode} env{i} step {t + 1}: {key}')
            if fused:
                np.testing.assert_array_equal(obs[i], ep['obs'][t + 1], err_msg=f'{cfg}/{mode} env{i} t={t + 1} fused obs')
        if not fused:
            check_obs(t + 1)
    eng.close()


@pytest.mark.parametrize('cfg', ALL_CFGS)
def test_replay_untouched_reference(cfg):
    """faithful mode == the unmodified reference (uid-equality artefact included); BOTH observation kernels."""
    _replay_batch(cfg, 'U', obs_kernels=[1, 2])


@pytest.mark.parametrize('cfg', ALL_CFGS)
def test_replay_identity_reference(cfg):
    """identity mode == identity-patched reference; BOTH observation kernels (direct and tiled)."""
    _replay_batch(cfg, 'I', obs_kernels=[1, 2])


@pytest.mark.parametrize('mode', ['U', 'I'])
@pytest.mark.parametrize('cfg', ALL_CFGS)
def test_replay_through_fused_step_observe(cfg, mode):
    """The same reference traces through `mfg_step_observe` (tape + auto_reset): the overlapped side-stream pipeline."""
    _replay_batch(cfg, mode, obs_kernels=[], fused=True)


@pytest.mark.parametrize('mode', ['U', 'I'])
@pytest.mark.parametrize('cfg', FULL_OBS_CFGS)
def test_replay_full_observability(cfg, mode):
    """pomdp_r = 0 (the reference's own `_obs_test.yaml` fixture): whole-level observation planes, ray radius min(H, W);
    served by the direct per-agent kernel."""
    _replay_batch(cfg, mode, obs_kernels=[0])


@pytest.mark.parametrize('cfg', ['cfg1', 'cfg2', 'cfg3', 'cfg4', 'stress2', 'eight_puzzle', 'narrow_corridor', 'stress4'])
@pytest.mark.parametrize('faithful', [False, True])
def test_freerun_matches_host_build_of_device_code(cfg, faithful):
    """Free-running (Philox spawn, dirt respawn, maintainer policy, in-kernel auto reset): the CUDA kernels and the
    g++ build of the same per-env code must agree bit-for-bit on every field, reward, done and observation."""
    es = spec_for(cfg)
    N, steps = 96, 70
    eng = _engine(es, N, faithful=faithful, seed=1234)
    sim = HostSim(es, N, faithful=faithful, seed=1234)
    eng.reset()
    sim.reset()
    acts = torch.zeros((N, es.n_agents), dtype=torch.int32, device='cuda:0')
    for t in range(steps):
        eng.random_actions(acts, seed=99, step_index=t)
        a = acts.cpu().numpy()
        obs, rew, done = eng.step_observe(acts, auto_reset=True)
        r2, d2 = sim.step(a, auto_reset=True)
        f = eng.fields_numpy()
        for name, arr in sim.fields.items():          # both are [rows, N] copies
            if name == 'ep_ret':
                np.testing.assert_allclose(f[name], arr, rtol=1e-12, atol=1e-12)
            else:
                np.testing.assert_array_equal(f[name], arr, err_msg=f'{cfg} t={t} field {name}')
        np.testing.assert_array_equal(rew.cpu().numpy(), r2, err_msg=f'{cfg} t={t} reward')
        np.testing.assert_array_equal(done.cpu().numpy(), d2, err_msg=f'{cfg} t={t} done')
        np.testing.assert_array_equal(obs.cpu().numpy(), sim.observe(), err_msg=f'{cfg} t={t} obs')
    s1, s2 = eng.stats(), sim.stats()
    np.testing.assert_array_equal(s1[:11], s2[:11])
    eng.close()


@pytest.mark.parametrize('faithful', [False, True])
@pytest.mark.parametrize('cfg', ['crowd', 'crowd12'])
def test_freerun_wide_agent_mappings_match_host_build(cfg, faithful):
    """7 and 12 agents per env: the 8- and 16-lane-per-env mappings of the tiled observation kernel (multi-env tiles for
    the odd channel count of `crowd`), AMAX = 8 / 16 step kernels, re-spawns every 60 steps."""
    es = spec_for(cfg)
    N, steps = 70, 130
    eng = _engine(es, N, faithful=faithful, seed=11)
    assert eng.info('tiled_ok') == 1
    sim = HostSim(es, N, faithful=faithful, seed=11)
    eng.reset()
    sim.reset()
    acts = torch.zeros((N, es.n_agents), dtype=torch.int32, device='cuda:0')
    for t in range(steps):
        eng.random_actions(acts, seed=17, step_index=t)
        obs, rew, done = eng.step_observe(acts, auto_reset=True)
        r2, d2 = sim.step(acts.cpu().numpy(), auto_reset=True)
        np.testing.assert_array_equal(done.cpu().numpy(), d2, err_msg=f't={t} done')
        np.testing.assert_array_equal(rew.cpu().numpy(), r2, err_msg=f't={t} reward')
        if t % 5 == 4:
            f = eng.fields_numpy()
            for name, arr in sim.fields.items():
                if name == 'ep_ret':
                    np.testing.assert_allclose(f[name], arr, rtol=1e-12, atol=1e-12)
                else:
                    np.testing.assert_array_equal(f[name], arr, err_msg=f't={t} field {name}')
            np.testing.assert_array_equal(obs.cpu().numpy(), sim.observe(), err_msg=f't={t} obs')
    assert eng.stats()[0] >= 2 * N
    eng.close()


@pytest.mark.parametrize('faithful', [False, True])
@pytest.mark.parametrize('cfg', ['cfg4', 'stress'])
def test_freerun_steady_state_matches_host_build(cfg, faithful):
    """Long free run (several episodes per env: dirt respawns and compaction, un-listed entities, door traffic, re-spawns
    through the side-stream pipeline): CUDA kernels == g++ build of the per-env code, every field / reward / done /
    observation, checked every few steps."""
    es = spec_for(cfg)           # `stress`: max_steps 200 => every env is re-spawned in the same step twice (full-size lists)
    N, steps = 160, 420
    eng = _engine(es, N, faithful=faithful, seed=77)
    sim = HostSim(es, N, faithful=faithful, seed=77)
    eng.reset()
    sim.reset()
    acts = torch.zeros((N, es.n_agents), dtype=torch.int32, device='cuda:0')
    for t in range(steps):
        eng.random_actions(acts, seed=13, step_index=t)
        obs, rew, done = eng.step_observe(acts, auto_reset=True)
        r2, d2 = sim.step(acts.cpu().numpy(), auto_reset=True)
        np.testing.assert_array_equal(done.cpu().numpy(), d2, err_msg=f't={t} done')
        np.testing.assert_array_equal(rew.cpu().numpy(), r2, err_msg=f't={t} reward')
        if t % 7 == 6 or t > steps - 5:
            f = eng.fields_numpy()
            for name, arr in sim.fields.items():
                if name == 'ep_ret':
                    np.testing.assert_allclose(f[name], arr, rtol=1e-12, atol=1e-12)
                else:
                    np.testing.assert_array_equal(f[name], arr, err_msg=f't={t} field {name}')
            np.testing.assert_array_equal(obs.cpu().numpy(), sim.observe(), err_msg=f't={t} obs')
    s1, s2 = eng.stats(), sim.stats()
    np.testing.assert_array_equal(s1[:11], s2[:11])
    assert s1[0] > (2 * N if cfg == 'stress' else 0)
    eng.close()


@pytest.mark.parametrize('faithful', [False, True])
@pytest.mark.parametrize('cfg', ['eight_puzzle', 'narrow_corridor', 'obs_test', 'cfg2', 'cfg4'])
def test_exact_kernel_variants_agree_at_scale(cfg, faithful):
    """The exact per-agent path in its two forms - block-staged state with the first-visit rank table where there is one
    (obs_kernel 1: the product path of full observability) and plain global state with the ray walk (3) - at a ragged batch
    size, free-running."""
    es = spec_for(cfg)
    N = 1024 + 37
    eng = _engine(es, N, faithful=faithful, seed=11)
    eng.reset()
    acts = torch.zeros((N, es.n_agents), dtype=torch.int32, device='cuda:0')
    for t in range(24):
        eng.random_actions(acts, seed=5, step_index=t)
        eng.step(acts, auto_reset=True)
        if t % 6 == 5:
            eng.set_option('obs_kernel', 1)
            o1 = eng.observe().clone()
            eng.set_option('obs_kernel', 3)
            assert torch.equal(o1, eng.observe()), f'{cfg} t={t}: block-staged exact kernel != plain exact kernel'
    eng.close()


@pytest.mark.parametrize('faithful', [False, True])
@pytest.mark.parametrize('cfg', ['cfg1', 'cfg2', 'cfg4', 'stress'])
def test_tiled_observation_kernel_equals_direct_kernel_at_scale(cfg, faithful):
    """Size-independent property at a larger batch: both observation kernels produce identical tensors, agents never
    stand on walls, door timers stay within the auto-close interval, dirt amounts within (0, 5]."""
    es = spec_for(cfg)
    N = 8192 + 5                     # ragged: not a multiple of the 32-env CTA tile
    eng = _engine(es, N, faithful=faithful, seed=7)
    eng.reset()
    acts = torch.zeros((N, es.n_agents), dtype=torch.int32, device='cuda:0')
    walls = torch.as_tensor(es.walls, device='cuda:0')
    for t in range(40):
        eng.random_actions(acts, seed=3, step_index=t)
        eng.step(acts, auto_reset=True)
        if t % 8 == 7:
            eng.set_option('obs_kernel', 1)
            o1 = eng.observe().clone()
            eng.set_option('obs_kernel', 2)
            o2 = eng.observe().clone()
            assert torch.equal(o1, o2), f'{cfg} t={t}: tiled != direct'
            eng.set_option('obs_store', 0)               # LDS/STG write-out instead of the TMA bulk store
            assert torch.equal(o1, eng.observe()), f'{cfg} t={t}: tiled (no bulk store) != direct'
            eng.set_option('obs_store', 1)
            eng.set_option('obs_cap', 3)                 # tiny sprite lists: most envs take the overflow path
            assert torch.equal(o1, eng.observe()), f'{cfg} t={t}: tiled (overflow path) != direct'
            eng.set_option('obs_cap', eng.info('obs_cap_max'))
            apos = eng.field('apos').to(torch.int64) & 0xFFFF
            assert not walls[apos >> 8, apos & 255].any()
            if es.n_doors:
                assert int(eng.field('door_timer').max()) <= 10
            if es.has_dirt:
                live = (eng.field('dirt_pos').to(torch.int64) & 0xFFFF) != 0xFFFF
                amt = eng.field('dirt_amt')[live]
                assert float(amt.min()) > 0 and float(amt.max()) <= 5.0
    eng.close()


@pytest.mark.parametrize('faithful', [False, True])
def test_full_size_properties_all_modules_262144_envs(faithful):
    """BASELINE configs[3] at its full size (all modules, 262 144 envs), through size-independent properties: the fused
    step_observe pipeline (side-stream re-spawn, list-mode observation) is deterministic, its observations equal the
    direct per-agent kernel on the whole batch, finished envs restart at step 0, agents never stand on walls."""
    es = spec_for('cfg4')
    N = 262144
    a, b = _engine(es, N, faithful=faithful, seed=21), _engine(es, N, faithful=faithful, seed=21)
    a.reset()
    b.reset()
    acts = torch.zeros((N, es.n_agents), dtype=torch.int32, device='cuda:0')
    walls = torch.as_tensor(es.walls, device='cuda:0')
    for t in range(24):
        a.random_actions(acts, seed=5, step_index=t)
        o1, r1, d1 = a.step_observe(acts, auto_reset=True)
        o2, r2, d2 = b.step_observe(acts, auto_reset=True)
        if t % 8 == 7:
            assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2), t
            assert int(d1.sum()) > 0
            assert int(a.field('step')[0][d1.bool()].max()) == 0          # finished envs were re-spawned in the same call
            o1 = o1.clone()
            a.set_option('obs_kernel', 1)
            assert torch.equal(o1, a.observe()), f't={t}: fused pipeline != direct kernel'
            a.set_option('obs_kernel', 0)
            apos = a.field('apos').to(torch.int64) & 0xFFFF
            assert not walls[apos >> 8, apos & 255].any()
    np.testing.assert_array_equal(a.stats()[:11], b.stats()[:11])
    a.close()
    b.close()


def test_inline_and_deferred_auto_reset_agree():
    """Three forms of auto-reset give identical results: the packed reset kernel overlapped with the observation kernel
    on a side stream (default of mfg_step_observe), the same kernel serialised, and the in-line reset inside k_step."""
    es = spec_for('cfg4')
    engs = [_engine(es, 640, faithful=True, seed=9) for _ in range(3)]
    a, b, c = engs
    b.set_option('defer_reset', 0)
    c.set_option('overlap_reset', 0)
    for e in engs:
        e.reset()
    acts = torch.zeros((640, es.n_agents), dtype=torch.int32, device='cuda:0')
    for t in range(60):
        a.random_actions(acts, seed=2, step_index=t)
        o1, r1, d1 = a.step_observe(acts, auto_reset=True)
        for other in (b, c):
            o2, r2, d2 = other.step_observe(acts, auto_reset=True)
            assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2), t
    np.testing.assert_array_equal(a.stats()[:11], b.stats()[:11])
    np.testing.assert_array_equal(a.stats()[:11], c.stats()[:11])
    assert a.stats()[0] > 0
    for e in engs:
        e.close()


@pytest.mark.parametrize('faithful', [True, False])
def test_step_kernel_variants_agree(faithful):
    """k_step launch shapes (barriers at the convergent points, dirt-uid rows left in HBM, 1-3 state blocks per CTA) are
    schedules of the same per-env program: every field, reward, done flag and observation stays bit-identical, also on a
    ragged batch (a partly filled last block and a CTA with fewer blocks than its siblings)."""
    es = spec_for('cfg4')
    N = 128 * 4 + 37
    variants = [(0, 1), (2, 1), (1, 1), (1, 2), (1, 3), (1, 0)]      # (step_kernel, step_blocks)
    engs = []
    for sk, nb in variants:
        e = _engine(es, N, faithful=faithful, seed=21)
        e.set_option('step_kernel', sk)
        e.set_option('step_blocks', nb)
        e.reset()
        engs.append(e)
    acts = torch.zeros((N, es.n_agents), dtype=torch.int32, device='cuda:0')
    for t in range(80):
        engs[0].random_actions(acts, seed=4, step_index=t)
        o1, r1, d1 = engs[0].step_observe(acts, auto_reset=True)
        for v, other in zip(variants[1:], engs[1:]):
            o2, r2, d2 = other.step_observe(acts, auto_reset=True)
            assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2), (t, v)
    f0 = engs[0].fields_numpy()
    for v, other in zip(variants[1:], engs[1:]):
        for name, arr in other.fields_numpy().items():
            np.testing.assert_array_equal(f0[name], arr, err_msg=f'{v} field {name}')
        np.testing.assert_array_equal(engs[0].stats()[:11], other.stats()[:11])
    assert engs[0].stats()[0] > 0
    for e in engs:
        e.close()


def test_env_shards_are_independent_of_the_partition():
    """Multi-GPU property on one device: envs [0, 256) as one engine == two engines of 128 with env_id_offset 0 / 128."""
    es = spec_for('cfg4')
    whole = _engine(es, 256, faithful=False, seed=5)
    parts = [_engine(es, 128, faithful=False, seed=5, env_id_offset=o) for o in (0, 128)]
    for e in [whole] + parts:
        e.reset()
    acts = torch.zeros((256, es.n_agents), dtype=torch.int32, device='cuda:0')
    for t in range(30):
        whole.random_actions(acts, seed=11, step_index=t)
        o, r, d = whole.step_observe(acts, auto_reset=True)
        for k, p in enumerate(parts):
            sl = slice(128 * k, 128 * (k + 1))
            o2, r2, d2 = p.step_observe(acts[sl].contiguous(), auto_reset=True)
            assert torch.equal(o[sl], o2) and torch.equal(r[sl], r2) and torch.equal(d[sl], d2)
    tot = whole.stats()
    np.testing.assert_array_equal(tot[:11], sum(p.stats()[:11] for p in parts))
    for e in [whole] + parts:
        e.close()


def test_factory_surface_unbatched_matches_reference_shapes():
    from marl_factory_grid_b200 import Factory
    from golden_util import CONFIGS
    f = Factory(CONFIGS / 'cfg4.yaml', device='cuda:0')
    ep = episodes('cfg4')[0]
    assert list(f.named_action_space) == ep['meta']['agent_names']
    assert f.named_action_space == ep['meta']['named_action_space']
    obs = f.reset()
    assert list(obs) == ep['meta']['agent_names']
    assert [o.shape for o in obs.values()] == [(7, 7, 7), (8, 7, 7), (7, 7, 7), (14, 7, 7)]
    _, o, r, d, info = f.step([0, 1, 2, 3])
    assert len(o) == 4 and len(r) == 4 and isinstance(d, bool) and info['step'] == 1
    f.close()


def test_factory_batched_host_path_and_stats():
    from marl_factory_grid_b200 import Factory
    from golden_util import CONFIGS
    N = 1024
    f = Factory(CONFIGS / 'cfg3.yaml', n_envs=N, device='cuda:0', parity='identity', auto_reset=True)
    obs = f.reset()
    assert obs['Agent[Wolfgang]'].shape == (N, 5, 7, 7)
    A = f.spec.n_agents
    h_act = torch.zeros((N, A), dtype=torch.int32).pin_memory()
    h_rew = torch.zeros((N, A), dtype=torch.float32).pin_memory()
    h_done = torch.zeros(N, dtype=torch.uint8).pin_memory()
    h_obs = torch.zeros((N, f.spec.total_channels, 7, 7), dtype=torch.float32).pin_memory()
    g = torch.Generator().manual_seed(0)
    for t in range(120):
        h_act.copy_(torch.randint(0, 12, (N, A), generator=g, dtype=torch.int32))
        f.engine.step_host(h_act, h_rew, h_done, h_obs, auto_reset=True)
    assert torch.equal(h_obs, f.engine.observe().cpu())      # the host copy is the observation of the current state
    assert h_rew.abs().sum() > 0
    st = f.episode_stats()
    assert st['episodes'] > 0 and st['steps'] >= st['episodes']
    f.close()


def test_batched_env_monitor_records_every_finished_episode(tmp_path):
    """EnvMonitor (utils/logging/envmonitor.py:15-73) over the batched Factory: one row per finished episode, episode
    lengths and returns equal to what the step stream says, totals equal to the engine's device-side statistics."""
    import pickle
    from marl_factory_grid_b200 import EnvMonitor, Factory
    from golden_util import CONFIGS
    N = 512
    f = Factory(CONFIGS / 'stress.yaml', n_envs=N, device='cuda:0', parity='faithful', auto_reset=True, seed=3)
    m = EnvMonitor(f, filepath=tmp_path / 'monitor.pick')
    m.reset()
    A = f.spec.n_agents
    g = torch.Generator(device='cuda:0').manual_seed(0)
    hi = torch.tensor(f.spec.n_actions, device='cuda:0')
    ret = np.zeros((N, A)); length = np.zeros(N, np.int64); want = []
    for t in range(230):                       # max_steps = 200: every env finishes at least once
        acts = (torch.rand((N, A), generator=g, device='cuda:0') * hi).to(torch.int32)
        _, _, r, d, _ = m.step(acts)
        ret += r.cpu().numpy().astype(np.float64); length += 1
        for e in np.nonzero(d.cpu().numpy())[0]:
            want.append((int(e), int(length[e]), ret[e].copy()))
            ret[e] = 0; length[e] = 0
    df = m.monitor_df
    assert len(df) == len(want) >= N and list(df['episode']) == list(range(len(df)))
    np.testing.assert_array_equal(df['env'].to_numpy(), [w[0] for w in want])
    np.testing.assert_array_equal(df['steps'].to_numpy(), [w[1] for w in want])
    np.testing.assert_allclose(df[list(df.columns[4:])].to_numpy(), np.stack([w[2] for w in want]), rtol=0, atol=1e-12)
    st = f.episode_stats()
    assert st['episodes'] == len(df) and st['steps'] == int(df['steps'].sum())
    np.testing.assert_allclose(st['return_sum'], df['step_reward'].sum(), rtol=1e-5)
    m.save_monitor()
    back = pickle.load(open(tmp_path / 'monitor.pick', 'rb'))
    assert len(back) == len(df) and 'index' in back.columns
    f.close()


def test_cuda_graph_replay_equals_eager_step_observe():
    """Engine.capture_step: the fused call captured in a CUDA graph (side-stream fork / join included) gives the same
    observations, rewards, done flags and state as eager calls."""
    es = spec_for('cfg4')
    N = 1024
    a, b = _engine(es, N, faithful=True, seed=4), _engine(es, N, faithful=True, seed=4)
    a.reset()
    b.reset()
    acts = torch.zeros((N, es.n_agents), dtype=torch.int32, device='cuda:0')
    graph = b.capture_step(acts, auto_reset=True)
    for t in range(40):
        a.random_actions(acts, seed=6, step_index=t)
        o1, r1, d1 = a.step_observe(acts, auto_reset=True)
        graph.replay()
        assert torch.equal(o1, b.obs) and torch.equal(r1, b.reward) and torch.equal(d1, b.done), t
    assert torch.equal(a.state, b.state)
    a.close()
    b.close()


def test_consecutive_resets_draw_fresh_layouts_and_match_host_build():
    """every `reset()` advances the per-env episode counter (Philox counter word), so `env.reset()` per episode
    (auto_reset off, the reference's usual loop) sees a new layout each time; reset(mask = all) == reset()."""
    es = spec_for('cfg4')
    N = 300
    eng, eng2 = _engine(es, N, faithful=True, seed=5), _engine(es, N, faithful=True, seed=5)
    sim = HostSim(es, N, faithful=True, seed=5)
    layouts = []
    for k in range(3):
        eng.reset()
        sim.reset()
        eng2.reset(None if k == 0 else torch.ones(N, dtype=torch.uint8))
        f = eng.fields_numpy()
        for name, arr in sim.fields.items():
            np.testing.assert_array_equal(f[name], arr, err_msg=f'reset {k} field {name}')
        for name, arr in eng2.fields_numpy().items():
            np.testing.assert_array_equal(f[name], arr, err_msg=f'reset {k} masked == unmasked: {name}')
        assert np.all(f['episode'] == k)
        layouts.append(f['apos'].copy())
    assert np.mean(np.any(layouts[0] != layouts[1], axis=0)) > 0.9 and np.mean(np.any(layouts[1] != layouts[2], axis=0)) > 0.9
    eng.close()
    eng2.close()


def test_factory_scalar_reward_config_raises_like_the_reference(tmp_path):
    """`individual_rewards: false`: the reference constructs and resets, then raises TypeError at
    environment/factory.py:217 on the first step."""
    import yaml
    from marl_factory_grid_b200 import Factory
    from golden_util import CONFIGS
    cfg = yaml.safe_load((CONFIGS / 'cfg2.yaml').read_text())
    cfg['General']['individual_rewards'] = False
    p = tmp_path / 'scalar.yaml'
    p.write_text(yaml.safe_dump(cfg, sort_keys=False))
    f = Factory(p, device='cuda:0')
    obs = f.reset()
    assert len(obs) == 2
    with pytest.raises(TypeError):
        f.step([0, 0])
    f.close()
