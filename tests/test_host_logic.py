"""CPU tests of the host-side logic: config compiler error behaviour, ray tables, state <-> snapshot round trips,
the C-ABI library's exported symbols, and property tests of the device code's free-running mode (host build)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import yaml

from golden_util import CONFIGS, ROOT, episodes, snap_at, spec_for
from marl_factory_grid_b200 import FactoryConfigParser
from marl_factory_grid_b200.rays import full_ray_table, window_ray_table
from marl_factory_grid_b200.state_io import columns_to_snapshot, snapshot_to_columns


def test_abi_library_exports_every_declared_symbol():
    """The built .so must export exactly the entry points include/mfg_b200.h declares (no compute calls here)."""
    header = (ROOT / 'include' / 'mfg_b200.h').read_text()
    declared = set(re.findall(r'\b(mfg_[a-z_]+)\s*\(', header))
    assert {'mfg_create', 'mfg_reset', 'mfg_step', 'mfg_observe', 'mfg_step_observe', 'mfg_stats'} <= declared
    lib_path = ROOT / 'marl_factory_grid_b200' / 'libmfg_b200.so'
    if not lib_path.exists():
        pytest.skip('libmfg_b200.so not built (run __graft_entry__.build())')
    lib = ctypes.CDLL(str(lib_path))
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in mfg_b200.h but not exported'
    lib.mfg_version.restype = ctypes.c_char_p
    assert b'sm_100a' in lib.mfg_version()


def test_ctypes_spec_mirror_has_the_c_struct_size():
    """sizeof(MfgSpec) seen by ctypes == sizeof seen by the C++ compiler (checked through the host build)."""
    from hostsim_util import HostSim
    sim = HostSim(spec_for('cfg4'), 2)     # hs_create would read garbage (and fail validation) on a layout mismatch
    assert sim.fields['apos'].shape == (4, 2) and sim._views['apos'].shape == (1, 4, 128)


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from marl_factory_grid_b200.engine import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(spec_for('cfg1'), 4)


@pytest.mark.parametrize('cfg', ['cfg1', 'cfg2', 'cfg3', 'cfg4', 'stress', 'stress2', 'default_config', 'clean_and_bring',
                                 'stress3', 'dest_all', 'dest_simul', 'obs_test', 'eight_puzzle', 'narrow_corridor', 'stress4'])
def test_named_spaces_equal_the_reference(cfg):
    """Agent names, named action space, action counts and the per-agent observation layer names recorded from the
    reference's own `Factory` (meta of the golden traces) == what the config compiler derives (no GPU needed)."""
    from marl_factory_grid_b200.config_parser import named_action_space
    meta, es = episodes(cfg)[0]['meta'], spec_for(cfg)
    assert [a.name for a in es.agents] == meta['agent_names']
    assert named_action_space(es) == meta['named_action_space']
    assert list(es.n_actions) == meta['n_actions']
    assert [[act.name for act in a.actions] for a in es.agents] == meta['action_names']
    assert [[ch.name for ch in a.channels] for a in es.agents] == meta['obs_layers']      # == Factory.named_observation_space
    assert [es.H, es.W] == meta['level_shape']


def test_ray_table_matches_survey_appendix_c():
    full = full_ray_table(7)
    assert len(full) == 44 and sum(len(r) for r in full) == 320
    assert [r[-1] for r in full][:5] == [(-7, -2), (-7, -1), (-7, 0), (-7, 1), (-7, 2)]
    win = window_ray_table(3)
    assert len(win) == 28 and all(len(r) == 3 for r in win)
    assert len({c for r in win for c in r}) == 48        # every window cell except the origin is reachable
    assert ((-1, -1), (-2, -2), (-3, -3)) in win and ((1, 1), (2, 2), (2, 3)) in win


def test_known_answer_visible_walls_obs_test_map():
    """SURVEY.md §4 known answer, extracted from the unmodified reference: `_obs_test.yaml`, random.seed(0), first
    reset(): visible wall counts per agent Wolfgang 18, Soeren 22, Juergen 18, Walter 19; Doors channel sums to 0;
    observation shape (6, 12, 12).  The committed fixture must show it and the oracle must reproduce it."""
    from oracle import OracleEnv, load_snapshot
    ep = episodes('obs_test')[0]
    obs0 = ep['obs'][0].reshape(4, 6, 12, 12)
    assert [int(obs0[a, 0].sum()) for a in range(4)] == [18, 22, 18, 19]
    assert float(np.abs(obs0[:, 1]).sum()) == 0.0
    assert all(float(obs0[a, c].sum()) == 1.0 for a in range(4) for c in (2, 3, 4))
    es = spec_for('obs_test')
    assert es.obs_shape == (12, 12)
    env = load_snapshot(OracleEnv(es, faithful=True), snap_at(ep, 0), ep['door_pos'])
    got = env.observe().reshape(4, 6, 12, 12)
    assert [int(got[a, 0].sum()) for a in range(4)] == [18, 22, 18, 19]


@pytest.mark.parametrize('cfg', ['cfg1', 'cfg4', 'stress2'])
def test_snapshot_roundtrip(cfg):
    es = spec_for(cfg)
    ep = episodes(cfg)[0]
    for t in (0, len(ep['actions']) // 2):
        snap = snap_at(ep, t)
        back = columns_to_snapshot(es, snapshot_to_columns(es, snap))
        for key in ('agent_pos', 'door_open', 'door_timer', 'dirt_n', 'dirt_pos', 'dirt_amt', 'dirt_uid', 'item_pos',
                    'maint_pos', 'dest_reached', 'battery', 'door_listed', 'dirt_listed', 'item_listed'):
            np.testing.assert_array_equal(back[key], snap[key], err_msg=key)


def test_snapshot_with_a_foreign_uid_model_is_rejected():
    """The step kernel's uid listing relies on the fresh-episode uid model (pile uids strictly increasing in creation order);
    a hand-made state that breaks it must not load."""
    es = spec_for('cfg4')
    snap = dict(snap_at(episodes('cfg4')[0], 0))
    assert int(snap['dirt_n']) >= 2
    uid = np.array(snap['dirt_uid']).copy()
    uid[1] = uid[0]
    snap['dirt_uid'] = uid
    with pytest.raises(ValueError, match='dirt uids'):
        snapshot_to_columns(es, snap)


def _write(tmp_path, mutate):
    cfg = yaml.safe_load((CONFIGS / 'cfg4.yaml').read_text())
    mutate(cfg)
    p = tmp_path / 'c.yaml'
    p.write_text(yaml.safe_dump(cfg, sort_keys=False))
    return p


def test_config_errors_raise_instead_of_exit(tmp_path):
    with pytest.raises(NotImplementedError):
        FactoryConfigParser(CONFIGS / 'cfg1.yaml', custom_modules_path='/tmp/x')
    with pytest.raises(NotImplementedError):
        FactoryConfigParser(_write(tmp_path, lambda c: c['Rules'].update(MyRule={}))).compile()
    with pytest.raises(NotImplementedError):
        FactoryConfigParser(_write(tmp_path, lambda c: c['Entities'].update(Unicorns={}))).compile()
    with pytest.raises(ValueError):      # observation of a group that is not configured
        FactoryConfigParser(_write(tmp_path, lambda c: c['Entities'].pop('Machines'))).compile()
    with pytest.raises(ValueError):      # action without its entities
        FactoryConfigParser(_write(tmp_path, lambda c: (c['Entities'].pop('DirtPiles'), c['Rules'].pop('RespawnDirt'),
                                                         c['Rules'].pop('EntitiesSmearDirtOnMove'),
                                                         c['Rules'].pop('DoneOnAllDirtCleaned')))).compile()
    with pytest.raises(TypeError):
        FactoryConfigParser(_write(tmp_path, lambda c: c['Rules']['DoneAtMaxStepsReached'].update(bogus=1))).compile()
    with pytest.raises(FileNotFoundError):
        FactoryConfigParser(_write(tmp_path, lambda c: c['General'].update(level_name='nope'))).compile()


def test_ragged_level_is_rejected(tmp_path):
    lvl = tmp_path / 'bad.txt'
    lvl.write_text('#####\n#--#\n#####\n')
    with pytest.raises(AssertionError):
        FactoryConfigParser(CONFIGS / 'cfg1.yaml').compile(custom_level_path=lvl)


def test_algorithmic_bytes_match_survey_8d():
    assert [spec_for(c).algorithmic_bytes_per_env_step() for c in ('cfg1', 'cfg2', 'cfg3', 'cfg4')] == [1225, 2427, 4097, 8105]


@pytest.mark.parametrize('cfg', ['cfg1', 'cfg2', 'cfg3', 'cfg4', 'stress', 'stress2', 'narrow_corridor', 'stress4'])
def test_freerun_invariants_host_build(cfg):
    """Philox spawn + free-running rules on the host build of the device code: structural invariants of the
    reference (SURVEY.md §4 iii) hold on every step, for a batch that auto-resets."""
    from hostsim_util import HostSim
    es = spec_for(cfg)
    N = 48
    sim = HostSim(es, N, faithful=True, seed=42)
    sim.reset()
    rng = np.random.default_rng(0)
    floor = set(map(tuple, es.floor.tolist()))
    doors = set(map(tuple, es.door_pos.tolist()))
    snap0 = [sim.snapshot(e) for e in range(N)]
    for s in snap0:
        pos = [tuple(p) for p in s['agent_pos']]
        assert all(p in floor and p not in doors for p in pos) and len(set(pos)) == len(pos)   # agents on EMPTY tiles
        for key in ('item', 'pod', 'dest', 'drop', 'machine', 'maint'):
            gp = [tuple(p) for p in s[f'{key}_pos']]
            assert len(set(gp)) == len(gp) and all(p in floor and p not in doors for p in gp)   # free tiles, distinct
        if es.has_dirt:
            assert s['dirt_n'] in (es.dirt_quantity - 1, es.dirt_quantity)
            amt = s['dirt_amt'][:s['dirt_n']]
            assert np.all(np.abs(amt - es.dirt_initial_amount) <= es.dirt_amount_var + 1e-12)
    if not all(a.positions for a in es.agents):              # (fixed start positions: every env starts the same)
        assert len({tuple(map(tuple, s['agent_pos'])) for s in snap0}) > N // 2                   # envs differ
    n_act = es.n_actions
    for t in range(120):
        a = np.stack([rng.integers(0, n, N) for n in n_act], 1).astype(np.int32)
        rew, done = sim.step(a, auto_reset=True)
        assert np.all(np.isfinite(rew))
        for e in range(0, N, 7):
            s = sim.snapshot(e)
            assert all(tuple(p) in floor for p in s['agent_pos'])
            assert all(tuple(p) in floor for p in s['maint_pos'])
            assert np.all(s['door_timer'] <= 10) and np.all(s['door_timer'] >= 0)
            if es.has_dirt:
                amt = s['dirt_amt'][:s['dirt_n']]
                assert np.all(amt > 0) and np.all(amt <= 5.0)
                dp = [tuple(p) for p in s['dirt_pos'][:s['dirt_n']]]
                assert len(set(dp)) == len(dp)
    st = sim.stats()
    assert st[9] == 0 and st[10] == 0                      # no dirt overflow; no spawn failure
    if cfg in ('stress', 'cfg3', 'cfg4'):
        assert st[0] > 0 and st[1] >= st[0]                # some episodes finished and were re-spawned in place


def test_every_reset_draws_a_fresh_layout_host_build():
    """a second full `reset()` must not replay episode 0.  Two consecutive unmasked resets give different
    spawn layouts, and reset(mask = all ones) == the second unmasked reset (both advance each env's episode counter)."""
    from hostsim_util import HostSim
    es = spec_for('cfg4')
    N = 40
    a, b = HostSim(es, N, faithful=True, seed=5), HostSim(es, N, faithful=True, seed=5)
    a.reset(); b.reset()
    first = a.fields
    np.testing.assert_array_equal(first['apos'], b.fields['apos'])
    assert np.all(first['episode'] == 0)
    a.reset()
    b.reset(np.ones(N, np.uint8))
    second = a.fields
    assert np.all(second['episode'] == 1)
    assert np.mean(np.any(second['apos'] != first['apos'], axis=0)) > 0.9          # fresh layouts
    for name, arr in b.fields.items():
        np.testing.assert_array_equal(arr, second[name], err_msg=name)


def test_finished_envs_are_counted_once_without_auto_reset_host_build():
    """with auto_reset off (the reference's convention) a finished env may keep being stepped; its episode
    enters the statistics once, not once per step."""
    from hostsim_util import HostSim
    es = spec_for('stress')                  # max_steps = 200, several early done rules
    N = 16
    sim = HostSim(es, N, faithful=True, seed=1)
    sim.reset()
    rng = np.random.default_rng(3)
    ever = np.zeros(N, bool)
    for t in range(230):
        a = np.stack([rng.integers(0, n, N) for n in es.n_actions], 1).astype(np.int32)
        _, done = sim.step(a, auto_reset=False)
        ever |= done.astype(bool)
    assert ever.all()
    st = sim.stats()
    assert st[0] == N and st[1] <= 200 * N
    assert np.all(sim.fields['finished'][0] == 1)
    sim.reset()
    assert np.all(sim.fields['finished'][0] == 0)


def test_limits_are_rejected_not_wrapped(tmp_path):
    from hostsim_util import HostSim
    with pytest.raises(ValueError):          # 16-bit step counter
        FactoryConfigParser(_write(tmp_path, lambda c: c['Rules']['DoneAtMaxStepsReached'].update(max_steps=70000))).compile()
    # individual_rewards: false is dead in the reference (TypeError at environment/factory.py:217 on the first step)
    es = FactoryConfigParser(_write(tmp_path, lambda c: c['General'].update(individual_rewards=False))).compile()
    assert es.individual_rewards is False
    with pytest.raises(RuntimeError, match='individual_rewards'):
        HostSim(es, 1)                       # the C-side spec validation (shared with mfg_create)
    from oracle.freerun import FreeRunEnv
    env = FreeRunEnv(es, faithful=True, seed=0)
    env.reset()
    with pytest.raises(TypeError):
        env.step_free([0] * es.n_agents)


def test_bound_destination_spawn_rules_host_build():
    """SpawnDestinationOnAgent + DoRandomInitialSteps (eight_puzzle) and SpawnDestinationsPerAgent (narrow_corridor) in the
    Philox spawn of the device code: destinations are bound one per agent; eight_puzzle keeps its single free tile and at most
    `random_steps` agents have left their destination; narrow_corridor puts every destination on the candidate tile that is not
    the agent's own (modules/destinations/rules.py:95-162, environment/rules.py:328-355)."""
    from hostsim_util import HostSim
    es = spec_for('eight_puzzle')
    assert es.dest_mode == 1 and es.dest_bound == list(range(8)) and es.random_initial_steps == 2
    sim = HostSim(es, 64, faithful=True, seed=9)
    sim.reset()
    moved_hist = np.zeros(4, int)
    for e in range(64):
        s = sim.snapshot(e)
        apos, dpos = [tuple(p) for p in s['agent_pos']], [tuple(p) for p in s['dest_pos']]
        assert len(set(apos)) == 8 and len(set(dpos)) == 8
        moved = sum(a != d for a, d in zip(apos, dpos))
        assert moved <= 2
        moved_hist[moved] += 1
    assert moved_hist[1] + moved_hist[2] > 0                  # the random initial steps do move agents
    es = spec_for('narrow_corridor')
    assert es.dest_mode == 2 and es.dest_bound == [0, 1]
    sim = HostSim(es, 32, faithful=True, seed=9)
    sim.reset()
    for e in range(32):
        s = sim.snapshot(e)
        apos, dpos = [tuple(p) for p in s['agent_pos']], [tuple(p) for p in s['dest_pos']]
        assert apos == [(2, 1), (2, 5)] and dpos == [(2, 5), (2, 1)]
