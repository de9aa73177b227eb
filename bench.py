#!/usr/bin/env python
"""bench.py - agent-steps/s including observations for the batched marl-factory-grid stepping engine.

One "step" = one pass of the hot path over the whole batch: draw uniform random actions on the device (Philox), then ONE
`mfg_step_observe` call (k_step: actions, rule hooks, done / reward; packed re-spawn of the finished envs on a side stream;
ray-cast observation tensor).  Workload at every N: BASELINE.json configs[3]/[4], the all-modules config (cfg4: level
`large`, 4 heterogeneous agents, POMDP r=3, every module incl. machines + maintainer) with 1,048,576 envs PER GPU
(the 1M..8M sweep of configs[4] => weak scaling; envs shard by global env id, no per-step collective, only the
episode-statistics vector is all-reduced over NCCL).  Headline = `faithful` parity mode (the untouched reference); the
identity-patched mode is measured in a second pass and reported under `other_parity_mode`.

    python bench.py --gpus 1 --steps 1000 --warmup 100           # this engine (defaults: SURVEY.md 8d protocol)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # CPU reference arm (oracle port, all host cores)

Prints ONE JSON line (rank 0).  Extra keys: `roofline` (dominant kernel vs measured HBM peak), `cpu_baseline`
(oracle port on the host cores, bounded sample), `e2e` (host buffers through mfg_step_host, copies timed),
`clocks`, `gpu_launches`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'agent-steps/sec incl. obs'
UNIT = 'agent-steps/s'
CONFIGS = ROOT / 'marl_factory_grid_b200' / 'configs'
WORKLOADS = {          # BASELINE.json configs[0..3] as restated in SURVEY.md App. D; cfg4 at 1M envs per GPU = configs[4]
    'cfg1': 'single-agent dirt clean-up (level simple, POMDP r=3)',
    'cfg2': '2-agent item pick-up / drop-off with doors (level rooms, POMDP r=3)',
    'cfg3': '4-agent batteries + charge pods + destinations (level large, POMDP r=3)',
    'cfg4': 'all-modules incl. machines + maintainer (level large, 4 agents, POMDP r=3)',
}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's CPU path, multiprocess over the host cores
# ---------------------------------------------------------------------------------------------------------------
class CpuPool:
    """Persistent multiprocess pool of CPU environments, one per host core.

    kind 'reference' = the UNMODIFIED reference installed into baseline/_ref (baseline/install_ref.sh, SURVEY.md 8d protocol:
    own Factory per worker, random.seed(worker), 30 warm-up steps, in-place reset on done, stdout suppressed);
    kind 'port' = the oracle restatement (oracle/freerun.py), used when the reference install is absent."""

    def __init__(self, cfg: str, faithful: bool, procs: int, kind: str = 'auto'):
        import multiprocessing as mp
        from baseline import ref_worker
        if kind == 'auto':
            kind = 'reference' if ref_worker.available() else 'port'
        self.kind, self.procs = kind, procs
        cfg_path = str(CONFIGS / f'{cfg}.yaml')
        if kind == 'reference':
            self._run = ref_worker.worker_run
            self.pool = mp.get_context('spawn').Pool(procs, initializer=ref_worker.worker_init, initargs=(cfg_path, 30))
        else:
            from oracle.freerun import worker_init, worker_run
            self._run = worker_run
            self.pool = mp.get_context('spawn').Pool(procs, initializer=worker_init, initargs=(cfg_path, faithful, 1000))

    def run(self, steps_per_worker: int):
        """Every worker advances its env by steps_per_worker; returns (agent_steps, slowest worker seconds)."""
        res = self.pool.map(self._run, [steps_per_worker] * self.procs, chunksize=1)
        return sum(r[0] for r in res), max(r[1] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_sample(cfg: str, faithful: bool, procs: int, kind: str, seconds: float):
    """Bounded sample: a short calibration run sizes the timed run to ~`seconds`.  Returns (value, steps per worker, slowest s)."""
    pool = CpuPool(cfg, faithful, procs, kind)
    a, t = pool.run(20)
    per = max(20, int(20 * seconds / max(t, 1e-3)))
    agent_steps, slowest = pool.run(per)
    pool.close()
    return agent_steps / slowest, per, slowest, pool.kind


def workload_config(args, world: int):
    """The `config` object of the JSON line: a function of the arguments only, so that the engine arm and the reference arm
    print the same one."""
    from marl_factory_grid_b200 import FactoryConfigParser
    es = FactoryConfigParser(CONFIGS / f'{args.config}.yaml').compile()
    n_local = args.envs_per_gpu
    step_bytes = es.algorithmic_bytes_per_env_step()
    return {'workload': f'{args.config} {WORKLOADS.get(args.config, "")}, {n_local} envs per GPU',
            'envs_total': world * n_local, 'agents': es.n_agents, 'parity': args.parity,
            'parity_note': 'faithful = the untouched reference incl. its uid-equality artefact (bit-exact vs oracle-U traces); '
                           'identity = the identity-patched reference (oracle-I), reported under other_parity_mode',
            'l2': f'per-step working set {step_bytes * n_local / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)',
            'actions': 'uniform random, regenerated every step (engine: device Philox inside the timed region)', 'auto_reset': True,
            'aged_steps': aged_steps(args),
            'aged_note': 'untimed env-steps run before the warm-up so that a short timed region sees the steady-state episode mix'}


def aged_steps(args) -> int:
    """Short runs (the driver's --steps 20) would otherwise time young episodes only: few dirt piles, no un-listed entities.
    Below 200 timed steps, ageing + warm-up together cover at least 300 env-steps (the SURVEY 8d protocol - 100 + 1000
    steps - spans the whole episode-age mix by itself)."""
    if args.age >= 0:
        return args.age
    return max(0, 300 - args.warmup) if args.steps < 200 else 0


def run_reference(args):
    """--impl reference: times the reference's own CPU implementation of the path on the host cores (the unmodified
    reference from baseline/_ref; the oracle port only if that install is absent).  Each bench "step" is a bounded sample:
    every worker advances its env by `per_step` env-steps."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    pool = CpuPool(args.config, args.parity == 'faithful', procs)
    a, t = pool.run(10)                                           # calibration: size a bench step to ~cpu_seconds / steps
    per_step = max(5, int(10 * args.cpu_seconds / max(args.steps, 1) / max(t, 1e-3)))
    for _ in range(max(args.warmup, 1)):
        pool.run(per_step)
    tot_steps, tot_time = 0, 0.0
    for _ in range(args.steps):
        agent_steps, slowest = pool.run(per_step)
        tot_steps += agent_steps
        tot_time += slowest
    pool.close()
    value = tot_steps / tot_time
    what = 'unmodified reference (baseline/_ref)' if pool.kind == 'reference' else 'oracle port of the reference step+obs'
    sample = (f'{procs} procs x {per_step} env-steps of {args.config} per bench step, {what} (own Factory per worker, random '
              f'actions, obs built every step, in-place reset on done)')
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * tot_time / max(args.steps, 1), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, int(os.environ.get('WORLD_SIZE', '1'))), 'cpu_arm': what,
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': pool.kind, 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index: int):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': reasons, 'samples': len(self.rows)}


def kernel_source_sha() -> str:
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / 'marl_factory_grid_b200' / 'csrc').glob('*.cu*')) + sorted((ROOT / 'marl_factory_grid_b200' / 'csrc').glob('*.h*')):
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def hbm_peak():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        try:
            return float(json.loads(p.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def measure(args, parity, steps, warmup, dev, rank, world, local, with_e2e):
    """One timed pass in one parity mode.  Returns the measurements of this rank (times max-reduced over ranks)."""
    import torch
    import torch.distributed as dist
    from marl_factory_grid_b200 import FactoryConfigParser
    from marl_factory_grid_b200.distributed import allreduce_max, allreduce_stats
    from marl_factory_grid_b200.engine import Engine

    es = FactoryConfigParser(CONFIGS / f'{args.config}.yaml').compile()
    n_local = args.envs_per_gpu
    A = es.n_agents
    eng = Engine(es, n_local, device=dev, faithful=parity == 'faithful', seed=es.env_seed, env_id_offset=rank * n_local)
    if args.obs_kernel:
        eng.set_option('obs_kernel', args.obs_kernel)
    eng.set_option('obs_store', args.obs_store)
    acts = torch.zeros((n_local, A), dtype=torch.int32, device=dev)
    eng.reset()

    def one_step(i):
        eng.random_actions(acts, seed=0, step_index=i)
        eng.step_observe(acts, auto_reset=True)      # ONE C-ABI call: k_step, then packed re-spawn || k_obs_tiled, redo passes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    n_age = aged_steps(args)
    for i in range(n_age):                 # untimed, not part of `steps` / `warmup` (config.aged_steps)
        one_step(i)
    for i in range(warmup):
        one_step(n_age + i)
    barrier()
    warmup = n_age + warmup                # step index offset of the timed region
    graph = None
    if args.graph:
        # launch-bound batch sizes: the fused call as ONE CUDA-graph launch.  Kernel durations cannot be recorded inside a
        # graph, so they are taken from a short eager pass first (same state, library event pairs).
        eng.set_option('timing', 1)
        for i in range(20):
            one_step(warmup + i)
        barrier()
        eager_ms = {k: eng.info(k) * 1e-6 / 20 for k in ('obs_ns', 'step_ns', 'reset_ns')}
        eng.set_option('timing', 0)
        graph = eng.capture_step(acts, auto_reset=True)

        def one_step(i):                      # noqa: F811
            eng.random_actions(acts, seed=0, step_index=i)
            graph.replay()
    launches0 = eng.info('launches')
    # per-kernel durations: CUDA event pairs recorded by the library around its launches, on the launching streams,
    # during the timed region below (mfg_set_option "timing"; read back after the region)
    eng.set_option('timing', 1)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        start.record()
        for i in range(steps):
            one_step(warmup + i)
        stop.record()
        barrier()
    elapsed_ms = start.elapsed_time(stop)
    obs_ms = eng.info('obs_ns') * 1e-6 / max(steps, 1)
    step_ms = eng.info('step_ns') * 1e-6 / max(steps, 1)
    reset_ms = eng.info('reset_ns') * 1e-6 / max(steps, 1)
    eng.set_option('timing', 0)
    launches = eng.info('launches') - launches0
    if graph is not None:
        obs_ms, step_ms, reset_ms = eager_ms['obs_ns'], eager_ms['step_ns'], eager_ms['reset_ns']
        launches = 8 * steps                  # kernels inside each replayed graph (+ k_random_actions), not counted by the library

    # episode statistics: the only cross-GPU exchange of the path (one small all-reduce over NVLink)
    stats = allreduce_stats(eng.stats(), device=dev)
    elapsed_ms, obs_ms = allreduce_max(elapsed_ms, dev), allreduce_max(obs_ms, dev)
    step_ms, reset_ms = allreduce_max(step_ms, dev), allreduce_max(reset_ms, dev)
    rest_ms = max(elapsed_ms / max(steps, 1) - obs_ms - step_ms, 0.0)

    # ---- e2e: host buffers through the C-ABI host entry point (H2D actions, D2H reward + done + obs inside the timed region)
    e2e = None
    if with_e2e:
        k2 = max(1, min(steps, args.e2e_steps))
        try:                             # pinned staging buffers: 7.4 GB of observations per rank
            h_act = torch.zeros((n_local, A), dtype=torch.int32).pin_memory()
            h_rew = torch.zeros((n_local, eng.n_rew), dtype=torch.float32).pin_memory()
            h_done = torch.zeros(n_local, dtype=torch.uint8).pin_memory()
            h_obs = torch.zeros((n_local, es.total_channels) + tuple(es.obs_shape), dtype=torch.float32).pin_memory()
            ok = 1.0
        except Exception:                # e.g. not enough lockable host memory for 8 ranks on one box
            ok = 0.0
        if allreduce_max(-ok, dev) < 0:  # every rank got its buffers
            gen = torch.Generator().manual_seed(rank)
            pool = [torch.stack([torch.randint(0, n, (n_local,), generator=gen, dtype=torch.int32) for n in es.n_actions], 1)
                    for _ in range(2)]
            eng.step_host(h_act, h_rew, h_done, h_obs, auto_reset=True)       # warm-up (allocates the staging buffers)
            barrier()
            t0 = time.perf_counter()
            for i in range(k2):
                h_act.copy_(pool[i % 2])                                        # the caller's host-side actions
                eng.step_host(h_act, h_rew, h_done, h_obs, auto_reset=True)
            barrier()
            t_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
            e2e = {'value': world * n_local * A * k2 / float(t_e2e[0]), 'unit': UNIT, 'steps': k2,
                   'h2d_bytes_per_step': int(h_act.numel() * 4) * world,
                   'd2h_bytes_per_step': int(h_rew.numel() * 4 + h_done.numel() + h_obs.numel() * 4) * world,
                   'note': 'mfg_step_host: pinned host actions in, reward + done + full observation tensor out, synchronous '
                           '(PCIe-bound: the dense f32 observation tensor is 7 KB per env)'}
            # second view: a GPU-resident learner reads the observation tensor in place; only actions go in and reward /
            # done come out over PCIe (mfg_step_host with h_obs = NULL).  Reported beside e2e, not instead of it.
            barrier()
            t0 = time.perf_counter()
            for i in range(k2 * 4):
                h_act.copy_(pool[i % 2])
                eng.step_host(h_act, h_rew, h_done, None, auto_reset=True)
            barrier()
            t_dev = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
            e2e['obs_left_on_device'] = {'value': world * n_local * A * k2 * 4 / float(t_dev[0]), 'unit': UNIT, 'steps': k2 * 4,
                                         'h2d_bytes_per_step': int(h_act.numel() * 4) * world,
                                         'd2h_bytes_per_step': int(h_rew.numel() * 4 + h_done.numel()) * world}
            # what the link can do: pinned D2H copy of the same observation tensor, alone (the end-to-end call adds the compute
            # in front of it and the small H2D / D2H transfers: it is synchronous by contract, results land in the caller's buffers)
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(2):
                h_obs.copy_(eng.obs, non_blocking=True)
            ev1.record()
            torch.cuda.synchronize(dev)
            d2h_gbs = 2 * h_obs.numel() * 4 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
            d2h_gbs = -allreduce_max(-d2h_gbs, dev)          # slowest rank (all ranks copy at the same time)
            e2e['d2h_link'] = {'pinned_d2h_gbs_per_rank_concurrent': d2h_gbs,
                               'e2e_d2h_gbs_per_rank': e2e['d2h_bytes_per_step'] / world * k2 / float(t_e2e[0]) / 1e9,
                               'frac_of_link': e2e['d2h_bytes_per_step'] / world * k2 / float(t_e2e[0]) / 1e9 / d2h_gbs}
            del h_obs
        else:
            e2e = {'value': None, 'unit': UNIT, 'note': 'pinned host buffers could not be allocated on every rank'}

    total_envs = world * n_local
    env_steps_per_s = total_envs * steps / (elapsed_ms * 1e-3)
    peak, peak_src = hbm_peak()
    obs_bytes = 4 * es.obs_d ** 2 * es.total_channels + es.algorithmic_state_bytes()       # obs write + state read, per env
    step_bytes = es.algorithmic_bytes_per_env_step()
    achieved = obs_bytes * n_local / (obs_ms * 1e-3) / 1e9
    tiled = bool(eng.info('tiled_ok')) and args.obs_kernel != 1
    # dram bytes of one launch of the dominant kernel cannot be measured in-run (ncu only): the committed capture is used
    # when it was taken from the current kernel sources (sha256 stamp), otherwise the key is null
    traffic, traffic_note = None, 'no ncu capture for this configuration'
    tp = ROOT / 'profiles' / 'obs_kernel_traffic.json'
    if tp.exists():
        try:
            rec = json.loads(tp.read_text())
            if rec.get('source_sha256') == kernel_source_sha():
                traffic = rec.get(f'{args.config}:{parity}:{n_local}')
                traffic_note = f'ncu --set full capture, {rec.get("captured", "?")}' if traffic else traffic_note
            else:
                traffic_note = 'committed ncu capture is older than the kernel sources: dropped'
        except Exception:
            traffic = None
    out = {
        'A': A, 'parity': parity, 'steps': steps, 'value': env_steps_per_s * A, 'env_steps_per_s': env_steps_per_s,
        'ms_per_step': elapsed_ms / max(steps, 1), 'tiled': tiled, 'step_bytes': step_bytes, 'n_local': n_local,
        'roofline': {'bound': 'hbm', 'kernel': 'k_obs_tiled' if tiled else 'k_obs_direct',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'traffic_note': traffic_note,
                     'peak_source': peak_src, 'bytes_per_env': obs_bytes, 'ms_per_launch': obs_ms,
                     'whole_step': {'bytes_per_env_step': step_bytes,
                                    'achieved': step_bytes * env_steps_per_s / world / 1e9,
                                    'frac': step_bytes * env_steps_per_s / world / 1e9 / peak}},
        'gpu_launches': launches,
        'obs_launch': {'threads': eng.info('obs_threads'), 'dyn_smem': eng.info('obs_smem'), 'ctas_per_sm': eng.info('obs_ctas_per_sm')},
        'kernel_ms': {'k_step': step_ms, 'k_obs_tiled+redo': obs_ms, 'k_reset_list (side stream, followed there by k_obs_tiled in list mode; overlaps the main k_obs_tiled)': reset_ms,
                      'rest (k_random_actions, joins, gaps)': rest_ms},
        'clocks': clocks.summary(),
        'episode_stats': {'episodes': int(stats[0]), 'env_steps_in_finished_episodes': int(stats[1]),
                          'collisions': int(stats[8]), 'dirt_overflow': int(stats[9]), 'spawn_fail': int(stats[10])},
        'e2e': e2e,
    }
    eng.close()
    del eng, acts
    torch.cuda.empty_cache()
    return out


def bind_to_gpu_cpus(local: int) -> str:
    """Pins this rank to the host cores that are local to its GPU (NVML's ideal CPU affinity = the GPU's NUMA node), so that
    the pinned host buffers of the end-to-end path are allocated and filled on that node: by default every rank inherits the
    same affinity and 8 ranks' 7.4 GB/step observation copies cross one socket.  Returns a short description for the JSON."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0)) or cpus
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f'{min(cpus)}-{max(cpus)} ({len(cpus)} cores, NVML affinity of GPU {local})'
    except Exception as exc:                # no NVML / no permission: keep the inherited affinity
        return f'inherited ({type(exc).__name__})'
    return 'inherited'


def run_engine(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    inherited = os.sched_getaffinity(0)
    affinity = bind_to_gpu_cpus(local)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    m = measure(args, args.parity, args.steps, args.warmup, dev, rank, world, local, with_e2e=not args.no_e2e)
    other = None
    if not args.no_other_mode:
        other_parity = 'identity' if args.parity == 'faithful' else 'faithful'
        other = measure(args, other_parity, args.steps, args.warmup, dev, rank, world, local, with_e2e=False)

    os.sched_setaffinity(0, inherited)       # the CPU baseline below uses every host core again
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    n_local, A = m['n_local'], m['A']
    line = {
        'metric': METRIC, 'value': m['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': m['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u16/f64 state, f32 obs', 'data': 'synthetic',
        'config': workload_config(args, world),
        'engine': {'obs_kernel': 'tiled' if m['tiled'] else 'direct', 'cuda_graph': bool(args.graph)},
        'env_steps_per_s': m['env_steps_per_s'],
        'roofline': m['roofline'], 'gpu_launches': m['gpu_launches'], 'obs_launch': m['obs_launch'], 'kernel_ms': m['kernel_ms'],
        'clocks': m['clocks'], 'episode_stats': m['episode_stats'],
    }
    if other is not None:
        line['other_parity_mode'] = {k: other[k] for k in ('parity', 'steps', 'value', 'env_steps_per_s', 'ms_per_step', 'kernel_ms')}
        line['other_parity_mode']['roofline'] = {k: other['roofline'][k] for k in ('kernel', 'achieved', 'frac', 'ms_per_launch', 'whole_step')}
    if m['e2e'] is not None:
        line['e2e'] = m['e2e']
        line['e2e']['host_cpu_affinity'] = affinity
    if world == 1 and not args.no_cpu:
        procs = os.cpu_count() or 1
        faithful = args.parity == 'faithful'
        try:
            v, per, slowest, kind = cpu_sample(args.config, faithful, procs, 'auto', args.cpu_seconds)
        except Exception as exc:          # a crash inside the reference must not cost the GPU line: time the oracle port instead
            print(f'[bench] reference CPU sample failed ({exc!r}); timing the oracle port', file=sys.stderr)
            v, per, slowest, kind = cpu_sample(args.config, faithful, procs, 'port', args.cpu_seconds)
        what = 'UNMODIFIED reference from baseline/_ref, SURVEY 8d protocol' if kind == 'reference' else 'oracle port of the reference step+obs'
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': procs, 'kind': kind,
                                'sample': f'{procs} procs x {per} env-steps of {args.config}, {what} (own Factory per worker, '
                                          f'random actions, in-place reset on done), {slowest:.1f} s'}
        if kind == 'reference':
            # the port beside it (what round 1 reported), and the other BASELINE configs on the reference (short samples)
            try:
                pv, pper, ps, _ = cpu_sample(args.config, faithful, procs, 'port', min(args.cpu_seconds, 6.0))
                line['cpu_baseline']['port'] = {'value': pv, 'sample': f'{procs} procs x {pper} env-steps, oracle port, {ps:.1f} s'}
                others = {}
                for c in ('cfg1', 'cfg2', 'cfg3'):
                    if c != args.config and not args.no_cpu_sweep:
                        ov, oper, os_, _ = cpu_sample(c, faithful, procs, 'reference', 4.0)
                        others[c] = {'value': ov, 'sample': f'{procs} procs x {oper} env-steps, {os_:.1f} s'}
                line['cpu_baseline']['other_configs'] = others
            except Exception as exc:
                line['cpu_baseline']['note'] = f'secondary CPU samples failed: {exc!r}'

    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=100)
    ap.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    ap.add_argument('--config', default='cfg4')
    ap.add_argument('--envs-per-gpu', type=int, default=1 << 20)
    ap.add_argument('--parity', default='faithful', choices=['identity', 'faithful'])
    ap.add_argument('--obs-kernel', type=int, default=0, help='0 auto, 1 direct, 2 tiled')
    ap.add_argument('--obs-store', type=int, default=1, help='1 TMA bulk store of the tile, 0 LDS/STG loop')
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='wall-clock budget of the CPU baseline sample (bounded)')
    ap.add_argument('--no-cpu-sweep', action='store_true', help='skip the short reference samples of the other BASELINE configs')
    ap.add_argument('--age', type=int, default=-1, help='untimed ageing env-steps before the warm-up (-1: 300 - warmup, at least 0)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--graph', action='store_true', help='replay the fused step as one CUDA graph (launch-bound batch sizes)')
    ap.add_argument('--no-other-mode', action='store_true', help='skip the short second pass in the other parity mode')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'engine':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_engine(args)


if __name__ == '__main__':
    main()
