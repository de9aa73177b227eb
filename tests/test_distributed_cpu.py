"""World-size-2 gloo test of the multi-process path (CPU): env sharding by global env id + the statistics all-reduce.
The per-env code is the test-only host build of the device functions, so this runs without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_util import spec_for
from marl_factory_grid_b200.distributed import allreduce_max, allreduce_stats, shard

N_TOTAL, STEPS, CFG, SEED = 37, 60, 'cfg4', 5      # 37 envs over 2 ranks: ragged shards (19 + 18)


def _actions(t, n_total, n_act):
    rng = np.random.default_rng(1000 + t)
    return np.stack([rng.integers(0, n, n_total) for n in n_act], 1).astype(np.int32)


def _run(offset, n_local, n_total):
    from hostsim_util import HostSim
    es = spec_for(CFG)
    sim = HostSim(es, n_local, faithful=True, seed=SEED, env_id_offset=offset)
    sim.reset()
    obs_sum = 0.0
    for t in range(STEPS):
        a = _actions(t, n_total, es.n_actions)[offset:offset + n_local]
        sim.step(a, auto_reset=True)
        obs_sum += float(sim.observe().astype(np.float64).sum())
    return sim.stats(), sim.fields['apos'].copy(), obs_sum


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        offset, n_local = shard(rank, world, N_TOTAL)
        stats, apos, obs_sum = _run(offset, n_local, N_TOTAL)
        total = allreduce_stats(stats)
        slowest = allreduce_max(float(rank + 1))
        gathered = [None] * world
        dist.all_gather_object(gathered, (offset, apos, obs_sum))
        if rank == 0:
            q.put((total, slowest, gathered))
    finally:
        dist.destroy_process_group()


def test_shard_partition_is_contiguous_and_complete():
    for world in (1, 2, 3, 8):
        parts = [shard(r, world, N_TOTAL) for r in range(world)]
        assert parts[0][0] == 0 and sum(n for _, n in parts) == N_TOTAL
        for (o1, n1), (o2, _) in zip(parts, parts[1:]):
            assert o1 + n1 == o2


@pytest.mark.timeout(300)
def test_two_rank_run_equals_single_process_run():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, slowest, gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_stats, ref_apos, ref_obs = _run(0, N_TOTAL, N_TOTAL)
    assert slowest == 2.0
    # integer counters add up exactly; the f64 return sums up to reassociation
    np.testing.assert_array_equal(total[:11], ref_stats[:11])
    np.testing.assert_allclose(total[11:].view(np.float64), ref_stats[11:].view(np.float64), rtol=1e-12, atol=1e-9)
    assert total[0] > 0                       # episodes finished (and were re-spawned) during the run
    apos = np.concatenate([g[1] for g in sorted(gathered, key=lambda g: g[0])], axis=1)
    np.testing.assert_array_equal(apos, ref_apos)     # env for env identical to the unsharded run
    assert abs(sum(g[2] for g in gathered) - ref_obs) < 1e-6
