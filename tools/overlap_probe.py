"""Experiment: do two half-batches on two streams overlap the latency-bound step kernel with the bandwidth-bound
observation kernel?  (not part of the product path)"""
import sys, time
sys.path.insert(0, '.')
import torch
from marl_factory_grid_b200 import FactoryConfigParser
from marl_factory_grid_b200.engine import Engine

es = FactoryConfigParser('marl_factory_grid_b200/configs/cfg4.yaml').compile()
N = 1 << 20
def run(parts, steps=40, warm=8):
    n = N // parts
    engs = [Engine(es, n, device='cuda:0', faithful=False, seed=69, env_id_offset=i * n) for i in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    acts = [torch.zeros((n, 4), dtype=torch.int32, device='cuda:0') for _ in range(parts)]
    for e, s in zip(engs, streams):
        with torch.cuda.stream(s):
            e.reset()
    torch.cuda.synchronize()
    def step(i):
        for e, s, a in zip(engs, streams, acts):
            with torch.cuda.stream(s):
                e.random_actions(a, 0, i)
                e.step(a, auto_reset=True)
                e.observe()
    for i in range(warm): step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps): step(warm + i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f'parts={parts}: {dt*1e3:.3f} ms/step  {N*4/dt:.3e} agent-steps/s')
    for e in engs: e.close()
for p in (1, 2, 4, 8):
    run(p)
