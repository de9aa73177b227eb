"""Development aid: a short free-running rollout in both parity modes (for compute-sanitizer memcheck / racecheck)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from marl_factory_grid_b200 import FactoryConfigParser
from marl_factory_grid_b200.engine import Engine
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
es = FactoryConfigParser(ROOT / 'marl_factory_grid_b200' / 'configs' / f'{cfg}.yaml').compile()
for parity in ('identity', 'faithful'):
    eng = Engine(es, N, device='cuda:0', faithful=parity == 'faithful', seed=3)
    acts = torch.zeros((N, es.n_agents), dtype=torch.int32, device='cuda:0')
    eng.reset()
    for t in range(steps):
        eng.random_actions(acts, seed=1, step_index=t)
        eng.step_observe(acts, auto_reset=True)
    torch.cuda.synchronize()
    print(parity, 'episodes', int(eng.stats()[0]), 'obs sum', float(eng.obs.sum()))
    eng.close()
