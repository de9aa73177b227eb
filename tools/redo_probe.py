"""Development aid: how many envs per step take the exact (redo) observation path, per parity mode."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from marl_factory_grid_b200 import FactoryConfigParser
from marl_factory_grid_b200.engine import Engine
es = FactoryConfigParser(ROOT / 'marl_factory_grid_b200' / 'configs' / (sys.argv[1] if len(sys.argv) > 1 else 'cfg4.yaml')).compile()
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
for parity in ('identity', 'faithful'):
    eng = Engine(es, N, device='cuda:0', faithful=parity == 'faithful', seed=es.env_seed)
    acts = torch.zeros((N, es.n_agents), dtype=torch.int32, device='cuda:0')
    eng.reset()
    counts = []
    for t in range(260):
        eng.random_actions(acts, seed=0, step_index=t)
        eng.step_observe(acts, auto_reset=True)
        if t % 20 == 19:
            counts.append(eng.info('obs_redo_count'))
    print(parity, counts)
    eng.close()
