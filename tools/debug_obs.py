"""Development aid: replay the golden episodes of a config on the GPU and print the first cell where the tiled
observation kernel differs from the direct kernel, with the entities involved."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
from golden_util import episodes, snap_at, spec_for
from hostsim_util import tape_respawn
from marl_factory_grid_b200.engine import Engine

cfg, mode = sys.argv[1], sys.argv[2]
es = spec_for(cfg)
eps = [ep for ep in episodes(cfg) if ep['meta']['mode'] == mode]
N, A, NM = len(eps), es.n_agents, es.n_maint
eng = Engine(es, N, device='cuda:0', faithful=mode == 'U')
for i, ep in enumerate(eps):
    eng.load_snapshot(i, snap_at(ep, 0))
T = [len(ep['actions']) for ep in eps]
nbad = 0
for t in range(max(T)):
    acts = np.zeros((N, A), np.int32); ma = np.full((N, max(NM, 1)), 8, np.uint8); rn = np.zeros(N, np.int8); rp = np.zeros((N, 8), np.uint16)
    for i, ep in enumerate(eps):
        if t < T[i]:
            acts[i] = ep['actions'][t]
            if NM: ma[i, :NM] = ep['maint_act'][t]
            rn[i], rp[i] = tape_respawn(ep, t)
    eng.step(acts, tape=dict(maint_action=ma[:, :NM] if NM else None, respawn_n=rn, respawn_pos=rp))
    eng.set_option('obs_kernel', 1); o1 = eng.observe().cpu().numpy().copy()
    eng.set_option('obs_kernel', 2); o2 = eng.observe().cpu().numpy().copy()
    if not np.array_equal(o1, o2):
        fields = eng.fields_numpy()
        for i in range(N):
            if t >= T[i] or np.array_equal(o1[i], o2[i]): continue
            bad = np.argwhere(o1[i] != o2[i])
            print(f't={t+1} env{i}: {len(bad)} cells differ; first (ch,x,y)={bad[0]} direct={o1[i][tuple(bad[0])]} tiled={o2[i][tuple(bad[0])]}')
            s = eng.snapshot(i, fields)
            for k in ('agent_pos', 'dirt_pos', 'dirt_uid', 'dirt_listed', 'item_pos', 'item_listed', 'maint_pos', 'maint_listed', 'door_open', 'door_listed'):
                print('   ', k, np.asarray(s[k]).tolist())
            nbad += 1
            if nbad > 3: sys.exit(0)
print('done, mismatching envs:', nbad)
