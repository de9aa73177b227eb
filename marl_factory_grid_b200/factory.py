"""Batched `Factory`: the reference's gym-style surface (marl_factory_grid/environment/factory.py:22-298)
over N independent environments that live on one B200.

Same constructor arguments, same yaml schema, same level files, same named action / observation spaces
and the same per-agent observation and reward layout - with a leading batch dimension:

    f = Factory('cfg4.yaml', n_envs=4096)
    obs = f.reset()                       # {agent_name: float32 tensor [N, C_a, D, D]}  (views of one packed tensor)
    _, obs, reward, done, info = f.step(actions)   # actions int32 [N, A]; reward f32 [N, A]; done bool [N]

`n_envs=None` gives the reference's un-batched shapes (numpy arrays, python lists, a bool) for drop-in use.
Differences to the reference, by design: config errors raise instead of `exit()`, gymnasium is
not required, `render()` is not provided, custom Python modules are rejected, `auto_reset=True` re-spawns
finished environments inside the step kernel (off by default, as in the reference).
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Optional, Union

import numpy as np

from .config_parser import FactoryConfigParser, named_action_space
from .engine import Engine
from .spec import EnvSpec


class Discrete:
    """Minimal stand-in for gymnasium.spaces.Discrete (gymnasium is optional)."""

    def __init__(self, n):
        self.n = int(n)

    def __repr__(self):
        return f'Discrete({self.n})'


class Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return f'Box({self.low}, {self.high}, {self.shape}, {np.dtype(self.dtype).name})'


def step_info(spec, flags, actions) -> dict:
    """The per-agent part of the reference's `info` dict for ONE env (utils/results.py:42-52 `get_infos`, folded by
    `summarize_step_results`, factory.py:222-239), rebuilt from the engine's per-step result flags (`mfg_bind_step_flags`):

      `"{agent}_{ActionClass}"`  the action's reward (valid / fail / ItemAction's drop-off reward); absent for a paralysed agent
      `"{agent}_Collisions"`     +1 when the agent's move "introduced a collision" (results.py:76-83), plus the WatchCollisions
                                 reward when the agent stands on a collision tile (rules.py:293-305) - the reference sums both
                                 under the one key

    Rule-value entries (`Global_DirtPiles_spawn`, `Global_DoorAutoClose`, `*_BatteryDecharge`, ...) are not produced.
    flags: uint8 [A + 1] (one row of the flags tensor), actions: int [A]."""
    from . import abi
    info = {}
    coll = next((r.params[0] for r in spec.rules if r.name == 'WatchCollisions'), None)
    for i, ag in enumerate(spec.agents):
        f = int(flags[i])
        if not f & abi.FLAG_SKIPPED:
            a = int(actions[i]) if 0 <= int(actions[i]) < len(ag.actions) else 0
            act = ag.actions[a]
            info[f'{ag.name}_{act.class_name}'] = float(act.aux_reward if f & abi.FLAG_AUX_REWARD else
                                                        act.valid_reward if f & abi.FLAG_VALID else act.fail_reward)
        c = (1.0 if f & abi.FLAG_MOVE_COLLISION else 0.0) + (float(coll) if f & abi.FLAG_COLLISION and coll is not None else 0.0)
        if f & (abi.FLAG_MOVE_COLLISION | abi.FLAG_COLLISION):
            info[f'{ag.name}_Collisions'] = c
    return info


class Factory:
    def __init__(self, config_file: Union[str, Path], custom_modules_path=None, custom_level_path=None,
                 n_envs: Optional[int] = None, device='cuda', parity: str = 'faithful', auto_reset: bool = False,
                 seed: Optional[int] = None, env_id_offset: int = 0, dirt_slots: int = 40, info: Optional[bool] = None):
        if parity not in ('faithful', 'identity'):
            raise ValueError("parity must be 'faithful' (untouched reference) or 'identity' (uid artefact off)")
        self._config_file = config_file
        self.conf = FactoryConfigParser(config_file, custom_modules_path)
        self.spec: EnvSpec = self.conf.compile(custom_level_path=custom_level_path, dirt_slots=dirt_slots)
        # `individual_rewards: false` is dead in the reference: construction and reset() work, the first step() raises
        # TypeError at factory.py:217 (`sum(reward)` of a float).  Mirrored: the engine always folds per agent.
        self._scalar_reward_defect = not self.spec.individual_rewards
        self.spec.individual_rewards = True
        self.unbatched = n_envs is None
        self.n_envs = 1 if n_envs is None else int(n_envs)
        self.parity = parity
        self.auto_reset = bool(auto_reset)
        self.engine = Engine(self.spec, self.n_envs, device=device, faithful=parity == 'faithful', seed=seed,
                             env_id_offset=env_id_offset)
        self.level_shape = (self.spec.H, self.spec.W)
        # per-step result flags (the batched form of the reference's `info`): on by default for the un-batched drop-in shape
        self._info = self.unbatched if info is None else bool(info)
        self._flags = self.engine.enable_step_flags() if self._info else None
        self._offsets = self.spec.channel_offsets
        self._needs_reset = True

    # ------------------------------------------------------------------ spaces (factory.py:24-63)
    @property
    def agent_names(self) -> List[str]:
        return [a.name for a in self.spec.agents]

    @property
    def action_space(self):
        return [Discrete(n) for n in self.spec.n_actions]

    @property
    def named_action_space(self) -> Dict[str, Dict[str, int]]:
        return named_action_space(self.spec)

    @property
    def observation_space(self):
        shape = tuple(self.spec.obs_shape)
        boxes = [Box(0, 1, (c,) + shape, np.float32) for c in self.spec.channels_per_agent]
        return boxes[0] if len(boxes) == 1 else boxes

    @property
    def named_observation_space(self) -> Dict[str, List[str]]:
        return {a.name: [ch.name for ch in a.channels] for a in self.spec.agents}

    @property
    def params(self) -> dict:
        return dict(self.conf.config)

    # ------------------------------------------------------------------ gym surface
    def _split(self, obs):
        return [obs[:, o:o + c] for o, c in zip(self._offsets, self.spec.channels_per_agent)]

    def reset(self, mask=None):
        """factory.py:134-148.  Returns {agent_name: observation}."""
        self.engine.reset(mask)
        self._needs_reset = False
        per_agent = self._split(self.engine.observe())
        if self.unbatched:
            return {n: o[0].cpu().numpy() for n, o in zip(self.agent_names, per_agent)}
        return dict(zip(self.agent_names, per_agent))

    def step(self, actions, tape=None):
        """factory.py:189-220.  Returns (None, [obs per agent], reward, done, info)."""
        if self._needs_reset:
            raise RuntimeError('call reset() before step()')
        if self._scalar_reward_defect:
            raise TypeError("'float' object is not iterable")      # factory.py:217 with individual_rewards: false
        if self.unbatched and not hasattr(actions, 'shape'):
            actions = np.asarray(actions if isinstance(actions, (list, tuple)) else [int(actions)], np.int32)[None]
        obs, reward, done = self.engine.step_observe(actions, tape=tape, auto_reset=self.auto_reset)
        self._last_actions = np.asarray(actions).reshape(-1) if self.unbatched else None
        per_agent = self._split(obs)
        if self.unbatched:
            r = reward[0].cpu().numpy().astype(np.float64)
            rew = [float(x) for x in r]
            step = int(self.engine.fields['step'][0, 0, 0].item()) & 0xFFFF
            info = self._info_dict(actions, r) if self._info else {}
            info.update(step_reward=float(np.sum(r)), step=step)
            return None, [o[0].cpu().numpy() for o in per_agent], rew, bool(done[0].item()), info
        return None, per_agent, reward, done.bool(), ({'flags': self._flags} if self._info else {})

    def _info_dict(self, actions, reward) -> dict:
        acts = np.asarray(actions.cpu() if hasattr(actions, 'cpu') else actions).reshape(-1)
        return step_info(self.spec, self._flags[0].cpu().numpy(), acts)

    # ------------------------------------------------------------------ entity access / state summaries
    def __getitem__(self, item):
        """factory.py:131 (`env[group_name]` -> the entity group).  Batched form: a dict of tensors, e.g.
        env['Agent'] = {'pos': int16 [N, A, 2]}, env['DirtPiles'] = {'pos': [N, slots, 2], 'amount': f64 [N, slots],
        'alive': bool [N, slots]}, env['Doors'] = {'pos': [ND, 2], 'open': bool [N, ND], 'time_to_close': uint8 [N, ND]}."""
        t = self.engine.torch
        eng, es = self.engine, self.spec

        def pos_of(field):
            p = eng.field(field).to(t.int64) & 0xFFFF                    # [rows, N]
            xy = t.stack([p >> 8, p & 255], -1).permute(1, 0, 2).to(t.int16)
            return xy, (p != 0xFFFF).permute(1, 0)

        if item in ('Agent', 'Agents'):
            return {'pos': pos_of('apos')[0], 'names': self.agent_names}
        small = {'Items': 'item_pos', 'ChargePods': 'pod_pos', 'Destinations': 'dest_pos', 'DropOffLocations': 'drop_pos',
                 'Machines': 'mach_pos', 'Maintainers': 'maint_pos'}
        if item in small and small[item] in eng.fields:
            xy, on_map = pos_of(small[item])
            out = {'pos': xy, 'on_map': on_map}
            if item == 'Destinations':
                r = eng.field('dest_reached')[0].to(t.int64)
                out['reached'] = ((r[:, None] >> t.arange(es.n_dest, device=r.device)) & 1).bool()
            return out
        if item == 'DirtPiles' and es.has_dirt:
            xy, alive = pos_of('dirt_pos')
            return {'pos': xy, 'alive': alive, 'amount': eng.field('dirt_amt').permute(1, 0)}
        if item == 'Doors' and es.n_doors:
            o = eng.field('door_open')[0].to(t.int64)
            return {'pos': t.as_tensor(es.door_pos.astype(np.int16)),
                    'open': ((o[:, None] >> t.arange(es.n_doors, device=o.device)) & 1).bool(),
                    'time_to_close': eng.field('door_timer').permute(1, 0)}
        if item == 'Batteries' and es.has_batteries:
            return {'charge_level': eng.field('bat').permute(1, 0)}
        if item == 'Walls':
            return {'pos': t.as_tensor(np.argwhere(es.walls).astype(np.int16))}
        raise KeyError(item)

    def summarize_header(self) -> dict:
        """factory.py:269-273: the static part of a recording (walls)."""
        walls = [dict(name=f'Wall[{k}]', x=int(x), y=int(y), can_collide=True) for k, (x, y) in enumerate(np.argwhere(self.spec.walls))]
        return {'rec_step': int(self.engine.fields['step'][0, 0, 0].item()) & 0xFFFF, 'recWalls': walls}

    def summarize_state(self, env: int = 0) -> dict:
        """factory.py:275-292 for one env of the batch: `{step, walls, agents, doors, items, batteries}` with the reference's
        per-entity keys (`EnvRecorder` / the renderer read these)."""
        s = self.engine.snapshot(env)
        es = self.spec
        out = {'step': int(s['step']),
               'walls': [dict(name=f'Wall[{k}]', x=int(x), y=int(y), can_collide=True) for k, (x, y) in enumerate(np.argwhere(es.walls))],
               'agents': [dict(name=a.name, x=int(p[0]), y=int(p[1]), can_collide=True) for a, p in zip(es.agents, s['agent_pos'])]}
        if self._flags is not None and self.unbatched and getattr(self, '_last_actions', None) is not None:
            fl = self._flags[0].cpu().numpy()
            for i, (rec, ag) in enumerate(zip(out['agents'], es.agents)):
                skipped = bool(fl[i] & 2)
                a = int(self._last_actions[i])
                rec.update(valid=True if skipped else bool(fl[i] & 1), action='Noop' if skipped else ag.actions[a].class_name)
        if es.n_doors:
            out['doors'] = [dict(name=f'Door[{k}]', x=int(p[0]), y=int(p[1]), can_collide=not bool(o),
                                 state='open' if o else 'closed', time_to_close=int(tm))
                            for k, (p, o, tm) in enumerate(zip(es.door_pos, s['door_open'], s['door_timer']))]
        if es.has_batteries:
            out['batteries'] = [dict(belongs_to=a.name, chargeLevel=float(b)) for a, b in zip(es.agents, s['battery'])]
        if es.n_items:
            out['items'] = [dict(name=f'Item[{k}]', x=int(p[0]), y=int(p[1]), can_collide=False) for k, p in enumerate(s['item_pos'])]
        return out

    def episode_stats(self, zero_after=False) -> dict:
        """Device-side episode statistics (replaces the pandas EnvMonitor, utils/logging/envmonitor.py:28-56)."""
        from . import abi
        s = self.engine.stats(zero_after)
        as_f64 = s.view(np.float64)
        n = max(int(s[abi.ST_EPISODES]), 1)
        return {'episodes': int(s[abi.ST_EPISODES]), 'steps': int(s[abi.ST_STEPS]),
                'mean_length': float(s[abi.ST_STEPS]) / n,
                'done_max_steps': int(s[abi.ST_DONE_MAX_STEPS]), 'done_all_dirt': int(s[abi.ST_DONE_ALL_DIRT]),
                'done_battery': int(s[abi.ST_DONE_BATTERY]), 'done_destination': int(s[abi.ST_DONE_DEST]),
                'done_maintainer': int(s[abi.ST_DONE_MAINT]), 'done_collision': int(s[abi.ST_DONE_COLLISION]),
                'collisions': int(s[abi.ST_COLLISIONS]), 'dirt_overflow': int(s[abi.ST_DIRT_OVERFLOW]),
                'spawn_fail': int(s[abi.ST_SPAWN_FAIL]), 'return_sum': float(as_f64[abi.ST_RETURN_SUM]),
                'mean_return_per_agent': [float(as_f64[abi.ST_RETURN_AGENT0 + i]) / n for i in range(self.spec.n_agents)]}

    # ------------------------------------------------------------------ state access
    def load_state(self, env: int, snapshot: dict):
        self.engine.load_snapshot(env, snapshot)
        self._needs_reset = False

    def export_state(self, env: int) -> dict:
        return self.engine.snapshot(env)

    def save_params(self, filepath):
        import shutil
        filepath = Path(filepath)
        filepath.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(self._config_file, filepath)

    def close(self):
        self.engine.close()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.close()
