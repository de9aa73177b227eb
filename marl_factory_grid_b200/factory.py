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


class Factory:
    def __init__(self, config_file: Union[str, Path], custom_modules_path=None, custom_level_path=None,
                 n_envs: Optional[int] = None, device='cuda', parity: str = 'faithful', auto_reset: bool = False,
                 seed: Optional[int] = None, env_id_offset: int = 0, dirt_slots: int = 40):
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
        per_agent = self._split(obs)
        if self.unbatched:
            r = reward[0].cpu().numpy().astype(np.float64)
            rew = [float(x) for x in r]
            step = int(self.engine.fields['step'][0, 0, 0].item()) & 0xFFFF
            info = dict(step_reward=float(np.sum(r)), step=step)
            return None, [o[0].cpu().numpy() for o in per_agent], rew, bool(done[0].item()), info
        return None, per_agent, reward, done.bool(), {}

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
