"""Batched EnvMonitor: per-episode records from the batched `Factory` (reference: marl_factory_grid/utils/logging/
envmonitor.py:15-73).

The reference wrapper collects the `info` dict of every step and, at `done`, aggregates the episode into one pandas row
(`sum` per column, `mean` for `*ount` columns) with a running `episode` index; `save_monitor` pickles the frame
(`reset_index()`).  Here N environments step at once on the GPU, so the per-episode sums are accumulated on the device
(f64, one row per env) and only the rows of the episodes that finished in a step travel to the host:

    columns: episode, env, steps, step_reward (sum over agents, = the reference's `step_reward` column summed over the
             episode), `<agent name>_reward` per agent (individual_rewards) or `reward`

The per-action / per-rule validity counters of the reference's `info` are not tracked by the engine.
"""
from __future__ import annotations

import pickle
from pathlib import Path
from typing import List, Optional, Union

import numpy as np


class EnvMonitor:
    ext = 'png'

    def __init__(self, env, filepath: Union[str, Path, None] = None):
        import torch
        self.env, self._filepath, self._torch = env, filepath, torch
        n, dev = env.n_envs, env.engine.device
        self._n_rew = env.engine.n_rew
        self._ret = torch.zeros((n, self._n_rew), dtype=torch.float64, device=dev)
        self._len = torch.zeros(n, dtype=torch.int64, device=dev)
        self._rows: List[np.ndarray] = []
        self._episodes = 0

    # reference surface -------------------------------------------------------------------------------------------
    def __getattr__(self, name):           # gymnasium.Wrapper forwards unknown attributes to the wrapped env
        return getattr(self.env, name)

    def reset(self, *a, **kw):
        self._ret.zero_()
        self._len.zero_()
        return self.env.reset(*a, **kw)

    def step(self, action):
        torch = self._torch
        obs_type, obs, reward, done, info = self.env.step(action)
        if self.env.unbatched:
            r = torch.as_tensor(np.atleast_1d(np.asarray(reward, np.float64)), device=self._ret.device)[None]
            d = torch.as_tensor([bool(done)], device=self._ret.device)
        else:
            r, d = reward.to(torch.float64), done.bool()
        self._ret += r
        self._len += 1
        if bool(d.any()):                   # one small D2H copy per step that finishes episodes
            idx = torch.nonzero(d, as_tuple=False).flatten()
            block = torch.cat([idx[:, None].to(torch.float64), self._len[idx, None].to(torch.float64), self._ret[idx]], 1)
            self._rows.append(block.cpu().numpy())
            self._ret[idx] = 0.0
            self._len[idx] = 0
        return obs_type, obs, reward, done, info

    # records -----------------------------------------------------------------------------------------------------
    @property
    def monitor_df(self):
        import pandas as pd
        names = ([f'{n}_reward' for n in self.env.agent_names] if self._n_rew > 1 else ['reward'])
        if not self._rows:
            return pd.DataFrame(columns=['episode', 'env', 'steps', 'step_reward'] + names)
        rows = np.concatenate(self._rows, 0)
        df = pd.DataFrame({'env': rows[:, 0].astype(np.int64), 'steps': rows[:, 1].astype(np.int64),
                           'step_reward': rows[:, 2:].sum(1)})
        for k, n in enumerate(names):
            df[n] = rows[:, 2 + k]
        df.insert(0, 'episode', np.arange(len(df)))
        return df

    def save_monitor(self, filepath: Union[Path, str, None] = None, auto_plotting_keys=None):
        """envmonitor.py:58-70: pickle of `monitor_df.reset_index()` (plotting is out of scope)."""
        filepath = Path(filepath or self._filepath)
        filepath.parent.mkdir(exist_ok=True, parents=True)
        with filepath.open('wb') as f:
            pickle.dump(self.monitor_df.reset_index(), f, protocol=pickle.HIGHEST_PROTOCOL)

    def report_possible_colum_keys(self):
        print(self.monitor_df.columns)


class EnvRecorder:
    """Episode recorder over the batched `Factory` (reference: marl_factory_grid/utils/logging/recorder.py:14-82): every
    `step` appends `env.summarize_state(env_index)` of ONE chosen env of the batch, `reset` starts a new episode with the
    header; `save_records` writes the list of episodes as yaml / json like the reference's non-protobuf path."""

    def __init__(self, env, env_index: int = 0, filepath: Union[str, Path, None] = None, episodes: Optional[List[int]] = None):
        self.env, self.env_index, self.filepath, self.episodes = env, int(env_index), filepath, episodes
        self._curr_episode = 0
        self._curr_ep_recorder, self._recorder_out_list = [], []

    def __getattr__(self, name):
        return getattr(self.env, name)

    def __getitem__(self, item):
        return self.env[item]

    def reset(self, *a, **kw):
        self._curr_ep_recorder, self._curr_episode = [], self._curr_episode + 1
        return self.env.reset(*a, **kw)

    def step(self, actions):
        out = self.env.step(actions)
        done = out[3]
        if self.episodes is None or self._curr_episode in self.episodes:
            self._curr_ep_recorder.append({'episode': self._curr_episode, **self.env.summarize_state(self.env_index)})
        finished = bool(done) if isinstance(done, bool) else bool(done[self.env_index])
        if finished and self._curr_ep_recorder:
            self._recorder_out_list.append({'steps': self._curr_ep_recorder, 'episode': self._curr_episode})
            self._curr_ep_recorder = []
        return out

    def save_records(self, filepath: Union[Path, str, None] = None, only_deltas: bool = False, save_occupation_map: bool = False,
                     save_trajectory_map: bool = False):
        import json
        filepath = Path(filepath or self.filepath)
        filepath.parent.mkdir(exist_ok=True, parents=True)
        out = {'n_episodes': self._curr_episode, 'env_params': self.env.params, 'header': self.env.summarize_header(),
               'episodes': self._recorder_out_list}
        filepath.write_text(json.dumps(out, default=lambda o: o.item() if hasattr(o, 'item') else str(o)))
        return out
