"""ctypes binding of the C-ABI library (include/mfg_b200.h) + torch-owned device buffers.

PyTorch is plumbing here: it owns the device memory (one uint8 state tensor, the observation / reward /
done tensors) and the CUDA stream; every computation is done by the hand-written sm_100a kernels in
`libmfg_b200.so`.  There is no CPU or eager-PyTorch path: if the library is missing or no CUDA device is
visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Dict, Optional

import numpy as np

from .abi import (ENV_BLOCK, FIELD_DTYPES, FIELD_VIEW, N_STATS, RESPAWN_TAPE_W, STATE_FIELD_NAMES, MfgField, MfgTape,
                  PackedSpec)
from .spec import EnvSpec
from .state_io import columns_to_snapshot, snapshot_to_columns

# MFG_B200_LIB: development aid (A/B runs of differently compiled builds); the product loads the in-tree library
LIB_PATH = Path(os.environ.get('MFG_B200_LIB') or Path(__file__).resolve().parent / 'libmfg_b200.so')
_lib = None

EXPORTS = ['mfg_create', 'mfg_destroy', 'mfg_last_error', 'mfg_version', 'mfg_state_bytes', 'mfg_state_field',
           'mfg_bind_state', 'mfg_reset', 'mfg_step', 'mfg_observe', 'mfg_step_observe', 'mfg_random_actions',
           'mfg_step_host', 'mfg_stats', 'mfg_set_option', 'mfg_get_info', 'mfg_bind_step_flags']


def load_library(path: Path = LIB_PATH):
    """Load libmfg_b200.so and declare the prototypes of include/mfg_b200.h.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not path.exists():
        raise RuntimeError(f'{path} is missing: build the CUDA extension first '
                           f'(python -c "import __graft_entry__ as g; g.build()").  There is no CPU fallback.')
    L = C.CDLL(str(path))
    vp, i64, u64, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
    L.mfg_create.argtypes = [vp, i64, i64, C.POINTER(vp)]
    L.mfg_destroy.argtypes = [vp]
    L.mfg_destroy.restype = None
    L.mfg_last_error.restype = C.c_char_p
    L.mfg_version.restype = C.c_char_p
    L.mfg_state_bytes.argtypes = [vp]
    L.mfg_state_bytes.restype = C.c_size_t
    L.mfg_state_field.argtypes = [vp, C.c_char_p, C.POINTER(MfgField)]
    L.mfg_bind_state.argtypes = [vp, vp]
    L.mfg_reset.argtypes = [vp, vp, vp]
    L.mfg_step.argtypes = [vp, vp, C.POINTER(MfgTape), vp, vp, i32, vp]
    L.mfg_observe.argtypes = [vp, vp, vp]
    L.mfg_step_observe.argtypes = [vp, vp, C.POINTER(MfgTape), vp, vp, vp, i32, vp]
    L.mfg_random_actions.argtypes = [vp, vp, u64, u64, vp]
    L.mfg_step_host.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    L.mfg_stats.argtypes = [vp, vp, i32, vp]
    L.mfg_bind_step_flags.argtypes = [vp, vp]
    L.mfg_set_option.argtypes = [vp, C.c_char_p, i64]
    L.mfg_get_info.argtypes = [vp, C.c_char_p]
    L.mfg_get_info.restype = i64
    _lib = L
    return L


class EngineError(RuntimeError):
    pass


_TORCH_DTYPES = None


def _torch_dtype(np_dtype):
    import torch
    global _TORCH_DTYPES
    if _TORCH_DTYPES is None:
        # unsigned fields are held as the signed torch dtype of the same width (bit patterns are what matter;
        # numpy re-views them as unsigned on the way out)
        _TORCH_DTYPES = {np.uint8: torch.uint8, np.uint16: torch.int16, np.uint32: torch.int32,
                         np.uint64: torch.int64, np.float64: torch.float64, np.int16: torch.int16}
    return _TORCH_DTYPES[np_dtype]


class Engine:
    """N independent environments of one EnvSpec on one CUDA device."""

    def __init__(self, es: EnvSpec, n_envs: int, device='cuda', faithful: bool = True, seed: Optional[int] = None,
                 env_id_offset: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise EngineError('No CUDA device visible: the marl_factory_grid_b200 engine runs on the GPU only.')
        self.torch = torch
        self.lib = load_library()
        self.es = es
        self.N = int(n_envs)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise EngineError(f'device must be a CUDA device, got {device}')
        self.faithful = bool(faithful)
        self.packed = PackedSpec(es, faithful=faithful, seed=seed)
        self.h = C.c_void_p()
        with torch.cuda.device(self.device):
            self._check(self.lib.mfg_create(self.packed.ptr, self.N, int(env_id_offset), C.byref(self.h)))
            nbytes = self.lib.mfg_state_bytes(self.h)
            self.state = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            self._check(self.lib.mfg_bind_state(self.h, self.state.data_ptr()))
        self.fields: Dict[str, 'torch.Tensor'] = {}       # name -> strided VIEW [n_blocks, rows, 128] of the state buffer
        for name in STATE_FIELD_NAMES:
            f = MfgField()
            self._check(self.lib.mfg_state_field(self.h, name.encode(), C.byref(f)))
            if f.rows == 0:
                continue
            # blocked layout: element (row r, env e) = offset + (e // 128) * block_bytes + (r * 128 + e % 128) * elem_size
            dt = _torch_dtype(FIELD_VIEW.get(name, FIELD_DTYPES[f.elem_size]))
            n_blocks = (self.N + ENV_BLOCK - 1) // ENV_BLOCK
            self.fields[name] = torch.as_strided(self.state.view(dt), (n_blocks, f.rows, ENV_BLOCK),
                                                 (f.block_bytes // f.elem_size, ENV_BLOCK, 1), f.offset // f.elem_size)
        A = es.n_agents
        self.n_rew = A
        self.obs = torch.zeros((self.N, es.total_channels) + tuple(es.obs_shape), dtype=torch.float32, device=self.device)
        self.reward = torch.zeros((self.N, self.n_rew), dtype=torch.float32, device=self.device)
        self.done = torch.zeros(self.N, dtype=torch.uint8, device=self.device)
        self._stats = torch.zeros(N_STATS, dtype=torch.int64, device=self.device)
        self._tape_keep = None

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            raise EngineError(f'mfg error {rc}: {self.lib.mfg_last_error().decode()}')

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if getattr(self, 'h', None) and self.h.value:
            self.torch.cuda.synchronize(self.device)
            self.lib.mfg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def enable_step_flags(self, on: bool = True):
        """Per-step result flags [N, A + 1] uint8 (MFG_FLAG_* per agent, then the env's done reason): the batched form of the
        reference's `info` dict.  Written by every following step."""
        self.flags = self.torch.zeros((self.N, self.es.n_agents + 1), dtype=self.torch.uint8, device=self.device) if on else None
        self._check(self.lib.mfg_bind_step_flags(self.h, self.flags.data_ptr() if on else None))
        return self.flags

    def set_option(self, name: str, value: int):
        self._check(self.lib.mfg_set_option(self.h, name.encode(), int(value)))

    def info(self, name: str) -> int:
        return int(self.lib.mfg_get_info(self.h, name.encode()))

    # ------------------------------------------------------------------ hot path
    def reset(self, mask=None):
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=self.torch.uint8).contiguous()
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mfg_reset(self.h, m.data_ptr() if m is not None else None, self._stream()))

    def _tape(self, tape):
        if tape is None:
            return None
        t = self.torch
        keep = []
        st = MfgTape()
        if tape.get('maint_action') is not None and self.es.n_maint:
            ma = t.as_tensor(np.ascontiguousarray(tape['maint_action'], np.uint8)).reshape(self.N, self.es.n_maint).to(self.device)
            keep.append(ma)
            st.d_maint_action = ma.data_ptr()
        if tape.get('respawn_n') is not None:
            rn = t.as_tensor(np.ascontiguousarray(tape['respawn_n'], np.int8)).reshape(self.N).to(self.device)
            rp = t.as_tensor(np.ascontiguousarray(tape['respawn_pos'], np.uint16).view(np.int16)).reshape(self.N, RESPAWN_TAPE_W).to(self.device)
            keep += [rn, rp]
            st.d_respawn_n, st.d_respawn_pos = rn.data_ptr(), rp.data_ptr()
        self._tape_keep = keep + [st]
        return C.byref(st)

    def _actions(self, actions):
        t = self.torch
        a = t.as_tensor(actions)
        a = a.to(device=self.device, dtype=t.int32).reshape(self.N, self.es.n_agents).contiguous()
        return a

    def step(self, actions, tape=None, auto_reset=False):
        a = self._actions(actions)
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mfg_step(self.h, a.data_ptr(), self._tape(tape), self.reward.data_ptr(),
                                          self.done.data_ptr(), int(auto_reset), self._stream()))
        return self.reward, self.done

    def observe(self):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mfg_observe(self.h, self.obs.data_ptr(), self._stream()))
        return self.obs

    def step_observe(self, actions, tape=None, auto_reset=False):
        a = self._actions(actions)
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mfg_step_observe(self.h, a.data_ptr(), self._tape(tape), self.reward.data_ptr(),
                                                  self.done.data_ptr(), self.obs.data_ptr(), int(auto_reset),
                                                  self._stream()))
        return self.obs, self.reward, self.done

    def capture_step(self, actions, auto_reset=True):
        """CUDA-graph form of `step_observe` for launch-bound batch sizes: the whole call (k_step, packed re-spawn on the
        side stream, tiled + list-mode observation kernels, their memsets and the fork / join events) becomes ONE graph
        launch.  `actions` must be the int32 [N, A] device tensor the caller refills before every replay; results land
        in `self.obs / self.reward / self.done`.  Returns the `torch.cuda.CUDAGraph` (call `.replay()`).

        One eager step runs first so that the library's lazy allocations happen outside the capture; the state buffer
        is restored afterwards (the statistics vector keeps that step)."""
        t = self.torch
        a = self._actions(actions)
        if a.data_ptr() != actions.data_ptr():
            raise ValueError('capture_step needs a contiguous int32 [N, A] tensor on the engine device')
        saved = self.state.clone()
        side = t.cuda.Stream(self.device)
        side.wait_stream(t.cuda.current_stream(self.device))
        with t.cuda.stream(side):
            self.step_observe(a, auto_reset=auto_reset)
        t.cuda.current_stream(self.device).wait_stream(side)
        t.cuda.synchronize(self.device)
        self.state.copy_(saved)
        graph = t.cuda.CUDAGraph()
        with t.cuda.graph(graph, stream=side, capture_error_mode='thread_local'):
            self.step_observe(a, auto_reset=auto_reset)
        self._graph_keep = (a, side)
        return graph

    def random_actions(self, out, seed: int, step_index: int):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mfg_random_actions(self.h, out.data_ptr(), int(seed), int(step_index), self._stream()))
        return out

    def step_host(self, h_actions, h_reward, h_done, h_obs, auto_reset=False):
        """Host-buffer path: (pinned) CPU tensors in and out, copies inside the call."""
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mfg_step_host(self.h, h_actions.data_ptr(), h_reward.data_ptr(), h_done.data_ptr(),
                                               h_obs.data_ptr() if h_obs is not None else None, int(auto_reset),
                                               self._stream()))

    def stats(self, zero_after=False) -> np.ndarray:
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mfg_stats(self.h, self._stats.data_ptr(), int(zero_after), self._stream()))
        return self._stats.cpu().numpy()

    # ------------------------------------------------------------------ snapshots (tests, replay, checkpoints)
    def load_snapshot(self, env: int, snap: dict):
        t = self.torch
        for name, col in snapshot_to_columns(self.es, snap).items():
            src = col.view(np.int16) if col.dtype == np.uint16 else col.view(np.int32) if col.dtype == np.uint32 \
                else col.view(np.int64) if col.dtype == np.uint64 else col
            self.fields[name][env // ENV_BLOCK, :, env % ENV_BLOCK] = t.as_tensor(src).to(self.device)

    def field(self, name: str):
        """Copy of one state field as a [rows, N] tensor (signed dtype of the field's width; same bit patterns)."""
        v = self.fields[name]
        return v.permute(1, 0, 2).reshape(v.shape[1], -1)[:, :self.N]

    def fields_numpy(self) -> Dict[str, np.ndarray]:
        out = {}
        for name in self.fields:
            arr = self.field(name).cpu().numpy()
            np_dt = FIELD_VIEW.get(name)
            if np_dt is None:
                np_dt = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[arr.dtype.itemsize]
            out[name] = arr.view(np_dt)
        return out

    def snapshot(self, env: int, fields: Optional[Dict[str, np.ndarray]] = None) -> dict:
        f = fields if fields is not None else self.fields_numpy()
        return columns_to_snapshot(self.es, {k: v[:, env] for k, v in f.items()})
