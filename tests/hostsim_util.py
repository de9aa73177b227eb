"""ctypes wrapper around the TEST-ONLY host build of the device functions (tests/hostsim/hostsim.cpp).

Lets the CPU-only container replay reference traces through the exact per-environment code the CUDA
kernels execute.  Never used by the product.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from marl_factory_grid_b200.abi import (ENV_BLOCK, FIELD_DTYPES, FIELD_VIEW, N_STATS, RESPAWN_TAPE_W, STATE_FIELD_NAMES,
                                        MfgField, PackedSpec, pos16)
from marl_factory_grid_b200.state_io import columns_to_snapshot, snapshot_to_columns

HERE = Path(__file__).resolve().parent / 'hostsim'
LIB = HERE / '_build' / 'libhostsim.so'


def build_hostsim(force=False):
    srcs = [HERE / 'hostsim.cpp'] + list((HERE.parent.parent / 'marl_factory_grid_b200' / 'csrc').glob('*.cuh')) + \
        list((HERE.parent.parent / 'marl_factory_grid_b200' / 'csrc').glob('*.hpp')) + \
        [HERE.parent.parent / 'include' / 'mfg_b200.h']
    if force or not LIB.exists() or any(s.stat().st_mtime > LIB.stat().st_mtime for s in srcs):
        LIB.parent.mkdir(exist_ok=True)
        subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', str(LIB), str(HERE / 'hostsim.cpp')],
                       check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build_hostsim()))
        L.hs_create.restype = C.c_char_p
        L.hs_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
        L.hs_destroy.argtypes = [C.c_void_p]
        L.hs_state_bytes.restype = C.c_size_t
        L.hs_state_bytes.argtypes = [C.c_void_p]
        L.hs_state_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(MfgField)]
        L.hs_bind_state.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_step.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int]
        L.hs_observe.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_bind_flags.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class HostSim:
    def __init__(self, es, n_envs=1, faithful=True, seed=None, env_id_offset=0):
        self.es, self.N = es, n_envs
        self.packed = PackedSpec(es, faithful=faithful, seed=seed)
        self.h = C.c_void_p()
        err = lib().hs_create(self.packed.ptr, n_envs, env_id_offset, C.byref(self.h))
        if err:
            raise RuntimeError(err.decode())
        self.buf = np.zeros(lib().hs_state_bytes(self.h), np.uint8)
        lib().hs_bind_state(self.h, self.buf.ctypes.data)
        self._views = {}
        for name in STATE_FIELD_NAMES:
            f = MfgField()
            assert lib().hs_state_field(self.h, name.encode(), C.byref(f)) == 0, name
            if f.rows == 0:
                continue
            dt = FIELD_VIEW.get(name, FIELD_DTYPES[f.elem_size])
            n_blocks = (n_envs + ENV_BLOCK - 1) // ENV_BLOCK
            typed = self.buf[f.offset:].view(dt) if (len(self.buf) - f.offset) % f.elem_size == 0 else \
                self.buf[f.offset:f.offset + (len(self.buf) - f.offset) // f.elem_size * f.elem_size].view(dt)
            # blocked layout: strided VIEW [n_blocks, rows, 128]
            self._views[name] = np.lib.stride_tricks.as_strided(
                typed, shape=(n_blocks, f.rows, ENV_BLOCK), strides=(f.block_bytes, ENV_BLOCK * f.elem_size, f.elem_size))
        A = es.n_agents
        self.n_rew = A
        self.reward = np.zeros((n_envs, self.n_rew), np.float32)
        self.done = np.zeros(n_envs, np.uint8)
        self.obs = np.zeros((n_envs, es.total_channels) + tuple(es.obs_shape), np.float32)

    def __del__(self):
        if getattr(self, 'h', None):
            lib().hs_destroy(self.h)
            self.h = None

    @property
    def fields(self):
        """name -> [rows, N] COPY of the field (same convention as Engine.fields_numpy())."""
        return {k: v.transpose(1, 0, 2).reshape(v.shape[1], -1)[:, :self.N].copy() for k, v in self._views.items()}

    def load_snapshot(self, env, snap):
        for name, colv in snapshot_to_columns(self.es, snap).items():
            self._views[name][env // ENV_BLOCK, :, env % ENV_BLOCK] = colv

    def snapshot(self, env):
        return columns_to_snapshot(self.es, {k: v[env // ENV_BLOCK, :, env % ENV_BLOCK] for k, v in self._views.items()})

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().hs_reset(self.h, None if m is None else m.ctypes.data)

    def step(self, actions, maint_act=None, respawn=None, auto_reset=False):
        """actions [N, A]; maint_act [N, NM] or None; respawn = (n [N] int8, pos [N, 8] uint16) or None."""
        a = np.ascontiguousarray(actions, np.int32).reshape(self.N, self.es.n_agents)
        ma = None if maint_act is None else np.ascontiguousarray(maint_act, np.uint8).reshape(self.N, -1)
        rn = rp = None
        if respawn is not None:
            rn = np.ascontiguousarray(respawn[0], np.int8).reshape(self.N)
            rp = np.ascontiguousarray(respawn[1], np.uint16).reshape(self.N, RESPAWN_TAPE_W)
        lib().hs_step(self.h, a.ctypes.data, None if ma is None or ma.size == 0 else ma.ctypes.data,
                      None if rn is None else rn.ctypes.data, None if rp is None else rp.ctypes.data,
                      self.reward.ctypes.data, self.done.ctypes.data, int(auto_reset))
        return self.reward, self.done

    def enable_step_flags(self):
        self.flags = np.zeros((self.N, self.es.n_agents + 1), np.uint8)
        lib().hs_bind_flags(self.h, self.flags.ctypes.data)
        return self.flags

    def observe(self):
        lib().hs_observe(self.h, self.obs.ctypes.data)
        return self.obs

    def stats(self):
        out = np.zeros(N_STATS, np.int64)
        lib().hs_stats(self.h, out.ctypes.data)
        return out


def tape_respawn(ep, t):
    """Golden episode -> (n int8, pos16[8]) for step index t (0-based)."""
    n = int(ep['respawn_n'][t])
    pos = np.zeros(RESPAWN_TAPE_W, np.uint16)
    if n > 0:
        pos[:n] = [pos16(p) for p in ep['respawn_tiles'][t][:n]]
    return np.int8(max(n, 0)), pos
