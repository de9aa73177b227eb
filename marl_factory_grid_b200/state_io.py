"""Conversion between the engine's struct-of-arrays state fields and per-environment snapshots.

A snapshot is a dict of small numpy arrays with the same keys the golden traces use
(tests/golden/make_golden.py::snapshot): agent_pos [A,2], door_open [ND], door_timer [ND], dirt_pos
[K,2] / dirt_amt / dirt_uid / dirt_n, item_pos ... and the `*_listed` bits of the uid listing model.
`Factory.load_state` / `Factory.export_state` use these helpers; the parity tests use them to start
an engine from a reference spawn table and to compare the state after every step.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .abi import NO_POS, pos16
from .spec import EnvSpec

GONE = -9999
_SMALL = (('item', 'n_items'), ('pod', 'n_pods'), ('dest', 'n_dest'), ('drop', 'n_dropoff'), ('machine', 'n_machines'),
          ('maint', 'n_maint'))
_FIELD_OF = {'item': 'item', 'pod': 'pod', 'dest': 'dest', 'drop': 'drop', 'machine': 'mach', 'maint': 'maint'}


def _mask(bits) -> int:
    m = 0
    for k, b in enumerate(bits):
        if b:
            m |= 1 << k
    return m


def snapshot_to_columns(es: EnvSpec, snap: dict) -> Dict[str, np.ndarray]:
    """One env: field name -> 1-D array of `rows` values (same dtypes as the state fields)."""
    A = es.n_agents
    col: Dict[str, np.ndarray] = {}
    col['step'] = np.array([int(snap.get('step', 0))], np.uint16)
    col['episode'] = np.array([0], np.uint32)
    apos = np.asarray(snap['agent_pos']).reshape(A, 2)
    col['apos'] = np.array([pos16(p) for p in apos], np.uint16)
    rank = np.asarray(snap['agent_rank']).reshape(A) if 'agent_rank' in snap else np.zeros(A, np.int64)
    col['astamp'] = np.array([int(rank[i]) * A + i for i in range(A)], np.uint32)
    col['clock'] = np.array([int(col['astamp'].max()) + 1], np.uint32)
    par = np.asarray(snap['paralysed']).reshape(A) if 'paralysed' in snap else np.zeros(A)
    col['aflag'] = par.astype(np.uint8)
    col['finished'] = np.array([0], np.uint8)
    col['ep_ret'] = np.zeros(A, np.float64)
    if es.has_batteries:
        col['bat'] = np.asarray(snap['battery'], np.float64).reshape(A)
    if es.n_doors:
        col['door_open'] = np.array([_mask(snap['door_open'])], np.uint64)
        listed = snap['door_listed'] if 'door_listed' in snap else np.ones(es.n_doors)
        col['door_listed'] = np.array([_mask(listed)], np.uint64)
        col['door_timer'] = np.asarray(snap['door_timer']).astype(np.uint8)
    if es.has_dirt:
        K = es.dirt_slots
        n = int(snap['dirt_n'])
        if n > K:
            raise ValueError(f'snapshot holds {n} dirt piles but the engine has {K} dirt slots')
        dpos = np.full(K, NO_POS, np.uint16)
        damt = np.zeros(K, np.float64)
        duid = np.zeros(K, np.uint16)
        for k in range(n):
            dpos[k] = pos16(snap['dirt_pos'][k])
            damt[k] = snap['dirt_amt'][k]
            duid[k] = snap['dirt_uid'][k] if 'dirt_uid' in snap else k
        # the uid model of a fresh episode (object.py:103-113 counters start at 0): strictly increasing in creation order, so
        # slot k holds a uid >= k; the step kernel's uid listing relies on it (Env::find_listed)
        if n and (np.any(np.diff(duid[:n].astype(np.int64)) <= 0) or np.any(duid[:n] < np.arange(n))):
            raise ValueError('snapshot: dirt uids must increase strictly in creation order (fresh-episode uid model)')
        listed = snap['dirt_listed'][:n] if 'dirt_listed' in snap else np.ones(n)
        col.update(dirt_pos=dpos, dirt_amt=damt, dirt_uid=duid,
                   dirt_listed=np.array([_mask(listed)], np.uint64),
                   dirt_end=np.array([n], np.uint8), dirt_n=np.array([n], np.uint8),
                   dirt_next_uid=np.array([int(snap.get('dirt_next_uid', n))], np.uint16),
                   dirt_next_spawn=np.array([int(snap.get('dirt_next_spawn', -1))], np.int16))
    for key, attr in _SMALL:
        n = getattr(es, attr)
        if not n:
            continue
        f = _FIELD_OF[key]
        pos = np.asarray(snap[f'{key}_pos']).reshape(n, 2)
        col[f'{f}_pos'] = np.array([NO_POS if int(p[0]) < 0 else pos16(p) for p in pos], np.uint16)
        listed = snap[f'{key}_listed'] if f'{key}_listed' in snap else (pos[:, 0] >= 0)
        col[f'{f}_listed'] = np.array([_mask(listed)], np.uint32)
    if es.n_dest:
        col['dest_reached'] = np.array([_mask(snap['dest_reached'])], np.uint32)
    if es.n_maint:
        n = es.n_maint
        col['maint_target'] = np.full(n, NO_POS, np.uint16)
        col['maint_rand'] = np.full(n, NO_POS, np.uint16)
        col['maint_remaining'] = np.zeros(n, np.uint32)
        col['maint_last'] = np.full(n, 0xFF, np.uint8)
    return col


def _unpos(p: int):
    return (GONE, GONE) if p == NO_POS else (p >> 8, p & 255)


def columns_to_snapshot(es: EnvSpec, col: Dict[str, np.ndarray], dirt_pad: int = 64) -> dict:
    """Inverse of snapshot_to_columns; dirt piles are listed in creation order without tombstones."""
    A = es.n_agents
    s = {'agent_pos': np.array([_unpos(int(p)) for p in col['apos']], np.int16).reshape(A, 2),
         'step': np.int32(col['step'][0]),
         'paralysed': (np.asarray(col['aflag']) & 1).astype(np.uint8)}
    nd = es.n_doors
    if nd:
        o, l = int(col['door_open'][0]), int(col['door_listed'][0])
        s['door_open'] = np.array([(o >> d) & 1 for d in range(nd)], np.uint8)
        s['door_listed'] = np.array([(l >> d) & 1 for d in range(nd)], np.uint8)
        s['door_timer'] = np.asarray(col['door_timer']).astype(np.int16)
    else:
        s['door_open'] = s['door_listed'] = np.zeros(0, np.uint8)
        s['door_timer'] = np.zeros(0, np.int16)
    dp = np.full((dirt_pad, 2), GONE, np.int16)
    da, du, dl = np.zeros(dirt_pad), np.full(dirt_pad, -1, np.int32), np.zeros(dirt_pad, np.uint8)
    n = 0
    if es.has_dirt:
        listed = int(col['dirt_listed'][0])
        for k in range(int(col['dirt_end'][0])):
            p = int(col['dirt_pos'][k])
            if p == NO_POS:
                continue
            dp[n], da[n], du[n], dl[n] = _unpos(p), col['dirt_amt'][k], col['dirt_uid'][k], (listed >> k) & 1
            n += 1
        assert n == int(col['dirt_n'][0]), 'dirt_n out of sync with the slot table'
        s['dirt_next_uid'] = np.int32(col['dirt_next_uid'][0])
        s['dirt_next_spawn'] = np.int32(col['dirt_next_spawn'][0])
    else:
        s['dirt_next_uid'] = np.int32(0)
        s['dirt_next_spawn'] = np.int32(-1)
    s.update(dirt_n=np.int32(n), dirt_pos=dp, dirt_amt=da, dirt_uid=du, dirt_listed=dl)
    for key, attr in _SMALL:
        cnt = getattr(es, attr)
        f = _FIELD_OF[key]
        if cnt:
            l = int(col[f'{f}_listed'][0])
            s[f'{key}_pos'] = np.array([_unpos(int(p)) for p in col[f'{f}_pos']], np.int16).reshape(cnt, 2)
            s[f'{key}_listed'] = np.array([(l >> k) & 1 for k in range(cnt)], np.uint8)
        else:
            s[f'{key}_pos'] = np.zeros((0, 2), np.int16)
            s[f'{key}_listed'] = np.zeros(0, np.uint8)
    r = int(col['dest_reached'][0]) if es.n_dest else 0
    s['dest_reached'] = np.array([(r >> k) & 1 for k in range(es.n_dest)], np.uint8)
    s['battery'] = np.asarray(col['bat'], np.float64).copy() if es.has_batteries else np.zeros(0, np.float64)
    return s
