"""Visibility ray tables (host-side precompute for the kernels).

The reference builds, per agent and per step, 100 angular rays of length `min(obs_shape)` (the window
DIAMETER, not the radius - observation_builder.py:244), rounds and de-duplicates the targets
(ray_caster.py:34-49) and rasterises each with an integer Bresenham walk (ray_caster.py:143-199).
The result only depends on the radius and is translation invariant, so it is computed once here:

  full_ray_table(D)     every ray to its target, origin cell first, in the reference's target order
                        (needed for the `faithful` first-visit de-duplication)
  window_ray_table(r)   the distinct ray prefixes that stay inside the (2r+1)^2 window (28 rays of 3
                        cells for r = 3) - what the tiled observation kernel marches
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Tuple

import numpy as np

Cell = Tuple[int, int]


def _targets(radius: int, n_rays: int = 100, degs: int = 360) -> List[Cell]:
    # rotate the vector (0, -radius) by n_rays angles from +180 down to -180 degrees; the float rounding of
    # cos/sin is part of the behaviour (the target set is not symmetric), so the same numpy ops are used
    angles = np.deg2rad(np.linspace(-degs // 2, degs // 2, n_rays)[::-1])
    vec = np.array([0, -1]) * radius
    pts = np.stack([np.array([[math.cos(t), -math.sin(t)], [math.sin(t), math.cos(t)]]) @ vec for t in angles], 0)
    uniq = np.unique(np.round(pts), axis=0).astype(int)       # lexicographic order == reference ray order
    return [(int(a), int(b)) for a, b in uniq]


def _line(tx: int, ty: int) -> List[Cell]:
    """Integer Bresenham from (0, 0) to (tx, ty) with the reference's tie-breaking."""
    x1 = y1 = 0
    x2, y2 = tx, ty
    steep = abs(y2 - y1) > abs(x2 - x1)
    if steep:
        x1, y1, x2, y2 = y1, x1, y2, x2
    flipped = x1 > x2
    if flipped:
        x1, x2, y1, y2 = x2, x1, y2, y1
    dx, dy = x2 - x1, abs(y2 - y1)
    err = int(dx / 2.0)
    sy = 1 if y1 < y2 else -1
    cells, y = [], y1
    for x in range(x1, x2 + 1):
        cells.append((y, x) if steep else (x, y))
        err -= dy
        if err < 0:
            y += sy
            err += dx
    return cells[::-1] if flipped else cells


@lru_cache(maxsize=None)
def full_ray_table(diameter: int) -> Tuple[Tuple[Cell, ...], ...]:
    return tuple(tuple(_line(tx, ty)) for tx, ty in _targets(diameter))


@lru_cache(maxsize=None)
def window_ray_table(pomdp_r: int) -> Tuple[Tuple[Cell, ...], ...]:
    """Distinct in-window ray prefixes WITHOUT the origin cell, first-seen order."""
    seen, out = set(), []
    for ray in full_ray_table(2 * pomdp_r + 1):
        pref = tuple(c for c in ray[1:] if abs(c[0]) <= pomdp_r and abs(c[1]) <= pomdp_r)
        # the in-window part of a ray is always a prefix (cells move monotonically away from the origin)
        assert pref == tuple(ray[1:1 + len(pref)])
        if pref and pref not in seen:
            seen.add(pref)
            out.append(pref)
    # drop prefixes that are themselves a prefix of a longer kept ray (they add no visibility)
    keep = [p for p in out if not any(q != p and q[:len(p)] == p for q in out)]
    return tuple(keep)
