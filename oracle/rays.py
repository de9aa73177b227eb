"""Restatement of the reference ray geometry (test infrastructure, see oracle/__init__.py).

Follows marl_factory_grid/utils/ray_caster.py:
  * build_ray_targets  :34-49   100 angles, rotate `north * radius`, round, np.unique  -> targets
  * bresenham_loop     :143-199 integer Bresenham with steep-axis swap and start/end swap
The radius handed to RayCaster is the window DIAMETER (`min(obs_shape)`, observation_builder.py:244),
i.e. 7 for pomdp_r = 3 (SURVEY.md defect B14).
"""
import math
from functools import lru_cache

import numpy as np


@lru_cache(maxsize=None)
def ray_targets(radius: int, n_rays: int = 100, degs: int = 360):
    north = np.array([0, -1]) * radius
    thetas = [np.deg2rad(deg) for deg in np.linspace(-degs // 2, degs // 2, n_rays)[::-1]]
    rot = np.stack([[[math.cos(t), -math.sin(t)], [math.sin(t), math.cos(t)]] for t in thetas], 0)
    rot = np.unique(np.round(rot @ north), axis=0)
    return tuple((int(a), int(b)) for a, b in rot.astype(int))


def bresenham(x1, y1, x2, y2):
    dx, dy = x2 - x1, y2 - y1
    steep = abs(dy) > abs(dx)
    if steep:
        x1, y1, x2, y2 = y1, x1, y2, x2
    swapped = False
    if x1 > x2:
        x1, x2, y1, y2 = x2, x1, y2, y1
        swapped = True
    dx, dy = x2 - x1, y2 - y1
    error = int(dx / 2.0)
    ystep = 1 if y1 < y2 else -1
    y = y1
    pts = []
    for x in range(int(x1), int(x2) + 1):
        pts.append((y, x) if steep else (x, y))
        error -= abs(dy)
        if error < 0:
            y += ystep
            error += dx
    if swapped:
        pts.reverse()
    return pts


@lru_cache(maxsize=None)
def full_rays(radius: int):
    """All rays as offsets from the origin, in `ray_targets` order, origin cell first."""
    return tuple(tuple(bresenham(0, 0, tx, ty)) for tx, ty in ray_targets(radius))
