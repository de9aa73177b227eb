"""Level `.txt` parsing (host side only).

Keeps the reference's file format (marl_factory_grid/utils/level_parser.py:26-60 and
utils/helpers.py:168-202): one text row per grid row, `#` is a wall, every other character is
floor, `D` additionally marks a door tile.  Coordinates are (x, y) = (row, column), enumerated
row-major exactly like `np.argwhere` does in the reference - that order defines door indices
and the wall / door uids used by the `faithful` parity mode.
"""
from __future__ import annotations

from pathlib import Path
from typing import Union

import numpy as np

SYMBOL_WALL = '#'
SYMBOL_DOOR = 'D'

LEVELS_DIR = Path(__file__).parent / 'levels'


class LevelParser:
    def __init__(self, level_file_path: Union[str, Path]):
        path = Path(level_file_path)
        if not path.exists():
            raise FileNotFoundError(f'Level file not found: {path}')
        rows = [list(line.strip()) for line in path.read_text().splitlines()]
        rows = [r for r in rows if r] if any(rows) else rows
        if not rows:
            raise ValueError(f'Level file {path} is empty.')
        if len({len(r) for r in rows}) > 1:
            # helpers.py:181-182
            raise AssertionError('Every row of the level string must be of equal length.')
        self.path = path
        self.grid = np.array(rows)
        self.level_shape = self.grid.shape
        if self.level_shape[0] > 255 or self.level_shape[1] > 255:
            raise ValueError(f'Level {path.name} is {self.level_shape}; at most 255 x 255 tiles are supported.')
        self.walls = self.grid == SYMBOL_WALL
        self.floor = np.argwhere(~self.walls).astype(np.int32)
        self.wall_pos = np.argwhere(self.walls).astype(np.int32)
        self.door_pos = np.argwhere(self.grid == SYMBOL_DOOR).astype(np.int32)

    def get_coordinates_for_symbol(self, symbol, negate=False) -> np.ndarray:
        """Same contract as the reference helper (level_parser.py:46-60)."""
        hit = self.grid == str(symbol)
        return np.argwhere(~hit if negate else hit)


def resolve_level_path(level_name: str, custom_level_path=None) -> Path:
    """factory.py:97-100: a custom path wins, otherwise `<package>/levels/<level_name>.txt`."""
    if custom_level_path is not None:
        return Path(custom_level_path)
    return LEVELS_DIR / f'{level_name}.txt'
