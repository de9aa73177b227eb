"""marl_factory_grid_b200 - B200-native batched stepping engine behind the marl-factory-grid surface.

Only the hot path of the reference is rebuilt here: batched `Factory.reset/step` including the
per-agent ray-cast observation build (marl_factory_grid/environment/factory.py:134-220), as
hand-written sm_100a CUDA kernels behind a C-ABI (include/mfg_b200.h).  The host side keeps the
reference's yaml schema, level `.txt` format and observation / reward layout.

`Factory` needs torch + a CUDA device; the config compiler (`FactoryConfigParser`, `EnvSpec`) is
pure Python and importable anywhere.
"""
from .config_parser import FactoryConfigParser, named_action_space  # noqa: F401
from .level_parser import LevelParser  # noqa: F401
from .spec import EnvSpec  # noqa: F401

__all__ = ['Factory', 'EnvMonitor', 'EnvRecorder', 'FactoryConfigParser', 'LevelParser', 'EnvSpec', 'named_action_space']


def __getattr__(name):
    if name == 'Factory':
        from .factory import Factory
        return Factory
    if name == 'EnvMonitor':
        from .monitor import EnvMonitor
        return EnvMonitor
    if name == 'EnvRecorder':
        from .monitor import EnvRecorder
        return EnvRecorder
    raise AttributeError(name)
