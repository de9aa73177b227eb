import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) where they cannot run: no CUDA device, or the library is not built."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    lib = ROOT / 'marl_factory_grid_b200' / 'libmfg_b200.so'
    if has_cuda and lib.exists():
        return
    why = 'no CUDA device' if not has_cuda else 'libmfg_b200.so not built'
    skip = pytest.mark.skip(reason=f'gpu test: {why}')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
