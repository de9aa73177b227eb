#!/bin/bash
# Installs the UNMODIFIED reference (illiumst/marl-factory-grid, /root/reference) into the git-ignored baseline/_ref/
# (offline pip install, no dependency resolution: numpy / numba / networkx / pyyaml / pandas come with the image).
# baseline/_ref travels to the GPU box with the gpurun snapshot; bench.py's CPU arms import it from there.
set -e
cd "$(dirname "$0")/.."
rm -rf baseline/_ref
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target baseline/_ref /root/reference
