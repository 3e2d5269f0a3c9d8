#!/bin/bash
# Side-by-side builds of the experimental compile-time variants next to the product libraries:
#   libartis_b200_<preset>_prepass.so  -DARTISB200_CHI_PREPASS=1
#   libartis_b200_<preset>_masum.so    -DARTISB200_MA_SUMMARY=1
#   libartis_b200_<preset>_both.so     both
#   libartis_b200_<preset>_celllanes.so  -DARTISB200_BUILD_CELL_LANES=1 (macro-atom table builder: lanes over cells)
# usage (development container, before gpurun): bash tools/build_variants.sh [presets...]   (default: kilonova_lte classic)
# then on the GPU box:
#   for s in _prepass _masum _both _celllanes; do ARTISB200_LIB_SUFFIX=$s python -m pytest tests/test_gpu_parity.py -q -k "packet_histories and (kilonova_toy or classic3d_toy)"; done
#   TUNE=":;_prepass:;_masum:;_both:;_celllanes:" bash tools/gpu_round.sh r2 tune
set -eu
cd "$(dirname "$0")/.."
export ARTISB200_BUILD_PRESETS="${*:-kilonova_lte classic}"
ARTISB200_LIB_SUFFIX=_prepass ARTISB200_NVCC_EXTRA="-DARTISB200_CHI_PREPASS=1" python -c "import __graft_entry__ as g; g.build_cuda()"
ARTISB200_LIB_SUFFIX=_masum ARTISB200_NVCC_EXTRA="-DARTISB200_MA_SUMMARY=1" python -c "import __graft_entry__ as g; g.build_cuda()"
ARTISB200_LIB_SUFFIX=_both ARTISB200_NVCC_EXTRA="-DARTISB200_CHI_PREPASS=1 -DARTISB200_MA_SUMMARY=1" python -c "import __graft_entry__ as g; g.build_cuda()"
ARTISB200_LIB_SUFFIX=_celllanes ARTISB200_NVCC_EXTRA="-DARTISB200_BUILD_CELL_LANES=1" python -c "import __graft_entry__ as g; g.build_cuda()"
ls -la artis_b200/_build | grep -E "_prepass|_masum|_both|_celllanes"
