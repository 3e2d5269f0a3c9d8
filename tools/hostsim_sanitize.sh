#!/bin/bash
# TEST INFRASTRUCTURE: compile the physics headers for the host (tests/hostsim) with AddressSanitizer + UBSan and replay
# oracle fixtures through every schedule. usage: [EXTRA_DEFS="-DARTISB200_HOSTSIM_FUZZ_LIBM ..."] bash tools/hostsim_sanitize.sh [preset config nts]
set -eu
cd "$(dirname "$0")/.."
PRESET=${1:-kilonova_lte}; CONFIG=${2:-kilonova_toy}; NTS=${3:-4}
OUT=/tmp/artis_b200_asan; mkdir -p $OUT
g++ -std=c++20 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -fPIC -shared -Wno-unknown-pragmas \
    -Wno-subobject-linkage -Iartis_b200/csrc ${EXTRA_DEFS:-} "-DARTISB200_PRESET_HEADER=\"options/preset_${PRESET}.h\"" tests/hostsim/hostsim.cc \
    -o $OUT/libhostsim_${PRESET}.so
cat > $OUT/run.py <<PY
import os, sys
sys.path.insert(0, "$(pwd)"); sys.path.insert(0, "$(pwd)/tools")
os.environ["ARTISB200_ALLOW_HOSTSIM"] = "1"
from tests import parity_checks
lib = "$OUT/libhostsim_${PRESET}.so"
for opts in ({"schedule": 1, "wf_tail": 0, "wf_resort_every": 1}, {"schedule": 0},
             {"schedule": 1, "wf_tail": 0, "wf_masteps": 1, "wf_ma_rounds": 3, "wf_masteps_last": 2}):
    parity_checks.check_packet_histories(lib, "${CONFIG}", ${NTS}, options=opts)
parity_checks.check_deterministic_kernels(lib, "${CONFIG}", ${NTS})
print("sanitized replay of ${CONFIG} ts${NTS} (${PRESET}): clean")
PY
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 python $OUT/run.py
