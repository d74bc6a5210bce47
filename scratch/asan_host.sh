#!/bin/bash
# Host-side AddressSanitizer run of the C-ABI shim (compute-sanitizer is closed on the pool; this covers the HOST half: handle
# life cycle, argument validation, pinned / IPC buffers, the ctypes boundary).  The device code is unchanged.
#   bash scratch/asan_host.sh [pytest args]        (default: the whole -m gpu suite; on a box without a GPU: tests/test_abi.py)
set -e
cd "$(dirname "$0")/.."
export DRONECU_OUT=$PWD/drone_rl_b200/libdronecu_asan.so
[ -f "$DRONECU_OUT" ] || DRONECU_DEFINES="-Xcompiler -fsanitize=address -Xcompiler -fno-omit-frame-pointer" python -m drone_rl_b200.build --force
export DRONECU_LIB=$DRONECU_OUT
export LD_PRELOAD=$(gcc -print-file-name=libasan.so)
# protect_shadow_gap=0: the CUDA driver maps memory in ASan's shadow gap; leaks: python and the CUDA runtime never free at exit
export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=0:halt_on_error=1
if [ $# -eq 0 ]; then set -- tests -m gpu -x -q; fi
python -m pytest "$@"
