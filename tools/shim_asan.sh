#!/bin/bash
# The CPU-shim kernel tests (tests/test_kernels_on_cpu_shim.py, test_frontend_on_cpu_shim.py, test_thinking_oracle.py) under the address sanitizer: the kernel SOURCES of csrc/ run one OS
# thread per CUDA thread with ASan red zones around every torch CPU allocation, static __shared__ array and dynamic shared block.
# Usage: bash tools/shim_asan.sh [pytest -k expression]
cd "$(dirname "$0")/.."
export UA2_SHIM_ASAN=1 ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1
LD_PRELOAD="$(gcc -print-file-name=libasan.so)" python -m pytest tests/test_kernels_on_cpu_shim.py tests/test_frontend_on_cpu_shim.py tests/test_thinking_oracle.py -x -q ${1:+-k "$1"}
