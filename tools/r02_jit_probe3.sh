#!/bin/bash
# K1v with one barrier per phase; mask kernel with software prefetch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jit.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -5
for j in 0 1; do echo "== C1 TPLX_JIT=$j"; TPLX_JIT=$j timeout 120 python tools/c1_probe.py 2>&1 | tail -1; done
echo "== C1 JIT minb 6"; TPLX_JIT_MINB_VEC=6 timeout 120 python tools/c1_probe.py 2>&1 | tail -1
echo "== zillow default"; timeout 120 python tools/kernel_probe.py 2>&1 | tail -3
echo "== zillow no mask prefetch"; TPLX_MASK_PREFETCH=0 timeout 120 python tools/kernel_probe.py 2>&1 | tail -3
echo "== zillow interpreter, mask prefetch"; TPLX_JIT=0 timeout 120 python tools/kernel_probe.py 2>&1 | tail -3
echo "== zillow mask minb 5"; TPLX_MASK_MINB=5 timeout 120 python tools/kernel_probe.py 2>&1 | tail -3
