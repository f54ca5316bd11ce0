#!/bin/bash
# specialiser: parity tests, then kernel-time probes with the specialiser on / off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_jit.py -x -q 2>&1 | tail -15 > $O/jit_pytest.log; cat $O/jit_pytest.log
export TPLX_JIT_DUMP=$O
for j in 0 1; do
  echo "== C1 TPLX_JIT=$j"; TPLX_JIT=$j timeout 120 python tools/c1_probe.py 2>&1 | tail -3
done
for mb in 3 5 6; do echo "== C1 JIT minb $mb"; TPLX_JIT_MINB_VEC=$mb timeout 120 python tools/c1_probe.py 2>&1 | tail -1; done
for j in 0 1; do
  echo "== zillow TPLX_JIT=$j"; TPLX_JIT=$j timeout 120 python tools/kernel_probe.py 2>&1 | tail -8
done
for mb in 1 3 4; do echo "== zillow JIT minb $mb"; TPLX_JIT_MINB=$mb timeout 120 python tools/kernel_probe.py 2>&1 | tail -3; done
echo "== parity suite with every stage specialised"
TPLX_JIT=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q 2>&1 | tail -8
