#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_jit.py -x -q 2>&1 | tail -4
for re in 4 2; do for mb in 5 4 6; do echo "== C1 K1r defer B=$re minb $mb"; TPLX_JIT_RE=$re TPLX_JIT_MINB_RE=$mb timeout 120 python tools/c1_probe.py 2>&1 | tail -1; done; done
echo "== C1 K1r no defer B=4 minb 5"; TPLX_JIT_DEFER=0 TPLX_JIT_RE=4 TPLX_JIT_MINB_RE=5 timeout 120 python tools/c1_probe.py 2>&1 | tail -1
echo "== C1 K1r defer B=8 minb 4"; TPLX_JIT_RE=8 TPLX_JIT_MINB_RE=4 timeout 120 python tools/c1_probe.py 2>&1 | tail -1
