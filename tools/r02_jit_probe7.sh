#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_jit.py -x -q 2>&1 | tail -5
for re in 4 8; do for mb in 6 5 4 8; do echo "== C1 K1r B=$re minb $mb"; TPLX_JIT_RE=$re TPLX_JIT_MINB_RE=$mb timeout 120 python tools/c1_probe.py 2>&1 | tail -1; done; done
echo "== C1 K1w (RE off)"; TPLX_JIT_RE=0 timeout 120 python tools/c1_probe.py 2>&1 | tail -1
