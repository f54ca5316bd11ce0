#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_jit.py -x -q 2>&1 | tail -5
for w in 4 2 0; do echo "== C1 JIT wide $w"; TPLX_JIT_WIDE=$w timeout 120 python tools/c1_probe.py 2>&1 | tail -1; done
for mb in 2 4; do echo "== C1 JIT wide 4 minb $mb"; TPLX_JIT_MINB_WIDE=$mb timeout 120 python tools/c1_probe.py 2>&1 | tail -1; done
echo "== C1 JIT wide 2 minb 5"; TPLX_JIT_WIDE=2 TPLX_JIT_MINB_WIDE=5 timeout 120 python tools/c1_probe.py 2>&1 | tail -1
