#!/bin/bash
cd "$(dirname "$0")/.."
for mb in 6 8; do echo "== C1 K1r defer B=1 minb $mb"; TPLX_JIT_RE=1 TPLX_JIT_MINB_RE=$mb timeout 120 python tools/c1_probe.py 2>&1 | tail -1; done
echo "== C1 default"; timeout 120 python tools/c1_probe.py 2>&1 | tail -1
echo "== C1 K1r B=2 minb 8"; TPLX_JIT_MINB_RE=8 timeout 120 python tools/c1_probe.py 2>&1 | tail -1
