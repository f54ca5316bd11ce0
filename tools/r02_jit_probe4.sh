#!/bin/bash
cd "$(dirname "$0")/.."
for mb in 8 6 4; do for j in 4 2; do echo "== C1 JIT J=$j minb $mb"; TPLX_JIT_VEC_J=$j TPLX_JIT_MINB_VEC=$mb timeout 120 python tools/c1_probe.py 2>&1 | tail -1; done; done
echo "== zillow default"; timeout 120 python tools/kernel_probe.py 2>&1 | tail -3
