#!/bin/bash
cd "$(dirname "$0")/.."
for w in 4 2; do echo "== tile phases, wide $w"; TPLX_JIT_WIDE=$w timeout 200 python tools/tile_times.py 2>&1 | tail -9; done
echo "== wide 2 minb 6"; TPLX_JIT_WIDE=2 TPLX_JIT_MINB_WIDE=6 timeout 120 python tools/c1_probe.py 2>&1 | tail -1
