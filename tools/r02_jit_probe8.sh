#!/bin/bash
cd "$(dirname "$0")/.."
echo "== tile phases K1r B=4 minb 5"; TPLX_JIT_RE=4 timeout 200 python tools/tile_times.py 2>&1 | tail -9
for mb in 8 6 5; do echo "== C1 K1r B=2 minb $mb"; TPLX_JIT_RE=2 TPLX_JIT_MINB_RE=$mb timeout 120 python tools/c1_probe.py 2>&1 | tail -1; done
echo "== ncu K1r B=4"; PROBE_ROWS=50000000 timeout 300 ncu --set full --clock-control none -k regex:tplx_jit_kernel -s 2 -c 1 -f -o gpurun_out/r02_k1r python tools/c1_probe.py > gpurun_out/ncu_k1r.log 2>&1; tail -1 gpurun_out/ncu_k1r.log
