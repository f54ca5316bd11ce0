#!/bin/bash
# round-2b probe: dense launch occupancy vs L1 carve-out
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" TPLX_TRACE=1 python tools/kernel_probe.py 2>&1 | tail -2; }
{
run TPLX_DENSE_PAD=1100
run TPLX_DENSE_PAD=1100 TPLX_DENSE_CARVEOUT=85
run TPLX_DENSE_CARVEOUT=85
run TPLX_DENSE_CARVEOUT=71
run TPLX_DENSE_CARVEOUT=57
run TPLX_DENSE_CARVEOUT=100
run TPLX_DENSE_PAD=0
} 2>&1 | tee gpurun_out/probe_r2b.log
