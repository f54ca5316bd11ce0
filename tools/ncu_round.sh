#!/bin/bash
# one ncu --set full capture each: K1f scan-mask kernel, dense launch, K1v vector kernel (C1); reports -> gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PROBE_CYCLES=${PROBE_CYCLES:-400}
ncu --set full --clock-control none --import-source on -k regex:stage_mask -s 3 -c 1 -f -o gpurun_out/r02b_scan python tools/kernel_probe.py > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stage_rows_kernel -s 3 -c 1 -f -o gpurun_out/r02b_dense python tools/kernel_probe.py > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stage_rows_vec -s 2 -c 1 -f -o gpurun_out/r02b_vec python tools/c1_probe.py > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/r02b_*.ncu-rep
