#!/bin/bash
# ncu --set full captures of the Zillow launches (mask kernel staged / unstaged, dense launch); reports -> gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PROBE_CYCLES=${PROBE_CYCLES:-400}
ncu --set full --clock-control none --import-source on -k regex:stage_mask -s 3 -c 1 -f -o gpurun_out/r02_mask_staged python tools/kernel_probe.py > gpurun_out/ncu1.log 2>&1
TPLX_MASK_STAGE=0 ncu --set full --clock-control none --import-source on -k regex:stage_mask -s 3 -c 1 -f -o gpurun_out/r02_mask_unstaged python tools/kernel_probe.py > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stage_rows_kernel -s 3 -c 1 -f -o gpurun_out/r02_dense python tools/kernel_probe.py > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
