#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PROBE_CYCLES=${PROBE_CYCLES:-400}
ncu --set full --clock-control none --import-source on -k regex:stage_mask -s 3 -c 1 -f -o gpurun_out/r02_scan_staged python tools/kernel_probe.py > gpurun_out/ncu4.log 2>&1
TPLX_MASK_STAGE=0 ncu --set full --clock-control none --import-source on -k regex:stage_mask -s 3 -c 1 -f -o gpurun_out/r02_scan_unstaged python tools/kernel_probe.py > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/r02_scan*.ncu-rep
