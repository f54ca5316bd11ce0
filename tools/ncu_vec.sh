#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PROBE_ROWS=${PROBE_ROWS:-50000000} ncu --set full --clock-control none --import-source on -k regex:stage_rows_vec -s 2 -c 1 -f -o gpurun_out/r02c_vec python tools/c1_probe.py > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/r02c_vec.ncu-rep
