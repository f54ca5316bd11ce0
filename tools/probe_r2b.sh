#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== zillow default"; TPLX_TRACE=1 python tools/kernel_probe.py 2>&1 | tail -3
echo "== c1 default"; python tools/c1_probe.py
} 2>&1 | tee gpurun_out/probe_r2b.log
