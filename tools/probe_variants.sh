#!/bin/bash
# kernel-time probe of the Zillow stage under the library's tuning knobs (one gpurun call, several variants)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== default"; python tools/kernel_probe.py
echo "== old prefilter (look-back kernel)"; TPLX_NO_MASK=1 python tools/kernel_probe.py
echo "== mask, no staging"; TPLX_MASK_STAGE=0 python tools/kernel_probe.py
echo "== mask MR=2"; TPLX_MASK_MR=2 python tools/kernel_probe.py
echo "== mask smem 100K"; TPLX_MASK_SMEM=102400 python tools/kernel_probe.py
echo "== mask slack 1.25"; TPLX_MASK_SLACK=1.25 python tools/kernel_probe.py
echo "== no scan hint (VM in the mask kernel)"; TPLX_NO_SCAN=1 python tools/kernel_probe.py
echo "== scan, no staging"; TPLX_MASK_STAGE=0 python tools/kernel_probe.py
echo "== scan, slack 1.25 MR=2"; TPLX_MASK_SLACK=1.25 TPLX_MASK_MR=2 python tools/kernel_probe.py
} 2>&1 | tee gpurun_out/probe_variants.log
