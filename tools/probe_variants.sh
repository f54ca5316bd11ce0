#!/bin/bash
# kernel-time probe of the Zillow stage under the library's tuning knobs (one gpurun call, several variants)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== default"; python tools/kernel_probe.py
echo "== old prefilter (look-back kernel)"; TPLX_NO_MASK=1 python tools/kernel_probe.py
echo "== mask, TMA staging"; TPLX_MASK_STAGE=1 python tools/kernel_probe.py
echo "== no scan hint (VM in the mask kernel)"; TPLX_NO_SCAN=1 python tools/kernel_probe.py
echo "== no inplace dense"; TPLX_NO_INPLACE=1 python tools/kernel_probe.py
} 2>&1 | tee gpurun_out/probe_variants.log
