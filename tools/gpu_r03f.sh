#!/bin/bash
# GPU session r03f: fused-kernel sweep only
set -x
mkdir -p gpurun_out
timeout 600 python tools/ab_v3.py quick cfg2 > gpurun_out/r03f_ab_cfg2.log 2>&1
timeout 600 python tools/ab_v3.py quick cfg4 > gpurun_out/r03f_ab_cfg4.log 2>&1
