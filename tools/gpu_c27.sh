#!/bin/bash
mkdir -p gpurun_out; T=${1:-c27}
timeout 300 python tools/kernel_bench.py --only mem > gpurun_out/${T}_ln_new.txt 2>&1; grep -i "layernorm" gpurun_out/${T}_ln_new.txt
WM_B200_LIB=tools/_diag/libwm_b200_lnold.so timeout 300 python tools/kernel_bench.py --only mem > gpurun_out/${T}_ln_old.txt 2>&1; grep -i "layernorm" gpurun_out/${T}_ln_old.txt
timeout 300 python tools/kernel_bench.py --only mem > gpurun_out/${T}_ln_new2.txt 2>&1; grep -i "layernorm" gpurun_out/${T}_ln_new2.txt
