#!/bin/bash
mkdir -p gpurun_out; T=${1:-lnab}
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "layernorm or ln" > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
for i in 1 2; do
timeout 300 python tools/kernel_bench.py --only mem 2>&1 | grep -i "layernorm_bwd" | cut -c1-150
WM_OPTIONS=ln_bwd_rows=14 timeout 300 python tools/kernel_bench.py --only mem 2>&1 | grep -i "encoder form" | sed 's/^/rows14 /' | cut -c1-150
WM_B200_LIB=tools/_diag/libwm_b200_lnold.so timeout 300 python tools/kernel_bench.py --only mem 2>&1 | grep -i "encoder form" | sed 's/^/old    /' | cut -c1-150
done
