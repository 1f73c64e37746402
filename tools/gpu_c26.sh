#!/bin/bash
mkdir -p gpurun_out; T=${1:-c26}
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python tools/kernel_bench.py --only mem > gpurun_out/${T}_ln_new.txt 2>&1; tail -6 gpurun_out/${T}_ln_new.txt
WM_B200_LIB=tools/_diag/libwm_b200_lnold.so timeout 300 python tools/kernel_bench.py --only mem > gpurun_out/${T}_ln_old.txt 2>&1; tail -6 gpurun_out/${T}_ln_old.txt
timeout 300 python tools/gemm_sites.py large > gpurun_out/${T}_sites.txt 2>&1; tail -12 gpurun_out/${T}_sites.txt
for v in "1 8 0" "1 16 0" "1 16 2"; do timeout 200 python tools/gemm_roofline_once.py $v >> gpurun_out/${T}_roof.txt 2>&1; done; cat gpurun_out/${T}_roof.txt
timeout 300 python bench.py --no-trainer > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 1200 gpurun_out/${T}_bench.json
