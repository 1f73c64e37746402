#!/bin/bash
mkdir -p gpurun_out; T=${1:-c24}
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
timeout 600 python tools/cublas_compare.py large > gpurun_out/${T}_cublas.txt 2>&1; tail -25 gpurun_out/${T}_cublas.txt
timeout 300 python tools/gemm_sites.py large > gpurun_out/${T}_sites.txt 2>&1; tail -3 gpurun_out/${T}_sites.txt
for v in "1 8 0" "1 16 0" "1 16 2"; do timeout 200 python tools/gemm_roofline_once.py $v >> gpurun_out/${T}_roof.txt 2>&1; done; cat gpurun_out/${T}_roof.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 4 -c 2 -f -o gpurun_out/${T}_roofline python tools/gemm_roofline_once.py 1 8 0 > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
