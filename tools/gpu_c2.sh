#!/bin/bash
mkdir -p gpurun_out; T=${1:-c2}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > gpurun_out/${T}_pytest_gemm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gemm.log; tail -8 gpurun_out/${T}_pytest_gemm.log
timeout 300 python tools/gemm_sites.py large > gpurun_out/${T}_gemm_sites.txt 2>&1; cat gpurun_out/${T}_gemm_sites.txt
timeout 300 python tools/cublas_compare.py large > gpurun_out/${T}_cublas.txt 2>&1; cat gpurun_out/${T}_cublas.txt
timeout 300 python tools/step_timeline.py > gpurun_out/${T}_timeline.txt 2>&1; cat gpurun_out/${T}_timeline.txt
