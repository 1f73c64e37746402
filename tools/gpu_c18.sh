#!/bin/bash
mkdir -p gpurun_out; T=${1:-c18}
timeout 120 python tools/gemm_roofline_once.py 2>&1 | tail -1
timeout 120 python tools/gemm_roofline_once.py 1 16 0 2>&1 | tail -1
timeout 120 python tools/gemm_roofline_once.py 0 16 2 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 300 python tools/gemm_sites.py large > gpurun_out/${T}_gemm_sites.txt 2>&1; cat gpurun_out/${T}_gemm_sites.txt | cut -c1-330
