#!/bin/bash
mkdir -p gpurun_out; T=${1:-c20}
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python tools/gemm_sites.py large > gpurun_out/${T}_sites.txt 2>&1; tail -12 gpurun_out/${T}_sites.txt
timeout 200 python tools/gemm_roofline_once.py > gpurun_out/${T}_roof.txt 2>&1; tail -4 gpurun_out/${T}_roof.txt
timeout 300 python bench.py --no-trainer > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 1500 gpurun_out/${T}_bench.json
