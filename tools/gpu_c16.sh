#!/bin/bash
mkdir -p gpurun_out; T=${1:-c16}
timeout 200 python tools/kernel_bench.py --workload large --only mem 2>&1 | grep -i "layernorm"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
