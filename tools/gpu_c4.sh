#!/bin/bash
# gpurun payload: new embed kernel + loader tests, mem-kernel bench, bench with the trainer leg, torch-gpu arm
mkdir -p gpurun_out; T=${1:-c4}
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_loader.py tests/test_gpu_model.py -m gpu -x -q -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -12 gpurun_out/${T}_pytest.log; grep "loader alone" gpurun_out/${T}_pytest.log
timeout 200 python tools/kernel_bench.py --workload large --only mem > gpurun_out/${T}_kb_mem.txt 2>&1; cat gpurun_out/${T}_kb_mem.txt
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
timeout 500 python bench.py --impl torch-gpu --steps 3 --warmup 2 > gpurun_out/${T}_torch_gpu.json 2> gpurun_out/${T}_torch_gpu.err; cat gpurun_out/${T}_torch_gpu.json; tail -3 gpurun_out/${T}_torch_gpu.err
