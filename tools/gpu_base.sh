#!/bin/bash
# gpurun payload: full single-GPU check of the current tree (tests, smoke, step bench, kernel bench, timeline)
mkdir -p gpurun_out; T=${1:-base}
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${T}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 300 python tools/kernel_bench.py --workload large > gpurun_out/${T}_kb.txt 2>&1; cat gpurun_out/${T}_kb.txt
timeout 300 python tools/step_timeline.py > gpurun_out/${T}_timeline.txt 2>&1; tail -32 gpurun_out/${T}_timeline.txt
