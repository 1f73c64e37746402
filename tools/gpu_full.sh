#!/bin/bash
# Full validation of the working tree on one B200: GPU tests, smoke, the default bench line, the in-step kernel timeline.
tag=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log; tail -3 gpurun_out/${tag}_smoke.log | cut -c1-600
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${tag}_bench.json
timeout 300 python tools/step_timeline.py > gpurun_out/${tag}_timeline.txt 2>&1; sed -n 3,4p gpurun_out/${tag}_timeline.txt; grep -A16 "^kernel " gpurun_out/${tag}_timeline.txt
