#!/bin/bash
mkdir -p gpurun_out; T=${1:-c11}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 120 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.txt 2>&1; tail -1 gpurun_out/${T}_smoke.txt | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json | cut -c1-2600; tail -3 gpurun_out/${T}_bench.err
timeout 300 python tools/step_timeline.py > gpurun_out/${T}_timeline.txt 2>&1; tail -28 gpurun_out/${T}_timeline.txt | cut -c1-140
