#!/bin/bash
mkdir -p gpurun_out; T=${1:-c8}
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_loader.py -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; grep -n "eager   :\|recorded:\|passed\|failed\|FAILED\|Error" gpurun_out/${T}_pytest.log | cut -c1-400
for w in mini small; do
WM_GRAPH_DEBUG=1 timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; cut -c1-420 gpurun_out/${T}_bench_$w.json; grep -o '"e2e_trainer": {[^}]*}' gpurun_out/${T}_bench_$w.json | cut -c1-200; tail -5 gpurun_out/${T}_bench_$w.err | cut -c1-250
done
