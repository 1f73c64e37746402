#!/bin/bash
mkdir -p gpurun_out; T=${1:-c12}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 200 python tools/kernel_bench.py --workload large --only mem 2>&1 | grep -i "adam\|embed"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json | cut -c1-700; tail -3 gpurun_out/${T}_bench.err
for w in mini small; do
timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; cut -c1-330 gpurun_out/${T}_bench_$w.json; grep -o '"gpu_launches": [0-9]*' gpurun_out/${T}_bench_$w.json; grep -o '"e2e_trainer": {[^}]*}' gpurun_out/${T}_bench_$w.json | cut -c1-200; tail -3 gpurun_out/${T}_bench_$w.err | cut -c1-250
done
