#!/bin/bash
# The other BASELINE configs (bench.py --workload ...), then the ncu launch list of the default bench (heuristic kernel
# variants: the start-up tuner's launches would fill the list). Usage: tools/gpu_workloads.sh <tag>
mkdir -p gpurun_out; T=${1:-wl}
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k gemm > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
for w in mini small medium yield; do
timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; cut -c1-200 gpurun_out/${T}_bench_$w.json; grep -o '"e2e_trainer": {[^}]*}' gpurun_out/${T}_bench_$w.json | cut -c1-160; tail -2 gpurun_out/${T}_bench_$w.err | cut -c1-200
done
WM_GEMM_TUNE=0 timeout 300 python bench.py --steps 2 --warmup 3 --no-trainer --no-cpu-baseline > gpurun_out/${T}_notune.json 2> gpurun_out/${T}_notune.err && \
WM_GEMM_TUNE=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-trainer --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log | cut -c1-300; cut -c1-200 gpurun_out/${T}_notune.json
