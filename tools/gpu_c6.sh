#!/bin/bash
# gpurun payload: recorded-step tests, embed timing, small-workload bench lines (eager vs graph), default bench
mkdir -p gpurun_out; T=${1:-c6}
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_kernels.py tests/test_gpu_loader.py tests/test_gpu_cli.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -25 gpurun_out/${T}_pytest.log
timeout 200 python tools/kernel_bench.py --workload large --only mem > gpurun_out/${T}_kb_mem.txt 2>&1; grep embed gpurun_out/${T}_kb_mem.txt
for w in mini small medium yield; do
timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; cut -c1-700 gpurun_out/${T}_bench_$w.json; grep -o '"e2e_trainer": {[^}]*}' gpurun_out/${T}_bench_$w.json | cut -c1-200; tail -3 gpurun_out/${T}_bench_$w.err
timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --graph 0 --no-trainer > gpurun_out/${T}_bench_${w}_eager.json 2> gpurun_out/${T}_bench_${w}_eager.err; cut -c1-330 gpurun_out/${T}_bench_${w}_eager.json
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json | cut -c1-3000; tail -5 gpurun_out/${T}_bench.err
