#!/bin/bash
# gpurun payload: yield head + embed tests, default bench (roofline first, tuned variant, 40-step trainer leg), other workloads
mkdir -p gpurun_out; T=${1:-c5}
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_yield_path.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -8 gpurun_out/${T}_pytest.log
timeout 200 python tools/kernel_bench.py --workload large --only mem > gpurun_out/${T}_kb_mem.txt 2>&1; grep embed gpurun_out/${T}_kb_mem.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
for w in small medium mini yield; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; cat gpurun_out/${T}_bench_$w.json; tail -3 gpurun_out/${T}_bench_$w.err
done
