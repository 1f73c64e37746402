#!/bin/bash
mkdir -p gpurun_out; T=${1:-c13}
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_graph.py tests/test_yield_path.py tests/test_siblings.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
for w in small medium mini; do
timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; cut -c1-330 gpurun_out/${T}_bench_$w.json; grep -o '"gpu_launches": [0-9]*' gpurun_out/${T}_bench_$w.json; grep -o '"e2e_trainer": {[^}]*}' gpurun_out/${T}_bench_$w.json | cut -c1-200; tail -3 gpurun_out/${T}_bench_$w.err | cut -c1-250
done
timeout 200 python tools/kernel_bench.py --workload large --only wgrad 2>&1 | tail -6
