#!/bin/bash
# gpurun --gpus N payload: multi-rank parity test, bench at N, bench at 1 on the same box
N=${1:-8}; T=${2:-dp$N}; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-trainer --no-cpu-baseline > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; cut -c1-300 gpurun_out/${T}_bench_1gpu.json
