#!/bin/bash
# gpurun --gpus N payload: multi-rank parity test, DP timeline, bench at N
N=${1:-2}; T=${2:-dp$N}; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_timeline.py > gpurun_out/${T}_timeline.txt 2> gpurun_out/${T}_timeline.err; cat gpurun_out/${T}_timeline.txt | cut -c1-260; tail -3 gpurun_out/${T}_timeline.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-900 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
