#!/bin/bash
# gpurun --gpus N payload: A/B of the gradient all-reduce placement (WM_DP_OVERLAP) and NCCL settings on the DP timeline
N=${1:-8}; T=${2:-dpab$N}; mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dp_timeline.py > gpurun_out/${T}_${tag}.txt 2> gpurun_out/${T}_${tag}.err
  echo "== $tag rc=$?"; head -16 gpurun_out/${T}_${tag}.txt | cut -c1-200; tail -2 gpurun_out/${T}_${tag}.err | cut -c1-200
}
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log | cut -c1-200
run single WM_DP_OVERLAP=0
run single_nvls WM_DP_OVERLAP=0 NCCL_ALGO=NVLS
run overlap_ctas4 WM_DP_OVERLAP=1 NCCL_MAX_CTAS=4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err | cut -c1-200
