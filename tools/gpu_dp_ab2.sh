#!/bin/bash
N=${1:-8}; T=${2:-dpab2_$N}; mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dp_timeline.py > gpurun_out/${T}_${tag}.txt 2> gpurun_out/${T}_${tag}.err
  echo "== $tag rc=$?"; head -8 gpurun_out/${T}_${tag}.txt | cut -c1-200; grep -i "Last error" -A1 gpurun_out/${T}_${tag}.err | head -3 | cut -c1-200
}
run single_simple WM_DP_OVERLAP=0 NCCL_PROTO=Simple
run single_nvls WM_DP_OVERLAP=0 "NCCL_ALGO=allreduce:nvls"
run single_tree WM_DP_OVERLAP=0 "NCCL_ALGO=allreduce:tree"
