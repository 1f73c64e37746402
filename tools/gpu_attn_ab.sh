#!/bin/bash
# gpurun payload: A/B of attention-forward variants (tools/_diag/libwm_b200_<name>.so), each on the same box, timed alone
mkdir -p gpurun_out; T=${1:-attnab}; shift
VARS=${@:-base}
for rep in 1 2; do
for v in $VARS; do
  if [ $v = base ]; then LIB=""; else LIB=tools/_diag/libwm_b200_$v.so; fi
  WM_B200_LIB=$LIB timeout 200 python tools/kernel_bench.py --workload large --only attn 2>&1 | grep attn_ | sed "s/^/$v rep$rep /" | tee -a gpurun_out/${T}_ab.txt
done; done
for v in $VARS; do
  if [ $v = base ]; then LIB=""; else LIB=tools/_diag/libwm_b200_$v.so; fi
  WM_B200_LIB=$LIB timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "attention or golden or oracle" 2>&1 | tail -2 | sed "s/^/$v /"
done
