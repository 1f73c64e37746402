#!/bin/bash
mkdir -p gpurun_out; T=${1:-c19}
timeout 120 python tools/gemm_two_variants.py > gpurun_out/${T}_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -c 6 -f -o gpurun_out/${T}_gemm python tools/gemm_two_variants.py > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
