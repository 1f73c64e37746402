#!/bin/bash
mkdir -p gpurun_out; T=${1:-c14}
timeout 300 python tools/cublas_compare.py large > gpurun_out/${T}_cublas.txt 2>&1; tail -22 gpurun_out/${T}_cublas.txt | cut -c1-220
timeout 300 python tools/kernel_bench.py --workload large > gpurun_out/${T}_kb.txt 2>&1; tail -34 gpurun_out/${T}_kb.txt | cut -c1-160
