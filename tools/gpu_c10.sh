#!/bin/bash
# gpurun payload: ncu --set full of the roofline GEMM (tuned variant) + ncu launch list of the default bench command
mkdir -p gpurun_out; T=${1:-c10}
timeout 120 python tools/gemm_roofline_once.py > gpurun_out/${T}_gemm_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 6 -c 2 -f -o gpurun_out/${T}_gemm python tools/gemm_roofline_once.py > gpurun_out/${T}_gemm_ncu.log 2>&1; cat gpurun_out/${T}_gemm_plain.log; tail -2 gpurun_out/${T}_gemm_ncu.log
WM_GEMM_TUNE=0 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > gpurun_out/${T}_bench_plain.log 2>&1 && \
WM_GEMM_TUNE=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > gpurun_out/${T}_bench_ncu.log 2>&1; tail -2 gpurun_out/${T}_bench_ncu.log | cut -c1-300; wc -l gpurun_out/${T}_launches.csv
