#!/bin/bash
# gpurun payload: attention timed on a cold GPU, parity tests + report vs both oracles, ncu --set full of the attention kernels
mkdir -p gpurun_out; T=${1:-c3}
timeout 200 python tools/kernel_bench.py --workload large --only attn > gpurun_out/${T}_kb_attn.txt 2>&1; cat gpurun_out/${T}_kb_attn.txt
rm -f gpurun_out/r02_test_parity.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q > gpurun_out/${T}_pytest_model.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_model.log; tail -5 gpurun_out/${T}_pytest_model.log; cat gpurun_out/r02_test_parity.log
timeout 400 python tools/parity_report.py --out gpurun_out/${T}_parity.txt > gpurun_out/${T}_parity_stdout.txt 2>&1; grep -E "^==|bf16-STORAGE|gradients:|SUMMARY" gpurun_out/${T}_parity_stdout.txt
timeout 120 python tools/attn_once.py 0.1 > gpurun_out/${T}_attn_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 3 -c 3 -f -o gpurun_out/${T}_attn python tools/attn_once.py 0.1 > gpurun_out/${T}_attn_ncu.log 2>&1; tail -3 gpurun_out/${T}_attn_ncu.log
