#!/bin/bash
# gpurun payload: the round's standard single-GPU check (tests, parity report, attention A/B, step bench, smoke)
mkdir -p gpurun_out; T=${1:-c1}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -15 gpurun_out/${T}_pytest.log
timeout 300 python tools/parity_report.py --out gpurun_out/${T}_parity.txt > gpurun_out/${T}_parity_stdout.txt 2>&1; tail -3 gpurun_out/${T}_parity_stdout.txt
timeout 200 python tools/kernel_bench.py --workload large --only attn > gpurun_out/${T}_kb_attn.txt 2>&1; cat gpurun_out/${T}_kb_attn.txt
if [ -f tools/_diag/libwm_b200_intpack.so ]; then
WM_B200_LIB=tools/_diag/libwm_b200_intpack.so timeout 200 python tools/kernel_bench.py --workload large --only attn > gpurun_out/${T}_kb_attn_intpack.txt 2>&1; cat gpurun_out/${T}_kb_attn_intpack.txt
fi
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 120 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt
