#!/bin/bash
mkdir -p gpurun_out; T=${1:-c9}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log | cut -c1-300
for w in mini small medium; do
timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; cut -c1-330 gpurun_out/${T}_bench_$w.json; grep -o '"e2e_trainer": {[^}]*}' gpurun_out/${T}_bench_$w.json | cut -c1-200; tail -3 gpurun_out/${T}_bench_$w.err | cut -c1-250
done
timeout 120 python tools/sanitize_once.py > gpurun_out/${T}_sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/${T}_memcheck.log python tools/sanitize_once.py > gpurun_out/${T}_memcheck_stdout.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${T}_sanitize_plain.log; tail -5 gpurun_out/${T}_memcheck.log; tail -3 gpurun_out/${T}_memcheck_stdout.log
