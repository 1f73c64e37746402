#!/bin/bash
# Final numbers of a build: default bench line, then the other BASELINE configs. Usage: tools/gpu_final.sh <tag>
mkdir -p gpurun_out; T=${1:-final}
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
for w in mini small medium yield; do
timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "$w rc=$?"
done
python - <<PY
import json
for n in ["bench", "bench_mini", "bench_small", "bench_medium", "bench_yield"]:
    d = json.loads(open(f"gpurun_out/${T}_{n}.json").read().strip().splitlines()[-1])
    t = d.get("e2e_trainer") or {}
    r = d.get("roofline") or {}
    print(n, round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "trainer", t.get("value") and round(t["value"]), "roof", r.get("frac") and round(r["frac"], 3), d["clocks"]["sm_mhz"], round(d.get("step_tensor_frac_of_sustained", 0), 3))
PY
