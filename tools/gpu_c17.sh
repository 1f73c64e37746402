#!/bin/bash
for rep in 1 2; do
for o in "ln_bwd_width=8,ln_fwd_width=8" "ln_bwd_width=4,ln_fwd_width=4"; do
WM_OPTIONS=$o timeout 200 python tools/kernel_bench.py --workload large --only mem 2>&1 | grep -i "layernorm_fwd\|encoder form" | sed "s/^/$o rep$rep /" | cut -c1-200
done; done
