#!/bin/bash
# A/B on ONE box: tools/ab_bench.sh <workload> <label>=<env assignments separated by commas> ...
# e.g. tools/ab_bench.sh ppi_bp base=SUBGNN_B200_LIB=build/ab/libsubgnn_b200_base.so new=
# Runs bench.py (no CPU baseline) twice per variant, interleaved, and prints ms/step + e2e.
wl=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    label=${v%%=*}; envs=${v#*=}
    line=$(env $(echo $envs | tr ',' ' ') timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 300 2>/dev/null)
    echo "$label rep$rep $(echo $line | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("ms/step %.4f value %.0f e2e %.0f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))')"
  done
done
