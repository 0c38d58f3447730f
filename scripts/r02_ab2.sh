#!/bin/bash
# A/B of the WIDE fused kernel (128 threads, block pairs, EXT forms) against the 256-thread kernels of round 1.
# Usage: gpurun --timeout 1200 -- 'bash scripts/r02_ab2.sh r02b [notest]'
R=${1:-r02b}
O=gpurun_out
mkdir -p $O
if [ "$2" != "notest" ]; then
 timeout 900 python -m pytest tests -m gpu -q -x > $O/${R}_pytest_gpu.log 2>&1; tail -4 $O/${R}_pytest_gpu.log
fi
run() {   # name, env...
 name=$1; shift
 for wl in layered qft; do
  f=$O/${R}_${name}_${wl}.json
  env "$@" timeout 400 python bench.py --workload $wl --steps 4 --warmup 3 --no-micro --no-cpu --no-parity --no-qft --no-configs > $f 2> $O/${R}_ab.err
  python - <<PY
import json
try:
    d = json.load(open("$f"))
    print("%-22s %-8s ms/step=%.1f  e2e_ms=%.1f  frac=%.3f" % ("$name", "$wl", d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]),
          {k: (x["launches"], round(x["ms_total"] / x["launches"], 1)) for k, x in d["kernels"].items()}, d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name $wl FAILED", e); print(open("$O/${R}_ab.err").read()[-1500:])
PY
 done
}
run default
