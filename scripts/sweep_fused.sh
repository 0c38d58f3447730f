#!/bin/bash
# GPU-side sweep of fused-kernel variants (run under gpurun): prints one line per configuration.
for wl in layered qft; do
 for v in 0 2; do
  for ca in 0 1; do
   QIPB_FUSED_VARIANT=$v QIPB_COST_AWARE=$ca timeout 300 python bench.py --workload $wl --steps 2 --warmup 1 --no-micro --no-cpu --strategy tile > gpurun_out/sw_${wl}_${v}_${ca}.json 2> gpurun_out/sw.err
   python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_${wl}_${v}_${ca}.json"))
    print("$wl var=$v cost_aware=$ca  ms/step=%.0f value=%.0f" % (d["ms_per_step"], d["value"]), {k:(x["launches"], round(x["ms_total"]), round(x["GBps"])) for k,x in d["kernels"].items()})
except Exception as e:
    print("$wl var=$v ca=$ca FAILED", e); print(open("gpurun_out/sw.err").read()[-1500:])
PY
  done
 done
done
