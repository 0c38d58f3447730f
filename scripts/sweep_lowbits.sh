#!/bin/bash
# GPU-side: bench the two workloads at 33 qubits for several tile shapes (contiguous low bits of the tile).
for wl in layered qft; do
 for lb in 5 6 7; do
  QIPB_MIN_LOW_BITS=$lb timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-micro --no-cpu > gpurun_out/lb_${wl}_${lb}.json 2> gpurun_out/lb.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/lb_${wl}_${lb}.json"))
    print("$wl low_bits=$lb ms/step=%.1f e2e_ms=%.1f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), d["config"]["stats"], {k:(x["launches"], round(x["ms_total"]/x["launches"],1), round(x["GBps"])) for k,x in d["kernels"].items()})
except Exception as e:
    print("$wl lb=$lb FAILED", e); print(open("gpurun_out/lb.err").read()[-1500:])
PY
 done
done
