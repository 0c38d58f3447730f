#!/bin/bash
# GPU-side, first call of round 2: decide the knobs that were written and CPU-validated while no GPU time was left.
#   1. GPU tier (includes tests/test_zz_gpu_ext.py: the EXT kernel on the device)
#   2. A/B on the same box: QIPB_FUSED_EXT (real 1-qubit sweeps, two QFT steps per sweep) and QIPB_PACK_1Q
#      (lone 1-qubit gates tensored inside a pass) on both workloads at 33 qubits
#      and QIPB_LAZY_INIT (product-state init fused into the first pass) on the e2e path
#   3. ncu --set full of the EXT kernel inside a QFT (30 qubits) for the sweep-level numbers
# Usage: gpurun --timeout 1500 -- 'bash scripts/round2_ab.sh r02'
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${R}_pytest_gpu.log 2>&1; tail -3 $O/${R}_pytest_gpu.log
for wl in qft layered; do
 for ext in 0 1; do
  for pack in 1 0; do
   [ "$wl" = qft ] && [ "$pack" = 0 ] && continue
   f=$O/${R}_ab_${wl}_ext${ext}_pack${pack}.json
   QIPB_FUSED_EXT=$ext QIPB_PACK_1Q=$pack timeout 400 python bench.py --workload $wl --steps 4 --warmup 3 --no-micro --no-cpu > $f 2> $O/${R}_ab.err
   python - <<PY
import json
try:
    d = json.load(open("$f"))
    print("$wl ext=$ext pack=$pack  ms/step=%.1f  e2e_ms=%.1f  passes=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["stats"].get("passes")),
          {k: (x["launches"], round(x["ms_total"] / x["launches"], 1), round(x["GBps"])) for k, x in d["kernels"].items()})
except Exception as e:
    print("$wl ext=$ext pack=$pack FAILED", e); print(open("$O/${R}_ab.err").read()[-1500:])
PY
  done
 done
done
# with EXT the passes move towards the HBM floor, so the pass COUNT matters more: 6 contiguous low bits give 6 free bits per pass
for lb in 6 7; do
 for wl in qft layered; do
  f=$O/${R}_ab_${wl}_ext1_low${lb}.json
  QIPB_FUSED_EXT=1 QIPB_MIN_LOW_BITS=$lb timeout 400 python bench.py --workload $wl --steps 4 --warmup 3 --no-micro --no-cpu > $f 2>> $O/${R}_ab.err
  python - <<PY
import json
try:
    d = json.load(open("$f"))
    print("$wl ext=1 low_bits=$lb  ms/step=%.1f  passes=%s" % (d["ms_per_step"], d["config"]["stats"].get("passes")),
          {k: (x["launches"], round(x["ms_total"] / x["launches"], 1), round(x["GBps"])) for k, x in d["kernels"].items()})
except Exception as e:
    print("$wl ext=1 low_bits=$lb FAILED", e)
PY
 done
done
for lazy in 0 1; do
 f=$O/${R}_ab_lazy${lazy}.json
 QIPB_LAZY_INIT=$lazy timeout 400 python bench.py --steps 3 --warmup 3 --no-micro --no-cpu > $f 2>> $O/${R}_ab.err
 python - <<PY
import json
try:
    d = json.load(open("$f"))
    print("lazy_init=$lazy  e2e_ms=%.1f" % d["e2e"]["ms_per_step"], d["e2e"]["breakdown_untimed_step"])
except Exception as e:
    print("lazy=$lazy FAILED", e)
PY
done
QIPB_FUSED_EXT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 2 -o $O/${R}_prof_fused_ext_qft \
    python bench.py --workload qft --qubits 30 --steps 1 --warmup 1 --no-micro --no-cpu > /dev/null 2>> $O/${R}_ab.err
timeout 300 python scripts/l2_block_probe.py --qubits 28 > $O/${R}_l2_block_probe.txt 2>&1; cat $O/${R}_l2_block_probe.txt
ls -la $O | tail -12
