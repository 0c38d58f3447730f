#!/bin/bash
# GPU-side, multi-GPU call of round 2 (gpurun --gpus N -- 'bash scripts/round2_ab_multi.sh N r02'):
#   1. the multi-GPU parity worker (incl. compiled-circuit replays with cached sharded programs)
#   2. A/B of the hoisted exchange (QIPB_SHARD_HOIST) and of the EXT forms on both workloads
N=${1:-2}
R=${2:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q > $O/${R}_pytest_sharded_n$N.log 2>&1; tail -3 $O/${R}_pytest_sharded_n$N.log
PORT=29600
for wl in qft layered; do
 for hoist in 0 1; do
  for ext in 0 1; do
   [ "$wl" = layered ] && [ "$ext" = 1 ] && continue
   PORT=$((PORT+1))
   f=$O/${R}_ab_n${N}_${wl}_hoist${hoist}_ext${ext}.json
   QIPB_SHARD_HOIST=$hoist QIPB_FUSED_EXT=$ext timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
       --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --workload $wl --steps 3 --warmup 3 > $f 2> $O/${R}_abm.err
   python - <<PY
import json
try:
    d = json.load(open("$f"))
    print("N=$N $wl hoist=$hoist ext=$ext  ms/step=%.1f  e2e_ms=%.1f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), d["config"]["stats"],
          {k: (x["launches"], round(x["ms_total"] / x["launches"], 1)) for k, x in d["kernels"].items()})
except Exception as e:
    print("N=$N $wl hoist=$hoist ext=$ext FAILED", e); print(open("$O/${R}_abm.err").read()[-1500:])
PY
  done
 done
done
