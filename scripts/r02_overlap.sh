#!/bin/bash
# Exchange / compute overlap on N GPUs: parity tier with the pipeline on, then A/B of QIPB_SHARD_OVERLAP on the bench line.
# Usage: gpurun --gpus N -- 'bash scripts/r02_overlap.sh N r02e'
N=${1:-2}
R=${2:-r02e}
O=gpurun_out
mkdir -p $O
if [ -z "$NOTEST" ]; then
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s > $O/${R}_pytest_sharded_n$N.log 2>&1; tail -8 $O/${R}_pytest_sharded_n$N.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29700
CFGS=${CFGS:-"ov0 QIPB_SHARD_OVERLAP=0|ov1_c3 QIPB_SHARD_OVERLAP=1 QIPB_OVERLAP_CHUNK_BITS=3|ov1_c2 QIPB_SHARD_OVERLAP=1 QIPB_OVERLAP_CHUNK_BITS=2"}
IFS='|' read -ra CFG_LIST <<< "$CFGS"
for cfg in "${CFG_LIST[@]}"; do
 set -- $cfg; name=$1; shift
 PORT=$((PORT+1))
 f=$O/${R}_bench_n${N}_$name.json
 env "$@" timeout 600 $TR --master-port $PORT bench.py --gpus $N --steps 6 --warmup 3 --no-parity > $f 2> $O/${R}_bench_$name.err
 python - <<PY
import json
try:
    d = json.load(open("$f"))
    q = d["qft"]
    print("N=$N %-12s layered ms/step=%.1f e2e=%.1f | qft %.3f s (err %.1e, nvlink %.0f ms)" % ("$name", d["ms_per_step"], d["e2e"]["ms_per_step"], q["seconds"], q["parity_max_err"], q.get("nvlink_ms", 0)),
          {k: (x["launches"], round(x["ms_total"] / x["launches"], 1)) for k, x in d["kernels"].items()}, {k: v for k, v in d["config"]["stats"].items() if "overlap" in k or k == "exchanges"})
except Exception as e:
    print("$name FAILED", e); print(open("$O/${R}_bench_$name.err").read()[-2500:])
PY
done
