#!/bin/bash
# 4-GPU call of round 2: parity tier, the bench line (35 qubits c128), 36-qubit complex64 QFFT, and a sweep of the remap CTA count
R=${1:-r02r}
N=4
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s > $O/${R}_pytest_sharded_n$N.log 2>&1; tail -4 $O/${R}_pytest_sharded_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29900
show() {
python - <<PY
import json
try:
    d = json.load(open("$1"))
    q = d.get("qft") or {}
    print("%-14s N=4 ms/step=%.1f e2e=%.1f | qft %s s | " % ("$2", d["ms_per_step"], d["e2e"]["ms_per_step"], q.get("seconds")),
          {k: (x["launches"], round(x["ms_total"] / x["launches"], 1)) for k, x in d["kernels"].items()}, {k: v for k, v in d["config"]["stats"].items() if "overlap" in k or k == "exchanges"}, d.get("parity") and (d["parity"]["cases"], d["parity"]["max_err"]))
except Exception as e:
    print("$2 FAILED", e)
PY
}
timeout 900 $TR --master-port 29901 bench.py --gpus $N --steps 6 --warmup 3 > $O/${R}_bench_n4.json 2> $O/${R}_bench_n4.err; show $O/${R}_bench_n4.json default
for cfg in "cta025 QIPB_XCHG_CTAS_PER_SM=0.25" "cta1 QIPB_XCHG_CTAS_PER_SM=1" "ov0 QIPB_SHARD_OVERLAP=0"; do
 set -- $cfg; name=$1; shift
 PORT=$((PORT+1))
 env "$@" timeout 600 $TR --master-port $PORT bench.py --gpus $N --steps 6 --warmup 3 --no-parity > $O/${R}_bench_n4_$name.json 2> $O/${R}_x.err; show $O/${R}_bench_n4_$name.json $name
done
timeout 600 $TR --master-port 29911 bench.py --gpus $N --workload qft --statetype complex64 --total-qubits 36 --steps 2 --warmup 1 --no-parity > $O/${R}_bench_qft_c64_36q_n4.json 2> $O/${R}_x.err
python -c "
import json; d = json.load(open('$O/${R}_bench_qft_c64_36q_n4.json')); print('qft c64 36q N=4 seconds', d.get('qft_seconds'), d['config']['qubits'], d['config']['stats'])" || tail -20 $O/${R}_x.err
