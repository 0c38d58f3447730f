#!/bin/bash
# Multi-GPU call of round 2: gpurun --gpus N -- 'bash scripts/r02_multi.sh N r02'
#   1. the multi-GPU parity tier (tests/test_gpu_sharded.py -> tests/dist_gpu_worker.py), log kept for profiles/
#   2. bench.py at N (parity block before the timed region, layered step, QFFT line = BASELINE metric M2, e2e)
#   3. the reference arm under torchrun (OMP fix), short
#   4. optional extras: 36-qubit complex64 QFT (N = 4, 8), strong scaling of a fixed 33-qubit state
N=${1:-2}
R=${2:-r02}
EXTRA=${3:-}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s > $O/${R}_pytest_sharded_n$N.log 2>&1; tail -25 $O/${R}_pytest_sharded_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29611 bench.py --gpus $N --steps 6 --warmup 3 > $O/${R}_bench_n$N.json 2> $O/${R}_bench_n$N.err
grep "^\[bench\]" $O/${R}_bench_n$N.err
python - <<PY
import json
try:
    d = json.load(open("$O/${R}_bench_n$N.json"))
    print("N=$N layered ms/step=%.1f value=%.0f e2e=%s" % (d["ms_per_step"], d["value"], {k: d["e2e"][k] for k in ("ms_per_step", "layers_per_step", "exchanges_per_step")}))
    print(" parity", d["parity"])
    print(" qft", {k: v for k, v in d["qft"].items() if k != "stats"})
    print(" kernels", {k: (x["launches"], round(x["ms_total"] / x["launches"], 1)) for k, x in d["kernels"].items()}, d["config"]["stats"])
except Exception as e:
    print("bench N=$N FAILED", e); print(open("$O/${R}_bench_n$N.err").read()[-3000:])
PY
timeout 600 $TR --master-port 29612 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/${R}_bench_reference_n$N.json 2> $O/${R}_ref.err
cut -c1-600 $O/${R}_bench_reference_n$N.json
if [ -n "$EXTRA" ]; then
 QIPB_SHARD_OVERLAP=0 timeout 600 $TR --master-port 29615 bench.py --gpus $N --steps 6 --warmup 3 --no-parity > $O/${R}_bench_n${N}_overlap0.json 2> $O/${R}_x.err
 python -c "
import json; d = json.load(open('$O/${R}_bench_n${N}_overlap0.json')); print('overlap OFF N=$N layered ms/step %.1f, qft %.3f s' % (d['ms_per_step'], d['qft']['seconds']))" || tail -20 $O/${R}_x.err
 timeout 600 $TR --master-port 29613 bench.py --gpus $N --workload qft --statetype complex64 --total-qubits 36 --steps 2 --warmup 1 --no-parity > $O/${R}_bench_qft_c64_n$N.json 2> $O/${R}_x.err
 python -c "
import json; d = json.load(open('$O/${R}_bench_qft_c64_n$N.json')); print('qft c64 N=$N qubits', d['config']['qubits'], 'seconds', d.get('qft_seconds'), d['config']['stats'])" || tail -20 $O/${R}_x.err
 timeout 600 $TR --master-port 29614 bench.py --gpus $N --workload qft --total-qubits 33 --steps 2 --warmup 1 --no-parity > $O/${R}_bench_qft_strong33_n$N.json 2> $O/${R}_x.err
 python -c "
import json; d = json.load(open('$O/${R}_bench_qft_strong33_n$N.json')); print('qft strong 33q N=$N seconds', d.get('qft_seconds'), d['config']['stats'])" || tail -20 $O/${R}_x.err
fi
ls -la $O | tail -8
