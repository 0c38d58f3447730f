#!/bin/bash
# strong scaling of a fixed 33-qubit complex128 QFFT on N GPUs
N=${1:-2}
R=${2:-r02str}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for wl in qft layered; do
 timeout 600 $TR --master-port 29950 bench.py --gpus $N --workload $wl --total-qubits 33 --steps 3 --warmup 2 --no-parity --no-qft > $O/${R}_${wl}_strong33_n$N.json 2> $O/${R}.err
 python -c "
import json; d = json.load(open('$O/${R}_${wl}_strong33_n$N.json')); print('strong 33q N=$N %-8s ms/step %.1f  stats %s' % ('$wl', d['ms_per_step'], {k: v for k, v in d['config']['stats'].items() if k in ('exchanges', 'overlapped_exchanges')}))" || tail -5 $O/${R}.err
done
