#!/bin/bash
R=${1:-r02c64}
O=gpurun_out
mkdir -p $O
LEAN="--no-micro --no-cpu --no-parity --no-qft --no-configs"
timeout 900 python -m pytest tests -m gpu -q -x > $O/${R}_pytest_gpu.log 2>&1; tail -2 $O/${R}_pytest_gpu.log
for wl in layered qft; do
 for st in complex64 complex128; do
  timeout 300 python bench.py --workload $wl --statetype $st --steps 8 --warmup 3 $LEAN > $O/${R}_${wl}_$st.json 2> $O/${R}.err
  python -c "
import json; d = json.load(open('$O/${R}_${wl}_$st.json')); print('%-8s %-10s n=%d ms/step %.1f frac %.3f clk %s' % ('$wl', '$st', d['config']['qubits'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))" || tail -5 $O/${R}.err
 done
done
