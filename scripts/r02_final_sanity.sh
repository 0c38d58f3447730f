#!/bin/bash
# last call of the round: the driver's sequence on the final tree (GPU tests, smoke, the default bench command)
R=${1:-r02fin}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${R}_pytest_gpu.log 2>&1; tail -2 $O/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.txt 2>&1; tail -1 $O/${R}_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
python -c "
import json; d = json.load(open('$O/${R}_bench_n1.json'))
print('N=1 ms/step %.1f frac %.3f e2e %.1f qft %.3f s (%.1e) parity %s clk %s' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['qft']['seconds'], d['qft']['parity_max_err'], (d['parity']['cases'], d['parity']['max_err'], d['parity']['failed']), d['clocks']['sm_mhz']))
print(d['micro']['other_kernels']); print({k: (v.get('latency_ms') or v.get('device_ms') or v.get('ms_per_iteration')) for k, v in d['configs'].items()})"
timeout 600 python bench.py --statetype complex64 --steps 4 --warmup 3 --no-micro --no-cpu --no-parity --no-configs > $O/${R}_bench_c64_n1.json 2>> $O/${R}_bench_n1.err
python -c "
import json; d = json.load(open('$O/${R}_bench_c64_n1.json')); print('c64 ms/step %.1f frac %.3f qft %.3f' % (d['ms_per_step'], d['roofline']['frac'], d['qft']['seconds']))"
