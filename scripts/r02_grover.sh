#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grover or refeed or device" 2>&1 | tail -3
timeout 300 python scripts/grover_probe.py 4 1 2>&1 | tee $O/r02_grover_probe_adopt.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-parity --no-qft --no-micro --no-cpu > $O/r02_bench_configs_check.json 2> $O/r02_bench_configs_check.err
python -c "
import json; d = json.load(open('$O/r02_bench_configs_check.json')); g = d['configs']['grover28']; print('grover', g['ms_per_iteration'], g['seconds_all_repeats'], g['abs_err_vs_closed_form'])"
