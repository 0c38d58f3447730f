#!/bin/bash
# GPU-side: the round's N=1 numbers and ncu evidence (run under gpurun; outputs land in gpurun_out/).
R=${1:-r02}
O=gpurun_out
mkdir -p $O
LEAN="--no-micro --no-cpu --no-parity --no-qft --no-configs"
timeout 900 python -m pytest tests -m gpu -x -q > $O/${R}_pytest_gpu.log 2>&1; tail -2 $O/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.txt 2>&1; tail -1 $O/${R}_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err; tail -c 900 $O/${R}_bench_n1.json; grep "^\[bench\]" $O/${R}_bench_n1.err
timeout 600 python bench.py --workload qft --steps 4 --warmup 3 $LEAN > $O/${R}_bench_qft_n1.json 2>> $O/${R}_bench_n1.err
timeout 600 python bench.py --statetype complex64 --steps 4 --warmup 3 $LEAN > $O/${R}_bench_c64_n1.json 2>> $O/${R}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench_n1.err
# launch list of the same command (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_bench_n1.csv \
    python bench.py --steps 2 --warmup 1 $LEAN > $O/${R}_bench_under_ncu.json 2>> $O/${R}_bench_n1.err
# DRAM traffic of the fused kernel at the bench size
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fused_kernel -c 8 --csv \
    --log-file $O/${R}_traffic_n33.csv python bench.py --steps 2 --warmup 1 $LEAN > /dev/null 2>> $O/${R}_bench_n1.err
# full capture of the fused kernel on the bench workload (30 qubits: ncu replays every launch ~40 times)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 4 -c 4 -o $O/${R}_prof_fused_layered \
    python bench.py --qubits 30 --steps 2 --warmup 1 $LEAN > /dev/null 2>> $O/${R}_bench_n1.err
timeout 300 python scripts/pass_probe.py --workload layered --steps 4 > $O/${R}_pass_probe_layered.txt 2>&1
timeout 300 python scripts/pass_probe.py --workload qft --steps 1 > $O/${R}_pass_probe_qft.txt 2>&1
ls -la $O | tail -14
