#!/bin/bash
# per-launch anatomy of the fused passes + one ncu capture of a layered pass (WIDE kernel)
R=${1:-r02c}
O=gpurun_out
mkdir -p $O
timeout 300 python scripts/pass_probe.py --workload layered --steps 4 > $O/${R}_pass_probe_layered.txt 2>&1; cat $O/${R}_pass_probe_layered.txt
timeout 300 python scripts/pass_probe.py --workload qft --steps 1 > $O/${R}_pass_probe_qft.txt 2>&1; cat $O/${R}_pass_probe_qft.txt
QIPB_FUSED_PAIR=0 timeout 300 python scripts/pass_probe.py --workload layered --steps 2 > $O/${R}_pass_probe_layered_nopair.txt 2>&1; cat $O/${R}_pass_probe_layered_nopair.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 6 -c 4 -o $O/${R}_prof_fused_wide_layered \
    python bench.py --workload layered --qubits 30 --steps 2 --warmup 1 --no-micro --no-cpu --no-parity --no-qft --no-configs > /dev/null 2> $O/${R}_ncu.err
ls -la $O | tail -5
