#!/bin/bash
# complex64 layered pass: which of the round-2 forms cost it (A/B of the knobs on one box)
R=${1:-r02c64}
O=gpurun_out
mkdir -p $O
LEAN="--no-micro --no-cpu --no-parity --no-qft --no-configs"
for cfg in "default X=1" "short0 QIPB_FUSED_SHORT_RUNS=0" "ride0 QIPB_FUSED_RIDE2=0" "short0_ride0 QIPB_FUSED_SHORT_RUNS=0 QIPB_FUSED_RIDE2=0" "trio0 QIPB_FUSED_TRIO=0" "all0 QIPB_FUSED_SHORT_RUNS=0 QIPB_FUSED_RIDE2=0 QIPB_FUSED_TRIO=0" "dyn0 QIPB_FUSED_DYNSCHED=0" "wide0 QIPB_FUSED_WIDE=0"; do
 set -- $cfg; name=$1; shift
 env "$@" timeout 300 python bench.py --statetype complex64 --steps 4 --warmup 3 $LEAN > $O/${R}_$name.json 2> $O/${R}.err
 python -c "
import json; d = json.load(open('$O/${R}_$name.json')); print('%-14s c64 ms/step %.1f frac %.3f' % ('$name', d['ms_per_step'], d['roofline']['frac']), {k: (x['launches'], round(x['ms_total']/x['launches'],1)) for k, x in d['kernels'].items()})" || tail -5 $O/${R}.err
done
