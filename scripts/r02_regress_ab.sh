#!/bin/bash
# same box: today's library against the build of commit a742c4f (trio sweeps, before ride2 / qft4 / short runs / qft_low),
# layered at complex128 and complex64 -- does the code growth of the WIDE kernel cost the layered passes anything?
R=${1:-r02reg}
O=gpurun_out
mkdir -p $O
LEAN="--no-micro --no-cpu --no-parity --no-qft --no-configs"
OLD=$PWD/qip_b200/csrc/libqipb200_trio.so
one() {  # name, statetype, env...
 name=$1; st=$2; shift 2
 env "$@" timeout 300 python bench.py --statetype $st --steps 8 --warmup 3 $LEAN > $O/${R}_${name}_$st.json 2> $O/${R}.err
 python -c "
import json; d = json.load(open('$O/${R}_${name}_$st.json')); print('%-16s %-10s ms/step %.1f frac %.3f clk %s' % ('$name', '$st', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))" || tail -5 $O/${R}.err
}
for st in complex128 complex64; do
 one new $st X=1
 one old_trio $st QIPB_LIB=$OLD
 one new_again $st X=1
 one new_noride $st QIPB_FUSED_SHORT_RUNS=0 QIPB_FUSED_RIDE2=0
 one new_notrio $st QIPB_FUSED_SHORT_RUNS=0 QIPB_FUSED_RIDE2=0 QIPB_FUSED_TRIO=0
done
