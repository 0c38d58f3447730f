#!/bin/bash
# Compile only the WIDE complex128 instantiation of the fused kernel and report what matters in its SASS
# (registers / spills, whether the descriptor index stayed on the uniform datapath, the hot loops).
cd $(dirname $0)/../qip_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xptxas -v -DQIPB_PROBE_WIDE_ONLY "$@" -c fused.cu -o /tmp/f_probe.o 2> /tmp/f_probe.log || { tail -20 /tmp/f_probe.log; exit 1; }
grep -A2 "fused_kernel" /tmp/f_probe.log | grep -E "registers|spill"
cuobjdump -sass /tmp/f_probe.o > /tmp/probe.sass
echo "UIMAD(desc) $(grep -c 'UIMAD.*0x150' /tmp/probe.sass)  IMAD(desc) $(grep -c ' IMAD.*0x150' /tmp/probe.sass)  LDCU $(grep -c LDCU /tmp/probe.sass)  LDC $(grep -c 'LDC\.' /tmp/probe.sass) lines $(wc -l < /tmp/probe.sass)"
