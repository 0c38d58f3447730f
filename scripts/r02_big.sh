#!/bin/bash
# FP64-tensor dense-gate kernel: parity tests, then the probe (double-buffered / single-buffered / batch widths)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_kq or raw_c_abi or multi_entry" 2>&1 | tail -5 | tee $O/r02big_pytest.log
timeout 900 python scripts/big_gate_probe.py 30 5 6 7 8 9 10 2>&1 | tee $O/r02big_probe_mma2.txt | tail -30
echo "--- one buffer, batch widths"
QIPB_BIG_NBUF=1 PROBE_GBS=16,32,64 timeout 900 python scripts/big_gate_probe.py 30 5 6 7 2>&1 | grep -v dagger | tee $O/r02big_probe_nbuf1.txt | tail -40
echo "--- two buffers, batch widths"
QIPB_BIG_NBUF=2 PROBE_GBS=8,16,32,64 timeout 900 python scripts/big_gate_probe.py 30 5 6 7 8 2>&1 | grep -v dagger | tee $O/r02big_probe_nbuf2.txt | tail -50
