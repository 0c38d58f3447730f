#!/bin/bash
# FP64-tensor dense-gate kernel: parity tests (default batch widths, then 128 columns forced), probe with width sweep at K = 5, 6
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_kq or raw_c_abi or multi_entry" 2>&1 | tail -2 | tee $O/r02big_pytest.log
QIPB_BIG_GB=128 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_kq" 2>&1 | tail -2 | tee -a $O/r02big_pytest.log
PROBE_GBS=64,128 timeout 300 python scripts/big_gate_probe.py 30 5 6 2>&1 | tee $O/r02big_probe_gb128.txt | tail -30
