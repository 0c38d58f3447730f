#!/bin/bash
# FP64-tensor dense-gate kernel: parity tests (complex128 + complex64), then the probe for both state types
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_kq or raw_c_abi or multi_entry or golden_stream_complex64" 2>&1 | tail -5 | tee $O/r02big_pytest.log
timeout 600 python scripts/big_gate_probe.py 30 5 6 7 8 9 10 2>&1 | tee $O/r02big_probe_c128_final.txt | tail -30
echo "--- complex64, tensor path"
PROBE_STATETYPE=complex64 timeout 600 python scripts/big_gate_probe.py 30 5 6 8 10 2>&1 | tee $O/r02big_probe_c64_mma.txt | tail -20
echo "--- complex64, scalar kernel"
PROBE_STATETYPE=complex64 QIPB_BIG_MMA=0 timeout 600 python scripts/big_gate_probe.py 30 5 6 8 2>&1 | tee $O/r02big_probe_c64_scalar.txt | tail -14
