"""One dense 6-qubit gate at n = 26 complex128, twice (the second launch is the one ncu captures):
ncu --set full --clock-control none -k regex:big_gate_mma --launch-skip 1 -c 1 -o gpurun_out/x python scripts/big_gate_ncu.py"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qip_b200 import B200Backend                      # noqa: E402
from qip_b200.circuits import haar_unitary            # noqa: E402

n, K = 26, int(sys.argv[1]) if len(sys.argv) > 1 else 6
rng = np.random.default_rng(0)
b = B200Backend.make_state(n, [], [])
b.fuse = False
u = haar_unitary(rng, 2 ** K)
qs = tuple(n - 1 - x for x in reversed(sorted(set(int(round(x)) for x in np.linspace(0, n - 1, K)))))
for _ in range(2):
    b.kronselect_dot({qs: u})
    b.flush()
print("ok", b.total_prob())
