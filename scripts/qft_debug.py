import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from qip_b200 import B200Backend
from qip_b200.circuits import qfft_stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
b = B200Backend.make_state(n, [], [], strategy="tile")
b.profile = []
for mats in qfft_stream(n):
    b.kronselect_dot(mats)
b.flush()
torch.cuda.synchronize()
for name, nbytes, e0, e1 in b.profile:
    print(name, round(e0.elapsed_time(e1), 2), "ms")
print(b.stats)
