"""Dense K-qubit gates (K = 5..10) on one GPU: time, GB/s and share of the FP64 rate, per batch width.

    python scripts/big_gate_probe.py [n] [K ...]        QIPB_BIG_MMA=0 selects the scalar kernel; PROBE_GBS=8,16,32,64 sweeps the batch width; PROBE_STATETYPE=complex64

Each case also applies U then U^dagger on a random state and reports the distance from the start (1e-15-ish)."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qip_b200 import B200Backend                      # noqa: E402
from qip_b200.circuits import haar_unitary            # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    ks = [int(a) for a in sys.argv[2:]] or [5, 6, 7, 8, 9, 10]
    rng = np.random.default_rng(0)
    st = np.complex64 if os.environ.get("PROBE_STATETYPE") == "complex64" else np.complex128
    b = B200Backend.make_state(n, [], [], statetype=st)
    b.fuse = False
    # a non-trivial state: Hadamard-like dense gates on a few qubits
    for q in (0, n // 2, n - 1):
        b.kronselect_dot({q: haar_unitary(rng, 2)})
    b.flush()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    nbytes = (16.0 if st == np.complex64 else 32.0) * 2.0 ** n
    for K in ks:
        u = haar_unitary(rng, 2 ** K)
        for name, bits in (("spread", sorted(set(int(round(x)) for x in np.linspace(0, n - 1, K)))),
                           ("low", list(range(K))), ("high", list(range(n - K, n)))):
            if len(bits) != K:
                continue
            qs = tuple(n - 1 - x for x in reversed(bits))
            for gb in [None] + [int(x) for x in os.environ.get("PROBE_GBS", "").split(",") if x]:
                if gb is None:
                    os.environ.pop("QIPB_BIG_GB", None)
                else:
                    if (2 ** K) * (gb + 2) * 16 + 8 * (2 ** K + gb) > 220 * 1024:
                        continue
                    os.environ["QIPB_BIG_GB"] = str(gb)
                b.kronselect_dot({qs: u})
                b.flush()
                ts = []
                for _ in range(3):                       # min of three: one launch each, the state (16 GiB at n = 30) is larger than L2
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    b.kronselect_dot({qs: u})
                    b.flush()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) / 1e3)
                t = min(ts)
                fma = 4.0 * 2 ** K * 2.0 ** n
                print("K=%2d %-6s gb=%-4s %8.2f ms  %7.1f GB/s  %5.1f GFMA/s/SM" % (K, name, gb or "auto", t * 1e3, nbytes / t / 1e9, fma / t / sms / 1e9), flush=True)
        os.environ.pop("QIPB_BIG_GB", None)
    # U then U^dagger returns the state (size-independent property), at a size where a copy fits
    m = min(n, 26)
    c = B200Backend.make_state(m, [], [], statetype=st)
    c.fuse = False
    for q in range(0, m, 3):
        c.kronselect_dot({q: haar_unitary(rng, 2)})
    c.flush()
    ref = c.state.clone()
    for K in ks:
        u = haar_unitary(rng, 2 ** K)
        qs = tuple(int(q) for q in rng.permutation(m)[:K])
        c.kronselect_dot({qs: u})
        c.flush()
        moved = float((c.state - ref).abs().max())
        c.kronselect_dot({qs: u.conj().T})
        c.flush()
        back = float((c.state - ref).abs().max()) / float(ref.abs().max())
        print("K=%2d U then U^dagger: moved %.2e, back to %.2e (relative)" % (K, moved, back), flush=True)
        assert back < (1e-5 if st == np.complex64 else 1e-12) and moved > 1e-6


if __name__ == "__main__":
    main()
