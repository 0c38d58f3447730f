"""Vectorised oracle functions for `F(func, reg1, reg2)`.

The reference tabulates `func` with 2^|reg1| python calls on every `func_apply`
(qip/ext/func_apply.pyx:66-69).  The functions here accept python ints AND numpy int64 arrays, so
B200Backend.func_apply tabulates them with one vectorised call; `tabulated(f, nbits)` attaches the table
itself (built once, reused by every application and by compiled circuits); `controlled(f, c, m)` is C(F) as a plain F
on a joined register."""
import numpy as np


def equals(x0: int):
    """x -> 1 if x == x0 else 0: the Grover oracle of examples/grovers_iterative.py:20-21."""
    def f(x):
        return (x == x0) * 1
    f.vectorized = True                     # accepts numpy int64 arrays (backend.tabulate trusts the vectorised call)
    return f


def modexp(base: int, modulus: int):
    """i -> base**i mod modulus: the Shor oracle of examples/shors.py:113 (`pow(x, i, N)`), by
    square-and-multiply on int64 arrays (exact while modulus < 2**31)."""
    base, modulus = int(base), int(modulus)
    if not (0 < modulus < 2 ** 31):
        raise ValueError("modexp needs 0 < modulus < 2**31")

    def f(i):
        if isinstance(i, (int, np.integer)):
            return pow(base, int(i), modulus)
        e = np.asarray(i, dtype=np.int64).copy()
        result = np.ones_like(e)
        b = np.int64(base % modulus)
        while np.any(e > 0):
            odd = (e & 1) == 1
            result[odd] = (result[odd] * b) % modulus
            b = (b * b) % modulus
            e >>= 1
        return result
    f.vectorized = True                     # exact on int64 arrays while modulus < 2**31 (checked above)
    return f


def controlled(func, n_controls: int, n_inputs: int):
    """Controlled F as a plain F on a joined register: apply `F(controlled(f, c, m), joined, reg2)` where `joined` holds the
    c control qubits FIRST and then the m input qubits (x is read big-endian over reg1, qip/ext/func_apply.pyx:81-103, so the
    controls are the most significant bits).  The result is f(x) when every control is 1 and 0 otherwise, and q xor 0 leaves
    reg2 alone.  The reference has no controlled F (`COp` only wraps `MatrixOp`s, qip/operators.py:183-231); this needs no
    new kernel: the extended table goes through the same `func_apply` path, on one GPU and sharded."""
    n_controls, n_inputs = int(n_controls), int(n_inputs)
    if n_controls < 0 or n_inputs < 0:
        raise ValueError("n_controls and n_inputs must be non-negative")
    full, mask = (1 << n_controls) - 1, (1 << n_inputs) - 1
    vec = bool(getattr(func, "vectorized", False))

    def f(xp):
        if isinstance(xp, (int, np.integer)):
            xp = int(xp)
            return int(func(xp & mask)) if (xp >> n_inputs) == full else 0
        xp = np.asarray(xp, dtype=np.int64)
        on = (xp >> n_inputs) == full
        out = np.zeros_like(xp)
        if on.any():
            x = xp[on] & mask
            out[on] = np.asarray(func(x), dtype=np.int64) if vec else np.array([int(func(int(v))) for v in x], dtype=np.int64)
        return out
    f.vectorized = True                     # arrays are handled here (element by element when `func` itself is scalar-only)
    return f


class tabulated(object):
    """`func` with its table over nbits input bits attached (`.table`, int64)."""

    def __init__(self, func, nbits: int):
        from .backend import tabulate
        self.func = func
        self.table = tabulate(func, nbits)

    def __call__(self, x):
        return self.table[x]
