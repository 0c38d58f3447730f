"""Symbolic matrix carriers accepted inside a `mats` dict.

These mirror the two carriers of the reference front-end so that op streams can be built without
the reference installed (tests, bench) -- `SwapMat` follows qip/operators.py:157-179 and `CMat`
qip/operators.py:233-268.  The backend recognises carriers by duck typing on `_kron_struct`
(3 = swap with `.n`, 2 = controlled with `.m`), exactly like qip/ext/kronprod.pyx:88-110, so the
reference's own qip.operators.CMat / SwapMat objects are accepted unchanged.
"""
import numpy


class SwapMat(object):
    """2^(2n) x 2^(2n) permutation that exchanges two n-qubit registers (key = a-indices ++ b-indices)."""
    _kron_struct = 3

    def __init__(self, n):
        self.n = int(n)
        self.shape = (2 ** (2 * self.n), 2 ** (2 * self.n))

    def __getitem__(self, item):
        if type(item) == tuple and len(item) == 2:
            mask = (1 << self.n) - 1
            r, c = item
            return 1.0 if (r & mask) == (c >> self.n) and (c & mask) == (r >> self.n) else 0.0
        raise ValueError("SwapMat can only be indexed with M[i,j] not M[{}]".format(item))

    def conj(self):
        return self

    @property
    def T(self):
        return self

    def numpy(self):
        d = self.shape[0]
        return numpy.array([[self[i, j] for j in range(d)] for i in range(d)], dtype=numpy.complex128)

    def __repr__(self):
        return "SwapMat({})".format(self.n)


class CMat(object):
    """Controlled wrapper: identity when the first key index is 0, `m` when it is 1; nests."""
    _kron_struct = 2

    def __init__(self, mat):
        self.m = numpy.array(mat) if type(mat) == list else mat
        self.shape = (self.m.shape[0] * 2, self.m.shape[1] * 2)

    def __getitem__(self, item):
        if type(item) == tuple and len(item) == 2:
            r, c = item
            h = self.shape[0] // 2
            if r < h and c < h:
                return 1.0 if r == c else 0.0
            if r >= h and c >= h:
                return self.m[r - h, c - h]
            return 0.0
        raise ValueError("CMat can only be indexed with M[i,j] not M[{}]".format(item))

    def conj(self):
        return CMat(self.m.conj())

    @property
    def T(self):
        return CMat(self.m.T)

    def numpy(self):
        d = self.shape[0]
        return numpy.array([[self[i, j] for j in range(d)] for i in range(d)], dtype=numpy.complex128)

    def __repr__(self):
        return "CMat({!r})".format(self.m)
