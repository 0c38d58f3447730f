"""Tests-only: random sessions against a backend with the reference's StateType surface and the CPU oracle side by side."""
import random

import numpy as np

from oracle import oracle as orc
from qip_b200.circuits import H2, X2, haar_unitary, rm_mat
from qip_b200.mats import CMat, SwapMat


def rand_mats(rng, n):
    kind = int(rng.integers(0, 9))
    qs = [int(x) for x in rng.permutation(n)]
    if kind == 0:
        return {qs[0]: H2}
    if kind == 1:
        return {qs[0]: rm_mat(int(rng.integers(1, 6)))}
    if kind == 2:
        return {(qs[0], qs[1]): CMat(X2)}
    if kind == 3:
        return {(qs[0], qs[1]): SwapMat(1)}
    if kind == 4:
        return {(qs[0], qs[1]): haar_unitary(rng, 4)}
    if kind == 5:
        return {(qs[0], qs[1], qs[2]): CMat(CMat(haar_unitary(rng, 2)))}
    if kind == 6:
        return {(qs[0], qs[1], qs[2]): haar_unitary(rng, 8)}
    if kind == 7:
        return {(qs[0], qs[1], qs[2], qs[3], qs[4]): CMat(SwapMat(2))}
    return {qs[0]: H2, qs[1]: haar_unitary(rng, 2), (qs[2], qs[3]): CMat(rm_mat(2))}


def rand_feeds(rng, n):
    """Random grouping of the qubits into vector feeds, one-hot int feeds and un-fed qubits."""
    qs = [int(x) for x in rng.permutation(n)]
    groups, product_feeds, oracle_feeds = [], [], []
    i = 0
    while i < n:
        L = int(min(n - i, rng.integers(1, 4)))
        g = qs[i:i + L]
        i += L
        r = rng.random()
        if r < 0.2:
            continue
        groups.append(g)
        if r < 0.4:
            v = int(rng.integers(0, 2 ** L))
            product_feeds.append(v)
            oracle_feeds.append(np.eye(2 ** L)[v])
        else:
            v = rng.normal(size=2 ** L) + 1j * rng.normal(size=2 ** L)
            v /= np.linalg.norm(v)
            product_feeds.append(v)
            oracle_feeds.append(v)
    return groups, product_feeds, oracle_feeds


def session(make_state, seed, n, steps=None, **kw):
    """One random session: gates of every kind, both measurement conventions, sampling with a fixed draw, func_apply."""
    rng = np.random.default_rng(seed)
    groups, pf, of = rand_feeds(rng, n)
    g = make_state(n, groups, pf, **kw)
    c = orc.OracleBackend.make_state(n, groups, of)
    for step in range(int(rng.integers(3, 28)) if steps is None else steps):
        r = rng.random()
        if r < 0.7:
            m = rand_mats(rng, n)
            g.kronselect_dot(m)
            c.kronselect_dot(m)
        elif r < 0.78:
            idx = [int(x) for x in rng.permutation(n)[:int(rng.integers(1, 4))]]
            assert np.allclose(g.measure_probabilities(np.array(idx, dtype=np.int32)), c.measure_probabilities(idx), rtol=0, atol=1e-12)
        elif r < 0.85:
            idx = [int(x) for x in rng.permutation(n)[:2]]
            random.seed(seed + step)
            a = g.measure(np.array(idx, dtype=np.int32))
            random.seed(seed + step)
            b = c.measure(idx)
            assert a[0] == b[0] and abs(a[1] - b[1]) < 1e-12, ("measure", a, b)
        elif r < 0.92:
            qq = [int(x) for x in rng.permutation(n)[:4]]
            f = lambda x: (3 * x + 1) % 4
            g.func_apply(qq[:2], qq[2:], f)
            c.func_apply(qq[:2], qq[2:], f)
        elif r < 0.96:
            idx = [int(x) for x in rng.permutation(n)[:2]]
            random.seed(seed + step)
            a = g.soft_measure(np.array(idx, dtype=np.int32))
            random.seed(seed + step)
            b = c.soft_measure(idx)
            assert a[0] == b[0] and abs(a[1] - b[1]) < 1e-12, ("soft_measure", a, b)
        else:
            assert abs(g.total_prob() - c.total_prob()) < 1e-12
    a, b = np.asarray(g.get_state()), c.get_state()
    err = float(np.max(np.abs(a - b))) / max(1e-300, float(np.max(np.abs(b))))
    assert err <= 1e-11, ("state", err)
    g.close()
