"""compute_rd_cost (rdo.rs:718-723): the oracle's fma() restatement pinned by exact rational
arithmetic — f64::mul_add rounds the exact value lambda * (rate / 8) + (distortion as f64) ONCE to
nearest even, which is what float(Fraction) computes."""
import ctypes as C
from fractions import Fraction

import numpy as np

from tests import oracle_lib as O


def L():
    l = O.lib()
    l.orc_compute_rd_cost.restype = C.c_double
    l.orc_compute_rd_cost.argtypes = [C.c_double, C.c_uint32, C.c_uint64]
    l.orc_compute_rd_cost_batch.restype = None
    l.orc_compute_rd_cost_batch.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    return l


def exact(lam, rate, dist):
    """one rounding of the exact value; `distortion.0 as f64` is itself a rounded conversion"""
    return float(Fraction(lam) * Fraction(int(rate), 8) + Fraction(float(int(dist))))


def cases(n, seed):
    rng = np.random.default_rng(seed)
    lam = np.concatenate([rng.uniform(0.01, 5000.0, n // 2), np.exp(rng.uniform(-20, 20, n - n // 2))])
    rate = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    dist = rng.integers(0, 1 << 63, n, dtype=np.uint64)
    dist[::3] >>= rng.integers(20, 60, len(dist[::3])).astype(np.uint64)   # realistic magnitudes too
    dist[:4] = [0, 1, (1 << 53) + 1, (1 << 64) - 1]                        # u64 -> f64 rounding edges
    rate[:4] = [0, 1, 7, (1 << 32) - 1]
    return lam, rate, dist


def test_oracle_rd_cost_is_the_single_rounded_fma():
    l = L()
    lam, rate, dist = cases(4000, 1)
    for a, r, d in zip(lam, rate, dist):
        got = l.orc_compute_rd_cost(float(a), int(r), int(d))
        assert got == exact(float(a), r, d), (a, r, d)


def test_a_separate_multiply_and_add_differs_somewhere():
    """the test above can tell a fused from an unfused evaluation (otherwise it pins nothing)"""
    lam, rate, dist = cases(4000, 2)
    unfused = lam * (rate.astype(np.float64) / 8.0) + dist.astype(np.float64)
    want = np.array([exact(float(a), r, d) for a, r, d in zip(lam, rate, dist)])
    assert (unfused != want).any()


def test_batch_form():
    l = L()
    _, rate, dist = cases(1000, 3)
    out = np.zeros(len(rate))
    l.orc_compute_rd_cost_batch(123.456, rate.ctypes.data, dist.ctypes.data, len(rate), out.ctypes.data)
    assert all(out[i] == exact(123.456, rate[i], dist[i]) for i in range(len(rate)))
