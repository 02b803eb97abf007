"""GPU parity for compute_rd_cost (rdo.rs:718-723): CUDA DFMA == oracle fma() == the exactly rounded
value, bit for bit (tolerance 0 ULP; the north star allows 1), plus the fused first-minimum."""
import numpy as np
import pytest

from tests import gpu_util as G
from tests.test_oracle_rd_cost import L, cases, exact

pytestmark = pytest.mark.gpu


def test_rd_cost_batch_is_bit_exact():
    c = G.ctx()
    l = L()
    _, rate, dist = cases(20000, 5)
    for lam in (0.37, 117.25, 3.0e-7, 8.1e9):
        want = np.zeros(len(rate))
        l.orc_compute_rd_cost_batch(lam, rate.ctypes.data, dist.ctypes.data, len(rate), want.ctypes.data)
        d_cost = G.dev_empty(8 * len(rate))
        c.compute_rd_cost_dev(lam, G.to_dev(rate), G.to_dev(dist), len(rate), d_cost)
        c.synchronize()
        got = G.from_dev(d_cost, np.float64)[:len(rate)]
        assert (got.view(np.uint64) == want.view(np.uint64)).all()           # 0 ULP
        assert all(got[i] == exact(lam, rate[i], dist[i]) for i in range(0, len(rate), 97))


def test_first_minimum_per_group_matches_the_serial_scan():
    c = G.ctx()
    rng = np.random.default_rng(9)
    ngroups = 700
    counts = rng.integers(0, 70, ngroups)
    offs = np.zeros(ngroups + 1, np.uint32)
    offs[1:] = np.cumsum(counts)
    n = int(offs[-1])
    rate = rng.integers(0, 50, n).astype(np.uint32) * 8          # few distinct costs: many exact ties
    dist = rng.integers(0, 6, n).astype(np.uint64) * 100
    lam = 12.5
    l = L()
    cost = np.zeros(n)
    l.orc_compute_rd_cost_batch(lam, rate.ctypes.data, dist.ctypes.data, n, cost.ctypes.data)
    want = np.full(ngroups, 0xffffffff, np.uint32)
    for g in range(ngroups):
        best = None
        for i in range(offs[g], offs[g + 1]):
            if best is None or cost[i] < cost[best]:           # `if rd < best.rd_cost`
                best = i
        if best is not None:
            want[g] = best - offs[g]
    d_cost, d_best = G.dev_empty(8 * n), G.dev_empty(4 * ngroups)
    c.compute_rd_cost_dev(lam, G.to_dev(rate), G.to_dev(dist), n, d_cost, G.to_dev(offs), ngroups, d_best)
    c.synchronize()
    np.testing.assert_array_equal(G.from_dev(d_best, np.uint32)[:ngroups], want)
    assert (G.from_dev(d_cost, np.float64)[:n].view(np.uint64) == cost.view(np.uint64)).all()


def test_per_call_form():
    from rav1e_b200 import backend as B
    lib = B.lib()
    for lam, r, d in ((1.0, 8, 5), (0.1, 12345, 1 << 40), (977.3, 0, (1 << 64) - 1)):
        assert lib.b200_compute_rd_cost(lam, r, d) == exact(lam, r, d)
