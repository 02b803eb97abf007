"""GPU parity for b200_activity_mask_dev == oracle (ActivityMask::from_plane + fill_scales).
(The per-thread variance function is also checked on the CPU by tests/test_activity.py.)"""
import numpy as np
import pytest

from tests import gpu_util as G
from tests.test_activity import oracle_mask

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10), (np.uint16, 12)])
def test_activity_mask_matches_oracle(dtype, bd):
    c = G.ctx()
    rng = np.random.default_rng(bd)
    img = rng.integers(0, 1 << bd, (270, 483)).astype(dtype)
    img[:32, :64] = 5
    var, sc, _ = oracle_mask(img, 16, bd)
    dpl = c.plane_from_host(img, 16)
    n = var.size
    d_var, d_sc = G.dev_empty(4 * n), G.dev_empty(4 * n)
    c.activity_mask_dev(dpl, bd, d_var, d_sc)
    c.synchronize()
    np.testing.assert_array_equal(G.from_dev(d_var, np.uint32)[:n].reshape(var.shape), var)
    np.testing.assert_array_equal(G.from_dev(d_sc, np.uint32)[:n].reshape(sc.shape), sc)
    c.plane_free(dpl)
