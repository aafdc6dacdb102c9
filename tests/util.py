"""Shared helpers for the GPU parity tests."""
import numpy as np

RTOL = 1e-12   # north_star: predictions within 1e-12 relative in fp64


def same_mean(got, exp):
    """Forest means: within the north_star tolerance, and -- because the engine adds leaf values
    in sequential tree order like ranger and the oracle -- bit-equal."""
    got, exp = np.asarray(got), np.asarray(exp)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    ok = ~np.isnan(exp)
    np.testing.assert_allclose(got[ok], exp[ok], rtol=RTOL, atol=0)
    np.testing.assert_array_equal(got[ok], exp[ok])
