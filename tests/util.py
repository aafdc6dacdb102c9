"""Shared helpers for the GPU parity tests."""
import numpy as np
import pytest

# set by the `sum_mode` fixture in conftest.py
EXACT = {"on": False}
RTOL = 1e-12   # north_star: predictions within 1e-12 relative in fp64


def same_mean(got, exp):
    """Forest means: within the north_star tolerance always; bit-equal in the default mode
    (sequential tree-order sum).  The opt-in slot-refill walk (RFSI_B200_REFILL=1) adds
    per-slot partial sums, which moves the result by a few ulp."""
    got, exp = np.asarray(got), np.asarray(exp)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    ok = ~np.isnan(exp)
    np.testing.assert_allclose(got[ok], exp[ok], rtol=RTOL, atol=0)
    if EXACT["on"]:
        np.testing.assert_array_equal(got[ok], exp[ok])


@pytest.fixture(autouse=True, params=["exact", "refill"])
def sum_mode(request, monkeypatch):
    """Import into a test module to run its tests in both traversal modes: the default
    (sequential tree-order sum, bit-equal to the oracle) and RFSI_B200_REFILL=1 (per-warp
    slot refill, <= 1e-12 rel)."""
    EXACT["on"] = request.param == "exact"
    monkeypatch.setenv("RFSI_B200_REFILL", "0" if request.param == "exact" else "1")
    yield
    EXACT["on"] = False
