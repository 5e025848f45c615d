"""CPU checks of the integer formulation behind the tensor-core SYRK (oracle/ozaki_oracle.py restates
csrc/syrk_i8.cu's arithmetic): digit identity, int32 exactness and the error bound vs float64."""
import numpy as np
import pytest

from oracle import ozaki_oracle as oz


def _Z(K, n, seed):
    rng = np.random.default_rng(seed)
    Z = rng.normal(size=(K, n)) * np.exp(rng.uniform(-8, 8, size=(1, n)))
    Z[rng.uniform(size=Z.shape) < 0.3] = 0.0
    Z[:, -2:] = 0.0
    return Z


@pytest.mark.parametrize("s", [3, 5, 7])
def test_balanced_digits_reconstruct_the_rounded_value(s):
    Z = _Z(200, 17, s)
    Z[0, 0] = np.abs(Z[:, 0]).max() * (1 - 2.0 ** -52)        # an entry just below the column scale
    D, e, X = oz.slices(Z, s)
    assert D.dtype == np.int8 and np.abs(D[0].astype(int)).max() <= 65          # top digit has the headroom
    rec = sum(D[p].astype(np.int64) * (256 ** (s - 1 - p)) for p in range(s))
    assert np.array_equal(rec, X)                                                 # digits are exact
    B = 8 * s - 2
    back = np.ldexp(X.astype(np.float64), (e - B)[None, :])
    scale = np.ldexp(1.0, e)[None, :]
    assert np.all(np.abs(back - Z) <= 2.0 ** -(B + 1) * scale + 1e-300)           # rounding to B fractional bits


@pytest.mark.parametrize("s,tol", [(7, 2.0 ** -44), (6, 2.0 ** -36), (5, 2.0 ** -28), (3, 2.0 ** -12)])
def test_syrk_error_bound(s, tol):
    Z = _Z(1500, 40, 10 + s)
    got = oz.syrk(Z, s)
    ref = Z.T @ Z
    bound = np.abs(Z).T @ np.abs(Z)
    assert np.all(np.abs(got - ref) <= tol * bound + 1e-300), (np.abs(got - ref) / (bound + 1e-300)).max()
    assert np.array_equal(got, got.T)


def test_dropped_orders_are_below_the_rounding():
    """Keeping orders beyond s+1 changes the result by less than the slicing's own rounding error."""
    Z = _Z(800, 24, 3)
    a = oz.syrk(Z, 6)
    b = oz.syrk(Z, 6, max_order=12)
    bound = np.abs(Z).T @ np.abs(Z)
    assert np.all(np.abs(a - b) <= 2.0 ** -36 * bound + 1e-300)


def test_int32_headroom_per_work_item():
    """|d| <= 128, at most 7 pairs per order: a work item of 256 k-blocks x 64 (csrc/syrk_i8.cu OZ_MAX_ITEM_KB) keeps
    every int32 accumulator exact even if all digits sit at -128; longer reductions are split into several items."""
    assert 7 * 128 * 128 * 256 * 64 < 2 ** 31
    assert 7 * 128 * 128 * 12288 < 2 ** 31            # C3 (K = 3 x 4096) fits in one item
