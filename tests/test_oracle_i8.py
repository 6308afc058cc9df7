"""CPU tests of the fixed-point / digit-plane scheme of the INT8 tensor-core kernels (oracle/i8_oracle.py restates every
kernel step in numpy): exactness of the digits, of the integer GEMMs and of the recombination, and the error bound
against a long-double product."""
from fractions import Fraction

import numpy as np
import pytest

from oracle.i8_oracle import (NPL, balanced_digits, column_scales, combine7, exact_fixed_sum, mask_contract_i8, to_fixed)


def _case(N, D, q, seed, spread=False):
    rng = np.random.RandomState(seed)
    W, Wv = rng.randn(D, q), rng.rand(D, q) + 0.5
    if spread:
        W[:, 0] *= 1e3
        W[:, 1] *= 1e-4
    ii, jj = np.tril_indices(q)
    G = W[:, ii] * W[:, jj] + np.where(ii == jj, Wv[:, ii], 0.0)       # node.py:219-224, packed lower triangle
    mask = (rng.rand(N, D) > 0.3)
    mask[0] = False
    mask[1] = True
    return mask, G


def test_digits_are_balanced_and_reconstruct_the_integer():
    rng = np.random.RandomState(0)
    v = np.concatenate([rng.randint(-2 ** 54, 2 ** 54, size=4000, dtype=np.int64),
                        np.array([0, 1, -1, 127, 128, -128, -129, 2 ** 54, -2 ** 54, 2 ** 54 - 1], dtype=np.int64)])
    d = balanced_digits(v)
    assert d.dtype == np.int8 and d.shape == (NPL, v.size)
    back = sum(int(256 ** t) * d[t].astype(object) for t in range(NPL))
    assert all(int(b) == int(x) for b, x in zip(back, v))


def test_fixed_point_keeps_every_entry_to_2_pow_minus_55_of_the_column_maximum():
    _, G = _case(4, 96, 8, seed=1, spread=True)
    s = column_scales(G)
    v = to_fixed(G, s)
    assert np.all(np.abs(v) < 2 ** 54) and np.all(np.log2(s) == np.round(np.log2(s)))       # power-of-two scales
    # exact comparison in rational arithmetic: every entry within scale 2^-55, the large ones exactly
    for d in range(G.shape[0]):
        for c in range(G.shape[1]):
            fx = Fraction(int(v[d, c])) * Fraction(float(s[c])) / 2 ** 54
            err = abs(fx - Fraction(float(G[d, c])))
            assert err <= Fraction(float(s[c])) / 2 ** 55
            if abs(G[d, c]) >= s[c] / 4:
                assert err == 0


@pytest.mark.parametrize("shape", [(7, 64, 4), (5, 256, 6)])
def test_integer_gemms_and_recombination_are_exact_up_to_one_rounding(shape):
    N, D, q = shape
    mask, G = _case(N, D, q, seed=D, spread=True)
    exact, scale = exact_fixed_sum(mask, G)                            # Python integers
    dig = balanced_digits(to_fixed(G, scale))
    acc = np.stack([mask.astype(np.int32) @ dig[t].astype(np.int32) for t in range(NPL)])
    got = combine7(acc)
    for n in range(N):
        for c in range(G.shape[1]):
            e = exact[n][c]
            assert sum(int(acc[t][n][c]) * 256 ** t for t in range(NPL)) == e          # the seven GEMMs hold the exact sum
            assert got[n][c] == float(e) or abs(Fraction(got[n][c]) - e) <= Fraction(abs(e), 2 ** 53)   # one rounding


@pytest.mark.parametrize("shape", [(40, 128, 8), (12, 1024, 5)])
def test_result_is_within_the_fixed_point_bound_of_the_true_product(shape):
    N, D, q = shape
    mask, G = _case(N, D, q, seed=N, spread=True)
    tau = 3.7
    got = mask_contract_i8(mask, G, tau=tau)
    ref = tau * (mask.astype(np.longdouble) @ G.astype(np.longdouble))
    scale = column_scales(G)
    bound = tau * mask.sum(1)[:, None] * scale[None, :] * 2.0 ** -55 + 4 * np.finfo(np.float64).eps * np.abs(ref)
    assert np.all(np.abs(got - ref) <= bound)
    # and it is at least as close to the truth as a float64 matmul is allowed to be (D rounding errors)
    f64 = tau * (mask.astype(np.float64) @ G)
    assert np.max(np.abs(got - ref)) <= np.max(np.abs(f64 - ref)) + np.max(bound)
    assert np.all(got[0] == 0.0)                                        # an all-missing row contributes nothing


@pytest.mark.parametrize("s", [1.0, 30.0, 1e3, 1e4, 1e6])
def test_guard_flags_every_row_whose_fixed_point_error_matters(s):
    """One data dimension of W scaled by s (an un-normalised feature): rows that do not observe it keep an absolute error of
    ~ s^2 2^-55 in a qprec of size O(D).  The guard (restated in i8_oracle.zstep_guard_rows) must flag every row whose error
    exceeds 1e-10 of its largest diagonal entry, and must leave normalised data (s = 1) alone."""
    from oracle.i8_oracle import zstep_guard_rows
    rng = np.random.RandomState(5)
    N, D, q, tau = 300, 128, 6, 20.0
    W, Wv = rng.randn(D, q), np.full((D, q), 1e-3)
    W[7] *= s
    ii, jj = np.tril_indices(q)
    G = W[:, ii] * W[:, jj] + np.where(ii == jj, Wv[:, ii], 0.0)
    mask = rng.rand(N, D) > 0.3
    P0 = np.eye(q)[ii, jj][None, :]
    got = mask_contract_i8(mask, G, tau=tau, add=P0)
    ref = (P0 + tau * (mask.astype(np.longdouble) @ G.astype(np.longdouble))).astype(np.float64)
    dmax = ref[:, ii == jj].max(1)
    err = np.max(np.abs(got - ref), axis=1) / dmax                 # per row, max norm
    flagged = zstep_guard_rows(dmax, G, tau, D)
    assert np.all(flagged | (err <= 1e-10)), (s, float(err[~flagged].max()))
    if s <= 30.0:
        assert not flagged.any()                                     # moderate ranges stay on the INT8 path
    if s >= 1e4:
        assert flagged.any() and err.max() > (1e-9 if s >= 1e5 else 1e-10)   # the hazard is real: unguarded it breaks 1e-9 per row


def test_stats_guard_flags_dimensions_hidden_from_outlier_rows():
    from oracle.i8_oracle import stats_guard_dims
    rng = np.random.RandomState(2)
    N, D, q = 4000, 32, 4
    Z = rng.randn(N, q)
    Z[:5] *= 1e5                                                     # outlier rows
    ii, jj = np.tril_indices(q)
    M2 = Z[:, ii] * Z[:, jj] + np.where(ii == jj, 0.1, 0.0)
    mask = rng.rand(N, D) > 0.3
    mask[:5, :8] = False                                             # dimensions 0..7 do not see the outliers
    got = mask_contract_i8(mask.T, M2)
    ref = (mask.T.astype(np.longdouble) @ M2.astype(np.longdouble)).astype(np.float64)
    dmax = ref[:, ii == jj].max(1)
    err = np.max(np.abs(got - ref), axis=1) / dmax
    flagged = stats_guard_dims(dmax, mask.sum(0), M2)
    assert np.all(flagged | (err <= 1e-10))
    assert flagged[:8].all() and not flagged[8:].any()
