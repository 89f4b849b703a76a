import numpy as np
import scipy.linalg

from oracle import fwht


def test_hadamard_matches_scipy():
    for n in (1, 2, 32, 256):
        assert np.array_equal(fwht.hadamard(n), scipy.linalg.hadamard(n))


def test_matrix_vs_butterfly_and_involution():
    rng = np.random.default_rng(4)
    for n in (32, 128, 1024, 4096):
        x = rng.integers(-(1 << 17), 1 << 17, (3, n))
        a = fwht.fwht_matrix(x)
        assert np.array_equal(a, fwht.fwht_butterfly(x))
        assert np.array_equal(fwht.fwht_butterfly(a.astype(np.int64)).astype(np.int64), (x * n).astype(np.int32))


def test_wraparound_is_mod_2_32():
    x = np.full((1, 1024), (1 << 31) - 1, dtype=np.int64)
    a = fwht.fwht_matrix(x)
    assert a[0, 0] == np.int32((((1 << 31) - 1) * 1024 + (1 << 31)) % (1 << 32) - (1 << 31))
    assert np.array_equal(a, fwht.fwht_butterfly(x))


def test_sequency_order_counts_sign_changes():
    n = 64
    H = fwht.hadamard(n)
    p = fwht.sequency_permutation(n)
    assert sorted(p.tolist()) == list(range(n))
    assert [(np.diff(H[p[k]]) != 0).sum() for k in range(n)] == list(range(n))
    x = np.random.default_rng(0).integers(-100, 100, (2, n))
    assert np.array_equal(fwht.fwht_matrix(x, "sequency"), fwht.fwht_matrix(x)[:, p])
