"""TEST INFRASTRUCTURE - CPU oracle for the Walsh-Hadamard transform.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.

**Parity unpinned**: the reference contains no FWHT code, fixture or number -
/root/reference/README.md:5 is the only mention ("a spectrogram of the Fast
Walsh-Hadamard Transform ... was plotted on a VGA screen").  The definition is
therefore the textbook one: unnormalised
Walsh-Hadamard transform in Sylvester (natural) order, X = H_N x with
H_1=[1], H_2N=[[H_N,H_N],[H_N,-H_N]], int32 in/out with wrap-around mod 2**32;
optional sequency (Walsh) ordering = natural order permuted by
bit-reversal o Gray-code.

``fwht_matrix`` is independent of any butterfly (dense Hadamard matmul) and is
the parity oracle; ``fwht_butterfly`` is the fast CPU baseline.
"""
from __future__ import annotations

import numpy as np

__all__ = ["hadamard", "fwht_matrix", "fwht_butterfly", "sequency_permutation"]


def hadamard(n: int) -> np.ndarray:
    """Sylvester Hadamard matrix: H[i,j] = (-1)**popcount(i & j), int64 [n,n]."""
    if n < 1 or n & (n - 1):
        raise ValueError("n must be a power of two")
    i = np.arange(n)
    a = i[:, None] & i[None, :]
    pc = np.zeros_like(a)
    while a.any():
        pc += a & 1
        a >>= 1
    return (1 - 2 * (pc & 1)).astype(np.int64)


def sequency_permutation(n: int) -> np.ndarray:
    """perm such that X_sequency[k] = X_natural[perm[k]] (k = number of sign changes)."""
    bits = n.bit_length() - 1
    k = np.arange(n)
    g = k ^ (k >> 1)                      # Gray code
    rev = np.zeros_like(g)
    for b in range(bits):                 # bit reversal
        rev |= ((g >> b) & 1) << (bits - 1 - b)
    return rev


def _wrap32(v: np.ndarray) -> np.ndarray:
    return ((v + (1 << 31)) % (1 << 32) - (1 << 31)).astype(np.int32)


def fwht_matrix(x: np.ndarray, ordering: str = "natural") -> np.ndarray:
    """x int [S,N] -> int32 [S,N] via dense H_N matmul in exact integer arithmetic."""
    x = np.asarray(x)
    n = x.shape[-1]
    H = hadamard(n)
    # |x| < 2**31, n <= 2**20  ->  |sum| < 2**51: exact in int64
    y = x.astype(np.int64).reshape(-1, n) @ H.T
    y = _wrap32(y).reshape(x.shape)
    if ordering == "sequency":
        y = y[..., sequency_permutation(n)]
    elif ordering != "natural":
        raise ValueError(ordering)
    return y


def fwht_butterfly(x: np.ndarray, ordering: str = "natural") -> np.ndarray:
    """In-place radix-2 butterflies, vectorised over spectra (CPU baseline)."""
    x = np.asarray(x)
    n = x.shape[-1]
    y = x.astype(np.int64).reshape(-1, n).copy()
    h = 1
    while h < n:
        v = y.reshape(-1, n // (2 * h), 2, h)
        a = v[:, :, 0, :] + v[:, :, 1, :]
        b = v[:, :, 0, :] - v[:, :, 1, :]
        v[:, :, 0, :] = a
        v[:, :, 1, :] = b
        h *= 2
    y = _wrap32(y).reshape(x.shape)
    if ordering == "sequency":
        y = y[..., sequency_permutation(n)]
    elif ordering != "natural":
        raise ValueError(ordering)
    return y
