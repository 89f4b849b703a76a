"""Q6.12 quantiser of the reference, restated (including its bug).

Follows ``float2fix(val, width, precision)`` in /root/reference/CNN.ipynb:1
(cell 23), used there as ``float2fix(x, 18, 12)`` (cells 24, 25) to print the
``18'b...`` literals that were pasted into ``cnn_test_latest1.sv``.

Semantics that matter for parity:

* ``int(val * 2**precision)`` truncates toward zero;
* negative values are emitted as ``'1' + bin(2**(width-1) - |int|)`` which is
  correct two's complement *unless* ``|int| == 0`` (i.e. ``-2**-p < val < 0``):
  then the magnitude field needs ``width`` bits and the literal comes out one
  bit too long (``1100000000000000000`` for width 18).  A sized Verilog literal
  keeps the low ``width`` bits, which makes that value ``-2**(width-1)``.
"""
from __future__ import annotations

from typing import Iterable, Union

import numpy as np

WIDTH = 18
PRECISION = 12

__all__ = [
    "WIDTH", "PRECISION", "float2fix", "bits_to_int", "int_to_bits", "quantize",
    "fix2float", "wrap_signed",
]


def float2fix(val: float, width: int = WIDTH, precision: int = PRECISION) -> str:
    """Bit-string literal exactly as the reference prints it (bug included)."""
    integer = abs(int(val * 2 ** precision))
    if val < 0:
        integer = 2 ** (width - 1) - integer
        return "1" + format(integer, "b").rjust(width - 1, "0")
    return format(integer, "b").rjust(width, "0")


def wrap_signed(v, width: int = WIDTH):
    """Wrap integer(s) to signed two's complement of ``width`` bits."""
    if isinstance(v, np.ndarray):
        m = np.int64(1) << width
        v = v.astype(np.int64) & (m - 1)
        return np.where(v >= (m >> 1), v - m, v)
    m = 1 << width
    v &= m - 1
    return v - m if v >= (m >> 1) else v


def bits_to_int(bits: str, width: int = WIDTH, overwidth: str = "verilog") -> int:
    """Value of an ``N'b<bits>`` literal as signed ``width``-bit integer.

    ``overwidth`` decides what a literal longer than ``width`` means:

    * ``"verilog"`` (default) - sized-literal semantics, keep the low ``width``
      bits (what a simulator does with the 19-bit strings the quantiser bug
      produces): ``1100000000000000000`` -> ``-131072``;
    * ``"zero"`` - treat it as what the quantiser *meant* (a tiny negative
      number that truncates to 0);
    * ``"error"`` - raise ``ValueError``.
    """
    if len(bits) > width:
        if overwidth == "zero":
            return 0
        if overwidth == "error":
            raise ValueError(f"literal {bits!r} is wider than {width} bits")
        if overwidth != "verilog":
            raise ValueError(f"unknown overwidth policy {overwidth!r}")
        bits = bits[-width:]
    # Verilog zero-extends short unsized-sign literals
    return wrap_signed(int(bits, 2), width)


def int_to_bits(v: int, width: int = WIDTH) -> str:
    return format(int(v) & ((1 << width) - 1), "b").rjust(width, "0")


def quantize(x: Union[float, Iterable[float], np.ndarray], width: int = WIDTH,
             precision: int = PRECISION, overwidth: str = "verilog") -> np.ndarray:
    """Vectorised ``bits_to_int(float2fix(x))`` -> int32 array.

    Computed in float64 exactly as Python does for ``val * 2**precision`` when
    ``val`` is a numpy float32/float64 scalar promoted to double.
    """
    a = np.asarray(x)
    scaled = a.astype(np.float64) * float(2 ** precision)
    q = np.trunc(scaled).astype(np.int64)
    tiny_neg = (a < 0) & (q == 0)
    q = wrap_signed(q, width)
    if overwidth == "verilog":
        q = np.where(tiny_neg, -(1 << (width - 1)), q)
    elif overwidth == "error":
        if tiny_neg.any():
            raise ValueError("value in (-2**-precision, 0) hits the float2fix over-width bug")
    elif overwidth != "zero":
        raise ValueError(f"unknown overwidth policy {overwidth!r}")
    return q.astype(np.int32)


def fix2float(q, precision: int = PRECISION) -> np.ndarray:
    return np.asarray(q, dtype=np.float64) / float(2 ** precision)
