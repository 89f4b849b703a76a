"""Checkpoint -> fixed-point fixtures: the quantise/export step of the reference, restated.

The reference produced its ``*.Weights.txt`` / ``*testData*.txt`` files with notebook cells
(/root/reference/CNN.ipynb:1 cells 23-25: ``float2fix(x, 18, 12)`` printed as
``18'dADDR: data = 18'bBITS;``) and a weight-dump script that is NOT in the repository; its
layout was recovered by exhaustive matching of the five checkpoint / text-file pairs
(pinned by tests/test_formats.py):

    conv table   [3f, 3f+1, 3f+2]      = fix(K[0,0,0,f]), fix(K[0,1,0,f]), fix(bias[f])
    dense table  (class c, row r)[129 f + p] = fix(DenseKernel[(r*129 + p)*F + f, c])
    dense bias   [c]                    = fix(dense_bias[c])

This module goes the same way (h5 -> tables -> text) so that fixtures round-trip, and builds the
10-filter integer model the SystemVerilog reserves ROM addresses for
(cnn_test_latest1.sv:555-599, table map :267-273) from ``convmodrecnets_CNN2_0.5.wts.h5`` +
``DenseWeights1.txt``.

    python -m modulationdetectioncnn_b200.export weights  model.wts.h5  out.Weights.txt
    python -m modulationdetectioncnn_b200.export vector   frame.npy     out.testData.txt
"""
from __future__ import annotations

import sys
from typing import Optional, Sequence

import numpy as np

from . import fixedpoint as fx
from . import svtext
from .svtext import QWeights

__all__ = ["quantize_checkpoint", "quantize_frame", "qweights_from_dense_dump", "export_weights", "export_vector"]


def quantize_checkpoint(weights: Sequence[np.ndarray], overwidth: str = "verilog") -> QWeights:
    """[conv_kernel (1,2,1,F), conv_bias (F), dense_kernel (258F,C), dense_bias (C)] -> QWeights.

    Uses ``float2fix`` semantics including its tiny-negative bug (``overwidth`` says what the
    resulting 19-bit literal means, see fixedpoint.bits_to_int).  The deployed ROMs differ from
    this in a handful of hand-edited entries: for bit-exact hardware parity load
    the literal text files instead.
    """
    ck, cb, dk, db = [np.asarray(w, dtype=np.float32) for w in weights]
    if ck.ndim != 4 or ck.shape[:3] != (1, 2, 1):
        raise ValueError(f"conv kernel must be (1,2,1,F), got {ck.shape}")
    F = ck.shape[-1]
    if cb.shape != (F,) or dk.ndim != 2 or dk.shape[0] != 258 * F or db.shape != (dk.shape[1],):
        raise ValueError("weights are not a TinyCNN2(F,C) checkpoint")
    C = dk.shape[1]
    conv = np.stack([fx.quantize(ck[0, 0, 0], overwidth=overwidth), fx.quantize(ck[0, 1, 0], overwidth=overwidth),
                     fx.quantize(cb, overwidth=overwidth)], axis=1).reshape(-1)
    d4 = dk.reshape(2, 129, F, C)                                   # Keras channels_last flatten: (r, p, f)
    tabs = np.empty((2 * C, 129 * F), dtype=np.int32)
    for c in range(C):
        for r in range(2):
            tabs[2 * c + r] = fx.quantize(d4[r, :, :, c].T.reshape(-1), overwidth=overwidth)   # [f*129 + p]
    return QWeights(conv.astype(np.int32), fx.quantize(db, overwidth=overwidth), tabs, source="quantize_checkpoint").validate()


def qweights_from_dense_dump(weights: Sequence[np.ndarray], dense_flat: np.ndarray,
                             overwidth: str = "verilog") -> QWeights:
    """The 10-filter integer model (weight set E): conv table and dense bias quantised from the
    checkpoint, dense tables taken LITERALLY from a ``DenseWeights1.txt``-style dump laid out
    ``[c][r][f][p]``."""
    q = quantize_checkpoint(weights, overwidth)
    F, C = q.filters, q.classes
    flat = np.asarray(dense_flat, dtype=np.int32)
    if flat.shape != (C * 2 * F * 129,):
        raise ValueError(f"dense dump must have {C * 2 * F * 129} entries, got {flat.shape}")
    return QWeights(q.conv_tab, q.dense_bias, flat.reshape(2 * C, 129 * F).copy(), source="dense dump + checkpoint").validate()


def quantize_frame(x: np.ndarray, overwidth: str = "verilog") -> np.ndarray:
    """float frame(s) (...,2,128) -> int32 (...,256) test vector(s): addresses 0-127 = I, 128-255 = Q
    (CNN.ipynb cell 24 prints the I row, then the Q row)."""
    a = np.asarray(x, dtype=np.float32)
    if a.shape[-2:] != (2, 128):
        raise ValueError(f"frames must end in (2,128), got {a.shape}")
    return fx.quantize(a.reshape(a.shape[:-2] + (256,)), overwidth=overwidth)


def export_weights(h5_path: str, out_path: str, overwidth: str = "verilog") -> QWeights:
    from .model import read_keras_weights
    qw = quantize_checkpoint(read_keras_weights(h5_path), overwidth)
    svtext.write_qweights(qw, out_path)
    return qw


def export_vector(frame: np.ndarray, out_path: str, header: Optional[str] = None) -> np.ndarray:
    v = quantize_frame(frame)
    if v.ndim != 1:
        raise ValueError("export_vector writes one frame")
    svtext.write_vector(v, out_path, header=header)
    return v


def main(argv: Optional[Sequence[str]] = None) -> int:
    a = list(sys.argv[1:] if argv is None else argv)
    if len(a) != 3 or a[0] not in ("weights", "vector"):
        sys.stderr.write(__doc__)
        return 2
    if a[0] == "weights":
        qw = export_weights(a[1], a[2])
        print(f"{a[2]}: F={qw.filters} C={qw.classes} conv={qw.conv_tab.tolist()} dense_bias={qw.dense_bias.tolist()}")
    else:
        v = export_vector(np.load(a[1]), a[2])
        print(f"{a[2]}: 256 entries, range [{v.min()}, {v.max()}]")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
