"""Integer (SystemVerilog-exact) CNN2: the fixed-point datapath behind a model-like facade.

Stands in for one reset-to-done run of ``layers_top`` per frame
(/root/reference/cnn_test_latest1.sv:144-209) fed by ``test_input``/``test_table``
(:71-142).  Weights are the literal ROM/table contents of the ``*.Weights.txt`` files
(never re-quantised checkpoints: the deployed ROMs contain hand edits).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib, svtext
from .svtext import QWeights

__all__ = ["FixedPointCNN2"]


def _host_frames(x):
    """(array [N, row bytes], MDC_IN_* format) for a host batch: uint8 -> raw interleaved I/Q bytes [N,128,2]
    (sample = (2u - 255) * 16, what sdr.ingest_u8 writes); int16 -> Q6.12 samples [N,256]; anything else -> the int32
    words of test_table [N,256]."""
    a = np.asarray(x)
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a).reshape(-1, 256), _lib.IN_U8IQ
    if a.dtype == np.int16:
        return np.ascontiguousarray(a).reshape(-1, 256), _lib.IN_I16
    return np.ascontiguousarray(a, dtype=np.int32).reshape(-1, 256), _lib.IN_I32


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


class FixedPointCNN2:
    def __init__(self, filters: int = 3, classes: int = 3, device: Optional[int] = None):
        from .model import default_device
        device = default_device() if device is None else int(device)
        self.filters, self.classes, self.device = filters, classes, device
        self._h = _lib.Handle(_lib.MODEL_TINY, filters, classes, _lib.MODE_Q612, device)
        self.tables: Optional[QWeights] = None

    # ---- weights
    def set_tables(self, qw: QWeights) -> None:
        qw.validate()
        if qw.filters != self.filters or qw.classes != self.classes:
            raise ValueError(f"tables are F={qw.filters},C={qw.classes}; model is F={self.filters},C={self.classes}")
        ct = np.ascontiguousarray(qw.conv_tab, dtype=np.int32)
        db = np.ascontiguousarray(qw.dense_bias, dtype=np.int32)
        dt = np.ascontiguousarray(qw.dense_tabs, dtype=np.int32)
        _lib.check(self._h._lib.mdc_set_weights_q612(self._h.ptr, ct.ctypes.data, db.ctypes.data, dt.ctypes.data))
        self.tables = qw

    def load_weights(self, filepath: str, *, conv_from: Optional[str] = None,
                     dense_bias: Optional[Sequence[int]] = None, overwidth: str = "verilog") -> None:
        """Load a ``*.Weights.txt`` file (see :func:`svtext.load_qweights`)."""
        self.set_tables(svtext.load_qweights(filepath, conv_from=conv_from, dense_bias=dense_bias,
                                             classes=self.classes, overwidth=overwidth))

    # ---- inference
    def _run(self, x, want: Sequence[str]) -> Dict[str, object]:
        lib, Cn = self._h._lib, self.classes
        if _is_torch(x):
            import torch
            if not x.is_cuda:
                raise ValueError("torch inputs must be CUDA tensors (pass numpy for the host path)")
            if x.device.index != self.device:
                raise ValueError(f"input lives on cuda:{x.device.index}, this model on cuda:{self.device}")
            if x.dtype == torch.uint8:
                xt, fmt = x.reshape(-1, 256).contiguous(), _lib.IN_U8IQ
            elif x.dtype == torch.int16:
                xt, fmt = x.reshape(-1, 256).contiguous(), _lib.IN_I16
            else:
                xt, fmt = x.reshape(-1, 256).to(torch.int32).contiguous(), _lib.IN_I32
            n = xt.shape[0]
            out: Dict[str, object] = {}
            with torch.cuda.device(xt.device):
                for k in want:
                    if k == "hist":
                        out[k] = torch.zeros((Cn,), dtype=torch.int64, device=xt.device)
                    elif k == "cls":
                        out[k] = torch.empty((n,), dtype=torch.int32, device=xt.device)
                    else:
                        out[k] = torch.empty((n, Cn), dtype=torch.int32, device=xt.device)
                ptr = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
                _lib.check(lib.mdc_predict_q612_raw(self._h.ptr, xt.data_ptr(), fmt, n, ptr("out"), ptr("pre"), ptr("cls"),
                                                    ptr("hist"), torch.cuda.current_stream(xt.device).cuda_stream))
            return out
        xa, fmt = _host_frames(x)
        n = xa.shape[0]
        out = {}
        for k in want:
            if k == "hist":
                out[k] = np.zeros((Cn,), dtype=np.uint64)
            elif k == "cls":
                out[k] = np.empty((n,), dtype=np.int32)
            else:
                out[k] = np.empty((n, Cn), dtype=np.int32)
        ptr = lambda k: out[k].ctypes.data if k in out else None  # noqa: E731
        _lib.check(lib.mdc_predict_q612_raw_host(self._h.ptr, xa.ctypes.data, fmt, n, ptr("out"), ptr("pre"), ptr("cls"),
                                                 ptr("hist")))
        if "hist" in out:
            out["hist"] = out["hist"].astype(np.int64)
        return out

    def predict(self, x, output: str = "out"):
        """x int32 [N,256] (0-127 I, 128-255 Q; values are wrapped to 18 bits like a sized
        Verilog literal), int16 [N,256] (the same address map, Q6.12) or uint8 [N,128,2] (raw RTL-SDR bytes
        I0 Q0 I1 Q1 ..., sample = (2u - 255) * 16 as ``sdr.ingest_u8`` writes it).  ``output``: "out" (``out_data``, ReLU'd Q.12 int32 [N,C]),
        "pre" (``pre_out_data``), "argmax"."""
        key = {"out": "out", "pre": "pre", "argmax": "cls"}.get(output)
        if key is None:
            raise ValueError("output must be 'out', 'pre' or 'argmax'")
        return self._run(x, [key])[key]

    def predict_async(self, x, output: str = "out"):
        """Streaming form of :meth:`predict` for host (numpy) batches, see ``CNN2Model.predict_async``."""
        from .model import PendingPrediction, pinned_empty
        import ctypes as C
        key = {"out": "out", "pre": "pre", "argmax": "cls"}.get(output)
        if key is None:
            raise ValueError("output must be 'out', 'pre' or 'argmax'")
        xa, fmt = _host_frames(x)
        n = xa.shape[0]
        out = pinned_empty((n,), np.int32) if key == "cls" else pinned_empty((n, self.classes), np.int32)
        ptr = lambda k: out.ctypes.data if k == key else None  # noqa: E731
        ticket = C.c_int64(0)
        _lib.check(self._h._lib.mdc_predict_q612_raw_host_async(self._h.ptr, xa.ctypes.data, fmt, n, ptr("out"), ptr("pre"),
                                                                ptr("cls"), None, C.byref(ticket)))
        return PendingPrediction(self, ticket.value, out, xa)

    def class_histogram(self, x):
        return self._run(x, ["hist"])["hist"]

    def predict_file(self, path: str, overwidth: str = "verilog") -> np.ndarray:
        """Run every 256-entry vector of a ``*testData*.txt`` file."""
        return self.predict(svtext.load_vectors(path, overwidth))

    def launch_count(self) -> int:
        return self._h.launch_count()

    def close(self) -> None:
        self._h.close()
