"""Keras-shaped inference facade over libmdc.so.

Mirrors the call surface the reference uses on its ``Sequential`` model:

* ``model.load_weights(filepath)``            /root/reference/cnn.py:147, CNN.ipynb cell 8
* ``model.predict(X, batch_size=...)``        cnn.py:198,237; CNN.ipynb cells 12,17,18
* ``model.evaluate(X, Y, batch_size, verbose)`` cnn.py:154; CNN.ipynb cell 9 (loss only:
  the model is compiled with ``loss='categorical_crossentropy'`` and no metrics, cnn.py:113)
* sub-models that stop at an inner layer (``Model(inputs, model.layers[4].output)``,
  CNN.ipynb cell 17) -> ``predict(..., output="dense")``

Batches are independent in Keras ``predict``; ``batch_size`` only changes the
chunking and never the result, so it is accepted and ignored (the library picks
its own chunk size for copy/compute overlap).

Inputs: ``numpy.ndarray`` (host path: H2D/D2H handled by the library) or a CUDA
``torch.Tensor`` (device path on torch's current stream; returns torch tensors).
float32 ``(N,2,128)``, row 0 = I, row 1 = Q.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from .h5lite import H5File

__all__ = ["CNN2Model", "PendingPrediction", "tiny_cnn2", "vt_cnn2", "load_model", "read_keras_weights"]

_MODES = {"fp32": _lib.MODE_FP32, "bf16": _lib.MODE_BF16, "tf32x3": _lib.MODE_TF32X3, "f16x3": _lib.MODE_F16X3}


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def default_device() -> int:
    """The CUDA ordinal a model lands on when none is given: torch's current device if torch has initialised CUDA,
    else LOCAL_RANK (torchrun: one rank per GPU), else 0."""
    import os
    import sys
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
        return int(torch.cuda.current_device())
    return int(os.environ.get("LOCAL_RANK", 0))


def pinned_empty(shape, dtype) -> np.ndarray:
    """Page-locked host array (device-to-host copies into it are truly asynchronous); the array keeps its torch
    storage alive."""
    import torch
    t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    return t.numpy()


def _frames_of(x):
    """(array [N, row bytes], MDC_IN_* format) for a host batch: float -> f32 [N,2,128]; uint8 -> raw interleaved
    I/Q bytes [N,128,2]; int16 -> Q6.12 samples [N,2,128] (the test_table address map)."""
    a = np.asarray(x)
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a).reshape(-1, 256), _lib.IN_U8IQ
    if a.dtype == np.int16:
        return np.ascontiguousarray(a).reshape(-1, 256), _lib.IN_I16
    return np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 256), _lib.IN_F32


def read_keras_weights(filepath: str) -> List[np.ndarray]:
    """Weights of a Keras 2.x ``.h5`` in topology order (what ``load_weights`` consumes).

    Handles both full-model files (``/model_weights/...``, as written by
    ``ModelCheckpoint``, cnn.py:143) and weights-only files (layers at the root).
    """
    f = H5File(filepath)
    root = "/model_weights" if "model_weights" in f.listdir("/") else "/"
    out: List[np.ndarray] = []
    for layer in f.attrs(root)["layer_names"]:
        layer = str(layer)
        g = f"{root.rstrip('/')}/{layer}"
        for wn in f.attrs(g).get("weight_names", []):
            out.append(np.asarray(f.dataset(f"{g}/{wn}"), dtype=np.float32))
    return out


def _config_of(filepath: str) -> Optional[dict]:
    a = H5File(filepath).attrs("/")
    return json.loads(a["model_config"]) if "model_config" in a else None


class CNN2Model:
    """TinyCNN2(F,C) or VT-CNN2(C) behind ``load_weights / predict / evaluate``."""

    def __init__(self, kind: str, filters: int = 3, classes: int = 3, mode: str = "fp32",
                 device: Optional[int] = None, flatten: str = "channels_last"):
        if kind not in ("tiny", "vt"):
            raise ValueError("kind must be 'tiny' or 'vt'")
        if mode not in _MODES:
            raise ValueError(f"mode must be one of {sorted(_MODES)}")
        if kind == "tiny" and mode != "fp32":
            raise ValueError("TinyCNN2 float inference is fp32 (integer mode: FixedPointCNN2)")
        device = default_device() if device is None else int(device)
        self.kind, self.filters, self.classes, self.mode, self.device = kind, filters, classes, mode, device
        self.flatten = flatten
        # f16x3: verify after every call that no value left the fp16 range and recompute in tf32x3 if one did
        # (the check synchronises; set False to keep CUDA-tensor calls asynchronous and poll range_ok() yourself)
        self.check_range = True
        self._fallback: Optional["CNN2Model"] = None
        self._weights: List[np.ndarray] = []
        self._h = _lib.Handle(_lib.MODEL_TINY if kind == "tiny" else _lib.MODEL_VT, filters, classes,
                              _MODES[mode], device)
        if kind == "vt":
            if flatten not in ("channels_last", "channels_first"):
                raise ValueError("flatten must be channels_last or channels_first")
            _lib.check(self._h._lib.mdc_set_option(self._h.ptr, _lib.OPT_FLATTEN_ORDER,
                                                   int(flatten == "channels_first")))

    # ------------------------------------------------------------------ weights
    def weight_shapes(self) -> List[tuple]:
        F, Cn = self.filters, self.classes
        if self.kind == "tiny":
            return [(1, 2, 1, F), (F,), (2 * 129 * F, Cn), (Cn,)]
        return [(1, 3, 1, 256), (256,), (2, 3, 256, 80), (80,), (10560, 256), (256,), (256, Cn), (Cn,)]

    def set_weights(self, weights: Sequence[np.ndarray]) -> None:
        """Keras ``model.set_weights``: arrays in topology order, Keras layouts."""
        shapes = self.weight_shapes()
        if len(weights) != len(shapes):
            raise ValueError(f"expected {len(shapes)} weight arrays, got {len(weights)}")
        ids = ([_lib.T_CONV1_K, _lib.T_CONV1_B, _lib.T_DENSE1_K, _lib.T_DENSE1_B] if self.kind == "tiny"
               else list(range(8)))
        ws = []
        for w, shp, tid in zip(weights, shapes, ids):
            a = np.ascontiguousarray(w, dtype=np.float32)
            if a.shape != shp:
                raise ValueError(f"weight {tid}: expected shape {shp}, got {a.shape}")
            _lib.check(self._h._lib.mdc_set_weights_f32(self._h.ptr, tid, a.ctypes.data, a.size))
            ws.append(a)
        self._weights = ws
        if self._fallback is not None:
            self._fallback.set_weights(ws)

    def get_weights(self) -> List[np.ndarray]:
        return [w.copy() for w in self._weights]

    def load_weights(self, filepath: str) -> None:
        self.set_weights(read_keras_weights(filepath))

    # ------------------------------------------------------------------ inference
    def _tf32_twin(self) -> "CNN2Model":
        """f16x3 only: the handle that reruns a batch whose values left the fp16 range (MDC_ERR_RANGE)."""
        if self._fallback is None:
            self._fallback = CNN2Model("vt", 0, self.classes, "tf32x3", self.device, self.flatten)
            self._fallback.set_weights(self._weights)
        return self._fallback

    def range_ok(self) -> bool:
        """f16x3: True if no call since the last check saw a value outside the fp16 range (synchronises)."""
        return self._h.range_flags(reset=True) == 0

    def reserve(self, max_frames: int) -> None:
        """Size the work space (and pack the weights) so that later calls only enqueue kernels (mdc_reserve)."""
        self._h.reserve(max_frames)

    def _run(self, x, want_probs: bool, want_dense: bool, want_cls: bool, want_hist: bool) -> Dict[str, object]:
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
                xt, fmt = x.reshape(-1, 256).to(torch.float32).contiguous(), _lib.IN_F32
            n = xt.shape[0]
            out: Dict[str, object] = {}
            with torch.cuda.device(xt.device):
                if want_probs:
                    out["probs"] = torch.empty((n, Cn), dtype=torch.float32, device=xt.device)
                if want_dense:
                    out["dense"] = torch.empty((n, Cn), dtype=torch.float32, device=xt.device)
                if want_cls:
                    out["cls"] = torch.empty((n,), dtype=torch.int32, device=xt.device)
                if want_hist:
                    out["hist"] = torch.zeros((Cn,), dtype=torch.int64, device=xt.device)
                ptr = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
                _lib.check(lib.mdc_predict_raw(self._h.ptr, xt.data_ptr(), fmt, n, ptr("probs"), ptr("dense"),
                                               ptr("cls"), ptr("hist"),
                                               torch.cuda.current_stream(xt.device).cuda_stream))
                if self.mode == "f16x3" and self.check_range and not self.range_ok():
                    return self._tf32_twin()._run(x, want_probs, want_dense, want_cls, want_hist)
            return out
        xa, fmt = _frames_of(x)
        n = xa.shape[0]
        out = {}
        if want_probs:
            out["probs"] = np.empty((n, Cn), dtype=np.float32)
        if want_dense:
            out["dense"] = np.empty((n, Cn), dtype=np.float32)
        if want_cls:
            out["cls"] = np.empty((n,), dtype=np.int32)
        if want_hist:
            out["hist"] = np.zeros((Cn,), dtype=np.uint64)
        ptr = lambda k: out[k].ctypes.data if k in out else None  # noqa: E731
        rc = lib.mdc_predict_raw_host(self._h.ptr, xa.ctypes.data, fmt, n, ptr("probs"), ptr("dense"),
                                      ptr("cls"), ptr("hist"))
        if rc == _lib.ERR_RANGE and self.mode == "f16x3":
            return self._tf32_twin()._run(x, want_probs, want_dense, want_cls, want_hist)
        _lib.check(rc)
        if want_hist:
            out["hist"] = out["hist"].astype(np.int64)
        return out

    def predict(self, x, batch_size: int = 32, verbose: int = 0, output: str = "softmax"):
        """``output``: "softmax" (the model output), "dense" (last Dense before softmax:
        Dense+ReLU for TinyCNN2 = ``model2`` of CNN.ipynb cell 17, logits for VT-CNN2),
        "argmax" (int32 class ids, what cnn.py:209,244 computes per row)."""
        if output == "softmax":
            return self._run(x, True, False, False, False)["probs"]
        if output == "dense":
            return self._run(x, False, True, False, False)["dense"]
        if output == "argmax":
            return self._run(x, False, False, True, False)["cls"]
        raise ValueError("output must be 'softmax', 'dense' or 'argmax'")

    def predict_async(self, x, output: str = "softmax") -> "PendingPrediction":
        """Streaming form of :meth:`predict` for host (numpy) batches: returns at once, the next batch's transfer
        then runs under this batch's kernels; ``.result()`` waits for and returns this batch's array (page-locked
        memory owned by the result).  ``x`` must stay unchanged until then; pinned input
        (``torch.from_numpy(a).pin_memory().numpy()``) is copied by DMA straight away, pageable input is first staged
        by the calling thread.  Results complete in submission order.  (Keras' ``predict`` has no counterpart: it is
        the call for a stream of batches, e.g. raw uint8 I/Q from an RTL-SDR.)"""
        key = {"softmax": "probs", "dense": "dense", "argmax": "cls"}.get(output)
        if key is None:
            raise ValueError("output must be 'softmax', 'dense' or 'argmax'")
        if _is_torch(x):
            raise ValueError("predict_async takes host (numpy) batches; CUDA tensors are already asynchronous")
        xa, fmt = _frames_of(x)
        n, Cn = xa.shape[0], self.classes
        out = pinned_empty((n,), np.int32) if key == "cls" else pinned_empty((n, Cn), np.float32)
        ptr = lambda k: out.ctypes.data if k == key else None  # noqa: E731
        ticket = C.c_int64(0)
        _lib.check(self._h._lib.mdc_predict_raw_host_async(self._h.ptr, xa.ctypes.data, fmt, n, ptr("probs"), ptr("dense"),
                                                           ptr("cls"), None, C.byref(ticket)))
        return PendingPrediction(self, ticket.value, out, xa, output)

    def predict_classes(self, x):
        return self.predict(x, output="argmax")

    def class_histogram(self, x):
        """int64 [C] count of argmax classes (the fused epilogue; one tiny D2H)."""
        return self._run(x, False, False, False, True)["hist"]

    def evaluate(self, x, y, batch_size: int = 32, verbose: int = 0) -> float:
        """Mean categorical cross-entropy, Keras semantics (probabilities re-normalised and
        clipped to [1e-7, 1-1e-7]) - cnn.py:154, CNN.ipynb cell 9."""
        p = self.predict(x)
        if _is_torch(p):
            p = p.cpu().numpy()
        if _is_torch(y):
            y = y.cpu().numpy()
        p = p.astype(np.float64)
        p = p / p.sum(axis=-1, keepdims=True)
        p = np.clip(p, 1e-7, 1 - 1e-7)
        y = np.asarray(y, dtype=np.float64).reshape(p.shape)
        return float(-(y * np.log(p)).sum(axis=-1).mean())

    # ------------------------------------------------------------------ misc
    def summary(self) -> str:
        if self.kind == "tiny":
            F, Cn = self.filters, self.classes
            rows = [("Reshape", (2, 128, 1), 0), ("ZeroPadding2D", (2, 130, 1), 0),
                    ("Conv2D+ReLU", (2, 129, F), 3 * F), ("Flatten", (258 * F,), 0),
                    ("Dense+ReLU", (Cn,), 258 * F * Cn + Cn), ("Softmax", (Cn,), 0)]
        else:
            Cn = self.classes
            rows = [("Reshape", (1, 2, 128), 0), ("ZeroPadding2D", (1, 2, 132), 0),
                    ("Conv2D 1x3 +ReLU", (256, 2, 130), 1024), ("ZeroPadding2D", (256, 2, 134), 0),
                    ("Conv2D 2x3 +ReLU", (80, 1, 132), 122960), ("Flatten", (10560,), 0),
                    ("Dense+ReLU", (256,), 2703616), ("Dense", (Cn,), 256 * Cn + Cn), ("Softmax", (Cn,), 0)]
        lines = [f"{n:<20}{str(s):<18}{p:>10}" for n, s, p in rows]
        lines.append(f"Total params: {sum(p for _, _, p in rows)}")
        text = "\n".join(lines)
        print(text)
        return text

    def launch_count(self) -> int:
        return self._h.launch_count()

    def close(self) -> None:
        if self._fallback is not None:
            self._fallback.close()
        self._h.close()


class PendingPrediction:
    """Handle of one :meth:`CNN2Model.predict_async` batch."""

    def __init__(self, model, ticket: int, out: np.ndarray, keep_alive, output: str = ""):
        self._model, self._ticket, self._out, self._keep, self._output = model, ticket, out, keep_alive, output

    def done(self) -> bool:
        """True once ``result()`` has been taken."""
        return self._ticket is None

    def result(self) -> np.ndarray:
        if self._ticket is not None:
            rc = self._model._h._lib.mdc_host_wait(self._model._h.ptr, self._ticket)
            if rc == _lib.ERR_RANGE and getattr(self._model, "mode", "") == "f16x3":
                self._out = self._model._tf32_twin().predict(self._keep, output=self._output)
            else:
                _lib.check(rc)
            self._ticket, self._keep = None, None
        return self._out


def tiny_cnn2(filters: int = 3, classes: int = 3, device: Optional[int] = None) -> CNN2Model:
    """The net of CNN.ipynb cell 6 (filters=3) / cnn.py-era checkpoint (filters=10)."""
    return CNN2Model("tiny", filters, classes, "fp32", device)


def vt_cnn2(classes: int = 11, mode: str = "f16x3", device: Optional[int] = None,
            flatten: str = "channels_last") -> CNN2Model:
    """VT-CNN2 of the example notebook (:231-243).  ``mode``: "f16x3" (default: tensor cores at fp32-level accuracy,
    within 1e-5 of the fp64 oracle like the reference's fp32 Keras predict), "tf32x3" (the same accuracy without a
    range restriction, half the speed), "bf16" (3x faster, ~6e-3), "fp32" (CUDA cores)."""
    return CNN2Model("vt", 0, classes, mode, device, flatten)


def load_model(filepath: str, device: Optional[int] = None, mode: str = "f16x3") -> CNN2Model:
    """Keras ``load_model``: topology from the ``model_config`` embedded in the ``.h5``
    (SURVEY.md Appendix B.1), then ``load_weights``."""
    cfg = _config_of(filepath)
    if cfg is None:
        raise ValueError(f"{filepath}: no model_config (weights-only file): build the model, then load_weights")
    layers = cfg["config"]["layers"]
    convs = [l["config"] for l in layers if l["class_name"] in ("Conv2D", "Convolution2D")]
    denses = [l["config"] for l in layers if l["class_name"] == "Dense"]
    if len(convs) == 1 and len(denses) == 1 and tuple(convs[0]["kernel_size"]) == (1, 2):
        m = tiny_cnn2(int(convs[0]["filters"]), int(denses[0]["units"]), device)
    elif len(convs) == 2 and len(denses) == 2 and int(convs[0]["filters"]) == 256 and int(convs[1]["filters"]) == 80:
        fmt = convs[0].get("data_format", "channels_last")
        m = vt_cnn2(int(denses[1]["units"]), mode, device, fmt)
    else:
        raise ValueError(f"{filepath}: topology is neither TinyCNN2 nor VT-CNN2")
    m.load_weights(filepath)
    return m
