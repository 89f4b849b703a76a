"""ctypes binding of libmdc.so (C ABI declared in include/mdc.h).

There is no CPU fallback: if the shared library is missing, or a compute entry
point is called without a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmdc.so")

# enums of include/mdc.h
MDC_OK = 0
MODEL_TINY, MODEL_VT = 0, 1
MODE_FP32, MODE_BF16, MODE_TF32X3, MODE_Q612, MODE_F16X3 = 0, 1, 2, 3, 4
IN_F32, IN_U8IQ, IN_I16, IN_I32 = 0, 1, 2, 3
ERR_RANGE = -5
T_CONV1_K, T_CONV1_B, T_CONV2_K, T_CONV2_B, T_DENSE1_K, T_DENSE1_B, T_DENSE2_K, T_DENSE2_B = range(8)
OPT_FLATTEN_ORDER = 0
FWHT_NATURAL, FWHT_SEQUENCY = 0, 1

EXPORTS = [
    "mdc_create", "mdc_destroy", "mdc_set_option", "mdc_set_weights_f32", "mdc_set_weights_q612",
    "mdc_predict_f32", "mdc_predict_f32_host", "mdc_predict_f32_host_async", "mdc_host_wait", "mdc_predict_q612", "mdc_predict_q612_host_async", "mdc_predict_q612_host",
    "mdc_predict_q612_raw", "mdc_predict_q612_raw_host", "mdc_predict_q612_raw_host_async",
    "mdc_fwht_i32", "mdc_fwht_i32_host", "mdc_sdr_ingest_u8", "mdc_confusion_i32", "mdc_confusion_grouped_i32", "mdc_last_error", "mdc_version",
    "mdc_launch_count", "mdc_profile_enable", "mdc_profile_read", "mdc_debug_read",
    "mdc_reserve", "mdc_predict_raw", "mdc_predict_raw_host", "mdc_predict_raw_host_async", "mdc_range_flags",
]


class MdcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmdc error {code}: {msg}")
        self.code = code


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m modulationdetectioncnn_b200.build` "
            "(needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    sig = {
        "mdc_create": (i32, [i32, i32, i32, i32, i32, C.POINTER(vp)]),
        "mdc_destroy": (i32, [vp]),
        "mdc_set_option": (i32, [vp, i32, i32]),
        "mdc_set_weights_f32": (i32, [vp, i32, vp, sz]),
        "mdc_set_weights_q612": (i32, [vp, vp, vp, vp]),
        "mdc_predict_f32": (i32, [vp, vp, i64, vp, vp, vp, vp, vp]),
        "mdc_predict_f32_host": (i32, [vp, vp, i64, vp, vp, vp, vp]),
        "mdc_predict_f32_host_async": (i32, [vp, vp, i64, vp, vp, vp, vp, C.POINTER(i64)]),
        "mdc_host_wait": (i32, [vp, i64]),
        "mdc_predict_q612": (i32, [vp, vp, i64, vp, vp, vp, vp, vp]),
        "mdc_predict_q612_host": (i32, [vp, vp, i64, vp, vp, vp, vp]),
        "mdc_predict_q612_host_async": (i32, [vp, vp, i64, vp, vp, vp, vp, C.POINTER(i64)]),
        "mdc_predict_q612_raw": (i32, [vp, vp, i32, i64, vp, vp, vp, vp, vp]),
        "mdc_predict_q612_raw_host": (i32, [vp, vp, i32, i64, vp, vp, vp, vp]),
        "mdc_predict_q612_raw_host_async": (i32, [vp, vp, i32, i64, vp, vp, vp, vp, C.POINTER(i64)]),
        "mdc_fwht_i32": (i32, [vp, vp, i64, i32, i32, vp]),
        "mdc_fwht_i32_host": (i32, [vp, vp, i64, i32, i32, i32]),
        "mdc_sdr_ingest_u8": (i32, [vp, i64, vp, vp, vp, vp]),
        "mdc_confusion_i32": (i32, [vp, vp, i64, i32, vp, vp]),
        "mdc_confusion_grouped_i32": (i32, [vp, vp, vp, i64, i32, i32, vp, vp]),
        "mdc_last_error": (C.c_char_p, []),
        "mdc_version": (C.c_char_p, []),
        "mdc_launch_count": (i64, [vp]),
        "mdc_profile_enable": (i32, [vp, i32]),
        "mdc_profile_read": (i32, [vp, C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_char_p)]),
        "mdc_debug_read": (i32, [vp, i32, vp, sz, C.POINTER(sz)]),
        "mdc_reserve": (i32, [vp, i64]),
        "mdc_predict_raw": (i32, [vp, vp, i32, i64, vp, vp, vp, vp, vp]),
        "mdc_predict_raw_host": (i32, [vp, vp, i32, i64, vp, vp, vp, vp]),
        "mdc_predict_raw_host_async": (i32, [vp, vp, i32, i64, vp, vp, vp, vp, C.POINTER(i64)]),
        "mdc_range_flags": (i32, [vp, C.POINTER(C.c_uint), i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != MDC_OK:
        raise MdcError(code, load().mdc_last_error().decode("utf-8", "replace"))


def version() -> str:
    return load().mdc_version().decode()


class Handle:
    """Owns one ``mdc_handle_t``."""

    def __init__(self, model: int, filters: int, classes: int, mode: int, device: int = 0):
        self._lib = load()
        self._h = C.c_void_p()
        check(self._lib.mdc_create(model, filters, classes, mode, device, C.byref(self._h)))
        self.model, self.filters, self.classes, self.mode, self.device = model, filters, classes, mode, device

    @property
    def ptr(self) -> C.c_void_p:
        if not self._h:
            raise RuntimeError("handle destroyed")
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.mdc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self) -> int:
        return int(self._lib.mdc_launch_count(self.ptr))

    def reserve(self, max_frames: int) -> None:
        check(self._lib.mdc_reserve(self.ptr, int(max_frames)))

    def range_flags(self, reset: bool = True) -> int:
        f = C.c_uint(0)
        check(self._lib.mdc_range_flags(self.ptr, C.byref(f), int(reset)))
        return int(f.value)

    def profile_enable(self, on: bool = True) -> None:
        check(self._lib.mdc_profile_enable(self.ptr, int(on)))

    def profile_read(self):
        ms, n, name = C.c_double(), C.c_int64(), C.c_char_p()
        check(self._lib.mdc_profile_read(self.ptr, C.byref(ms), C.byref(n), C.byref(name)))
        return ms.value, n.value, (name.value or b"").decode()
