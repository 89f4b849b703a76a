"""Diagnostic: host time of the streaming call (submit) against the batch's total time, per mode and size."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import _lib, synth  # noqa: E402
from modulationdetectioncnn_b200.model import vt_cnn2  # noqa: E402

w = synth.vt_cnn2_weights(11, 1602)
for mode in ("f16x3", "bf16"):
    m = vt_cnn2(11, mode=mode)
    m.set_weights(w)
    lib, h = m._h._lib, m._h
    for n in (65536, 4 * 65536):
        xp = torch.randn((n, 2, 128)).mul_(2.0 ** -7).pin_memory().numpy()
        out = torch.empty((n, 11)).pin_memory().numpy()
        for rep in range(3):
            t = C.c_int64(0)
            t0 = time.perf_counter()
            _lib.check(lib.mdc_predict_raw_host_async(h.ptr, xp.ctypes.data, 0, n, out.ctypes.data, None, None, None, C.byref(t)))
            t1 = time.perf_counter()
            _lib.check(lib.mdc_host_wait(h.ptr, t.value))
            t2 = time.perf_counter()
            print(f"{mode} n={n} rep={rep}: raw ABI submit {1e3 * (t1 - t0):.2f} ms, total {1e3 * (t2 - t0):.2f} ms", flush=True)
        for rep in range(3):
            t0 = time.perf_counter()
            p = m.predict_async(xp)
            t1 = time.perf_counter()
            p.result()
            t2 = time.perf_counter()
            print(f"{mode} n={n} rep={rep}: facade submit {1e3 * (t1 - t0):.2f} ms, total {1e3 * (t2 - t0):.2f} ms", flush=True)
    m.close()
