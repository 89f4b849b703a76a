"""Where does the host-buffer predict call spend its time?  (diagnostic)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import _lib, synth  # noqa: E402
from modulationdetectioncnn_b200.model import vt_cnn2  # noqa: E402

n = 65536
xh = torch.randn((n, 2, 128)).mul_(2.0 ** -7).pin_memory()
xd = torch.empty_like(xh, device="cuda")
for sz in (n, n // 8):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        for c in range(0, n, sz):
            xd[c:c + sz].copy_(xh[c:c + sz], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 10
    print(f"H2D 64 MiB in {n // sz} copies: {dt * 1e3:.3f} ms  {n * 1024 / dt / 1e9:.1f} GB/s")
m = vt_cnn2(11, mode="bf16")
m.set_weights(synth.vt_cnn2_weights(11, 1602))
lib, h = m._h._lib, m._h
x = xh.numpy()
p = torch.empty((n, 11)).pin_memory().numpy()
hist = np.zeros(11, np.uint64)
for what, args in (("probs+hist", (p.ctypes.data, None, None, hist.ctypes.data)), ("probs", (p.ctypes.data, None, None, None)),
                   ("hist only", (None, None, None, hist.ctypes.data))):
    for _ in range(3):
        _lib.check(lib.mdc_predict_f32_host(h.ptr, x.ctypes.data, n, *args))
    t = time.perf_counter()
    for _ in range(10):
        _lib.check(lib.mdc_predict_f32_host(h.ptr, x.ctypes.data, n, *args))
    dt = (time.perf_counter() - t) / 10
    print(f"predict_f32_host {what}: {dt * 1e3:.3f} ms  {n / dt:.4g} frames/s")
xdev = xd.copy_(xh)
pd = torch.empty((n, 11), device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.mdc_predict_f32(h.ptr, xdev.data_ptr(), n, pd.data_ptr(), None, None, None, stream))
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10):
    _lib.check(lib.mdc_predict_f32(h.ptr, xdev.data_ptr(), n, pd.data_ptr(), None, None, None, stream))
torch.cuda.synchronize()
print(f"predict_f32 device: {(time.perf_counter() - t) / 10 * 1e3:.3f} ms")

# streaming form: two calls in flight
import ctypes as C
xs = [torch.randn((n, 2, 128)).mul_(2.0 ** -7).pin_memory().numpy() for _ in range(2)]
ps = [torch.empty((n, 11)).pin_memory().numpy() for _ in range(2)]
def stream(k):
    pend = None
    for i in range(k):
        t = C.c_int64(0)
        _lib.check(lib.mdc_predict_f32_host_async(h.ptr, xs[i % 2].ctypes.data, n, ps[i % 2].ctypes.data, None, None, None, C.byref(t)))
        if pend is not None:
            _lib.check(lib.mdc_host_wait(h.ptr, pend))
        pend = t.value
    _lib.check(lib.mdc_host_wait(h.ptr, pend))
stream(4)
t = time.perf_counter()
stream(20)
dt = (time.perf_counter() - t) / 20
print(f"predict_f32_host_async stream: {dt * 1e3:.3f} ms  {n / dt:.4g} frames/s")
