"""Host-buffer e2e of the HBM-bound paths per frame format: python tools/e2e_small.py  (MDC_HOST_CHUNK=frames per chunk)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import _lib, synth  # noqa: E402
from modulationdetectioncnn_b200.model import tiny_cnn2  # noqa: E402
from modulationdetectioncnn_b200.qmodel import FixedPointCNN2  # noqa: E402
from modulationdetectioncnn_b200.svtext import QWeights  # noqa: E402

nh = 1 << 18
g = np.load(os.path.join(ROOT, "tests", "golden", "qweights.npz"))
hw = np.load(os.path.join(ROOT, "tests", "golden", "h5_weights.npz"))
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
x32 = pin(synth.q612_frames(nh))
x16 = pin(synth.q612_frames(nh).astype(np.int16))
xu8 = pin(np.random.default_rng(1).integers(0, 256, (nh, 128, 2), dtype=np.uint8))
xf = pin(synth.iq_frames(nh))
oq = pin(np.empty((nh, 3), np.int32))
pf = pin(np.empty((nh, 3), np.float32))


def t(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    print(f"chunk {os.environ.get('MDC_HOST_CHUNK', 'default')} {name}: {nh / dt:.3e} frames/s", flush=True)


q = FixedPointCNN2(3, 3)
q.set_tables(QWeights(g["A_conv_tab"], g["A_dense_bias"], g["A_dense_tabs"]))
lib = q._h._lib
for name, x, fmt in (("q612 i32", x32, _lib.IN_I32), ("q612 i16", x16, _lib.IN_I16), ("q612 u8", xu8, _lib.IN_U8IQ)):
    t(name, lambda: _lib.check(lib.mdc_predict_q612_raw_host(q._h.ptr, x.ctypes.data, fmt, nh, oq.ctypes.data, None, None, None)))
w = [hw[f"E_f10_{k}"] for k in ("conv_k", "conv_b", "dense_k", "dense_b")]
m = tiny_cnn2(10, 3)
m.set_weights(w)
for name, x, fmt in (("tiny10 f32", xf, _lib.IN_F32), ("tiny10 u8", xu8, _lib.IN_U8IQ)):
    t(name, lambda: _lib.check(lib.mdc_predict_raw_host(m._h.ptr, x.ctypes.data, fmt, nh, pf.ctypes.data, None, None, None)))
