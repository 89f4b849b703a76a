"""GPU debugging aid: VT-CNN2 tf32x3 path, layer by layer, against the fp64 oracle."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from modulationdetectioncnn_b200 import _lib, synth          # noqa: E402
from modulationdetectioncnn_b200.model import vt_cnn2          # noqa: E402
from oracle import cnn2_float as cf                            # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 37
w = synth.vt_cnn2_weights(11, 1602)
x = synth.iq_frames(n, seed=7)
x[: max(1, n // 8)] *= 64
m = vt_cnn2(11, mode="tf32x3")
m.set_weights(w)
z = m.predict(x, output="dense")
lib = m._h._lib
buf = np.empty((2, n * 132, 80), np.float32)
got = C.c_size_t()
_lib.check(lib.mdc_debug_read(m._h.ptr, 0, buf.ctypes.data, buf.nbytes, C.byref(got)))
hb = np.empty((n, 256), np.float32)
_lib.check(lib.mdc_debug_read(m._h.ptr, 1, hb.ctypes.data, hb.nbytes, C.byref(got)))
kw = cf.vt_cnn2_init(11, 1602)
c2 = cf.vt_cnn2_forward(x, **kw, output="conv2")
h = cf.vt_cnn2_forward(x, **kw, output="dense1")
lg = cf.vt_cnn2_forward(x, **kw, output="logits")
act = (buf[0].astype(np.float64) + buf[1].astype(np.float64)).reshape(n, 132, 80)
print("hi/lo split: lo max / hi max =", np.abs(buf[1]).max() / np.abs(buf[0]).max())
print("conv2 rel err (vs max):", np.abs(act - c2).max() / np.abs(c2).max(), " hi-only:", np.abs(buf[0].reshape(n, 132, 80) - c2).max() / np.abs(c2).max())
print("dense1 rel err:", np.abs(hb - h).max() / np.abs(h).max())
print("logits rel err:", (np.abs(z - lg) / np.abs(lg).max(-1, keepdims=True)).max())
# fp32 reference of the same quantities for scale
c2f = cf.vt_cnn2_forward(x, **kw, output="conv2", dtype=np.float32)
print("numpy fp32 conv2 rel err:", np.abs(c2f - c2).max() / np.abs(c2).max())
# RZ-accumulation hypothesis: the tensor core truncates every accumulate toward zero, so positive sums
# come out systematically LOW by about (chain length) x 2^-25
big = c2 > 0.25 * c2.max()
print("conv2 signed rel err on large outputs: mean %.3e  min %.3e  max %.3e" % (((act - c2) / c2)[big].mean(), ((act - c2) / c2)[big].min(), ((act - c2) / c2)[big].max()))
bigh = h > 0.25 * h.max()
print("dense1 signed rel err on large outputs: mean %.3e  min %.3e  max %.3e" % (((hb - h) / h)[bigh].mean(), ((hb - h) / h)[bigh].min(), ((hb - h) / h)[bigh].max()))
