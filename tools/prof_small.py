"""Minimal driver for ncu: the HBM-bound paths (integer SV-exact, TinyCNN2 F=3 / F=10, FWHT).

    python tools/prof_small.py [which=all|q612|tiny3|tiny10|fwht] [log2_frames=21] [passes=3]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import _lib  # noqa: E402
from modulationdetectioncnn_b200.model import tiny_cnn2  # noqa: E402
from modulationdetectioncnn_b200.qmodel import FixedPointCNN2  # noqa: E402
from modulationdetectioncnn_b200.svtext import QWeights  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 21)
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
g = np.load(os.path.join(ROOT, "tests", "golden", "qweights.npz"))
hw = np.load(os.path.join(ROOT, "tests", "golden", "h5_weights.npz"))
lib = _lib.load()


def timeit(name, fn, bytes_per_unit, units):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(passes):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print(f"{name}: {ms:.3f} ms  {units / ms * 1e3:.4g} units/s  {bytes_per_unit * units / ms / 1e6:.0f} GB/s")


if which in ("all", "q612"):
    x = torch.randn((n, 256), device=dev).mul_(32).trunc_().to(torch.int32)
    o = torch.empty((n, 3), dtype=torch.int32, device=dev)
    q = FixedPointCNN2(3, 3)
    q.set_tables(QWeights(g["A_conv_tab"], g["A_dense_bias"], g["A_dense_tabs"]))
    timeit("q612", lambda: _lib.check(lib.mdc_predict_q612(q._h.ptr, x.data_ptr(), n, o.data_ptr(), None, None, None, stream)), 1036, n)
    del x, o
if which in ("all", "q612f10"):
    from modulationdetectioncnn_b200 import export
    x = torch.randn((n, 256), device=dev).mul_(32).trunc_().to(torch.int32)
    o = torch.empty((n, 3), dtype=torch.int32, device=dev)
    q10 = FixedPointCNN2(10, 3)
    q10.set_tables(export.qweights_from_dense_dump([hw[f"E_f10_{k}"] for k in ("conv_k", "conv_b", "dense_k", "dense_b")], g["E_dense_flat"]))
    timeit("q612f10", lambda: _lib.check(lib.mdc_predict_q612(q10._h.ptr, x.data_ptr(), n, o.data_ptr(), None, None, None, stream)), 1036, n)
    del x, o
for tag, key in (("tiny3", "A_3conv"), ("tiny10", "E_f10")):
    if which in ("all", tag):
        w = [hw[f"{key}_{k}"] for k in ("conv_k", "conv_b", "dense_k", "dense_b")]
        m = tiny_cnn2(w[0].shape[-1], 3)
        m.set_weights(w)
        x = torch.randn((n, 2, 128), device=dev).mul_(2.0 ** -7)
        p = torch.empty((n, 3), device=dev)
        timeit(tag, lambda: _lib.check(lib.mdc_predict_f32(m._h.ptr, x.data_ptr(), n, p.data_ptr(), None, None, None, stream)), 1036, n)
        del x, p
if which in ("all", "fwht"):
    s = n >> 3
    x = torch.randn((s, 1024), device=dev).mul_(32).trunc_().to(torch.int32)
    y = torch.empty_like(x)
    timeit("fwht", lambda: _lib.check(lib.mdc_fwht_i32(x.data_ptr(), y.data_ptr(), s, 10, 0, stream)), 8192, s)
