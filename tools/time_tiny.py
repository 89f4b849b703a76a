"""Device-resident timing of the TinyCNN2 fp32 kernels (real checkpoints): python tools/time_tiny.py [log2n=21]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import _lib  # noqa: E402
from modulationdetectioncnn_b200.model import tiny_cnn2  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
n = 1 << log2n
hw = np.load(os.path.join(ROOT, "tests", "golden", "h5_weights.npz"))
x = torch.randn((n, 2, 128), device="cuda").mul_(2.0 ** -7)
out = torch.empty((n, 3), device="cuda")
st = torch.cuda.current_stream().cuda_stream
for tag in ("A_3conv", "E_f10"):
    w = [hw[f"{tag}_{k}"] for k in ("conv_k", "conv_b", "dense_k", "dense_b")]
    m = tiny_cnn2(w[0].shape[-1], 3)
    m.set_weights(w)
    lib, h = m._h._lib, m._h
    for nn, reps in ((n, 10), (65536, 64)):
        offs = [(i * nn) % max(n - nn + 1, 1) for i in range(reps)]
        for _ in range(3):
            _lib.check(lib.mdc_predict_f32(h.ptr, x.data_ptr(), nn, out.data_ptr(), None, None, None, st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for o in offs:
            _lib.check(lib.mdc_predict_f32(h.ptr, x[o:].data_ptr(), nn, out.data_ptr(), None, None, None, st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"F={w[0].shape[-1]} n={nn}: {ms * 1e3:.1f} us  "
              f"{nn / ms * 1e3:.3e} frames/s  {nn * 1036 / ms / 1e6:.0f} GB/s", flush=True)
    m.close()
