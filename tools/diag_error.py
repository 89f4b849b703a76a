"""Diagnostic: max relative logit error of each VT-CNN2 mode against the fp64 oracle (n frames)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modulationdetectioncnn_b200 import synth  # noqa: E402
from modulationdetectioncnn_b200.model import vt_cnn2  # noqa: E402
from oracle import cnn2_float as cf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
modes = (sys.argv[2] if len(sys.argv) > 2 else "f16x3,tf32x3,fp32,bf16").split(",")
w = cf.vt_cnn2_init(11, 1602)
x = synth.iq_frames(n, seed=7)
x[: n // 8] *= 64
ref = cf.vt_cnn2_forward(x, **w, output="logits")
scale = np.abs(ref).max(-1, keepdims=True)
for mode in modes:
    m = vt_cnn2(11, mode=mode)
    m.set_weights([w[k] for k in ("w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4")])
    z = m.predict(x, output="dense")
    e = (z - ref) / scale
    print(f"{mode}: max |err| {np.abs(e).max():.3e}  mean err {e.mean():+.3e}  rms {np.sqrt((e ** 2).mean()):.3e}  "
          f"mean err on largest logit {((z - ref)[np.arange(n), np.abs(ref).argmax(-1)] / scale[:, 0] * np.sign(ref[np.arange(n), np.abs(ref).argmax(-1)])).mean():+.3e}", flush=True)
    m.close()
