"""Soak of the streaming host call: random batch sizes, up to 3 calls in flight, every result compared with the
blocking call's.  python tools/stress_async.py [iterations=150]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import synth  # noqa: E402
from modulationdetectioncnn_b200.model import vt_cnn2  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(11)
m = vt_cnn2(11, mode="bf16")
m.set_weights(synth.vt_cnn2_weights(11, 1602))
pool = torch.from_numpy(synth.iq_frames(70000, seed=5)).pin_memory().numpy()
inflight, bad = [], 0
for it in range(iters):
    n = int(rng.choice([1, 7, 300, 2048, 2049, 8192, 16384, 16385, 33000, 65536, 70000]))
    off = int(rng.integers(0, 70000 - n + 1))
    x = pool[off:off + n]
    inflight.append((m.predict_async(x, output="dense"), x))
    if len(inflight) > int(rng.integers(1, 4)):
        p, xx = inflight.pop(0)
        got = p.result()
        want = m.predict(xx, output="dense")          # blocking call while others are still in flight
        if not np.array_equal(got, want):
            bad += 1
            print(f"iteration {it}: mismatch for n={len(xx)}")
for p, xx in inflight:
    if not np.array_equal(p.result(), m.predict(xx, output="dense")):
        bad += 1
print(f"{iters} streaming calls, {bad} mismatches")
sys.exit(1 if bad else 0)
