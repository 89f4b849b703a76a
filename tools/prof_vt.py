"""Minimal driver for ncu: a few VT-CNN2 predict passes on device-resident frames.

    python tools/prof_vt.py [mode=bf16] [batch=65536] [passes=3]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import _lib, synth  # noqa: E402
from modulationdetectioncnn_b200.model import vt_cnn2  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
m = vt_cnn2(11, mode=mode)
m.set_weights(synth.vt_cnn2_weights(11, 1602))
x = torch.randn((batch, 2, 128), device=dev).mul_(2.0 ** -7)
probs = torch.empty((batch, 11), device=dev)
hist = torch.zeros(11, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
m._h.profile_enable(True)
for i in range(passes):
    e0.record()
    _lib.check(m._h._lib.mdc_predict_f32(m._h.ptr, x.data_ptr(), batch, probs.data_ptr(), None, None, hist.data_ptr(), stream))
    e1.record()
    torch.cuda.synchronize()
    kms, kl, kn = m._h.profile_read()
    print(f"pass {i}: {e0.elapsed_time(e1):.3f} ms  {batch / e0.elapsed_time(e1) * 1e3:.4g} frames/s   {kn}: {kms / max(kl, 1):.3f} ms")
if not os.environ.get("MDC_VT_DEBUG"):
    assert int(hist.sum()) == batch * passes
