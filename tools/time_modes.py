"""Device-resident timing of the VT-CNN2 arithmetic modes: whole step and the conv kernel's share (CUDA events).
    python tools/time_modes.py [modes=f16x3,bf16,tf32x3] [steps=10]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import _lib, synth  # noqa: E402
from modulationdetectioncnn_b200.model import vt_cnn2  # noqa: E402

modes = (sys.argv[1] if len(sys.argv) > 1 else "f16x3,bf16,tf32x3").split(",")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = 65536
w = synth.vt_cnn2_weights(11, 1602)
xs = [torch.randn((n, 2, 128), device="cuda").mul_(2.0 ** -7) for _ in range(4)]
out = torch.empty((n, 11), device="cuda")
st = torch.cuda.current_stream().cuda_stream
for mode in modes:
    m = vt_cnn2(11, mode=mode)
    m.set_weights(w)
    lib, h = m._h._lib, m._h
    for i in range(3):
        _lib.check(lib.mdc_predict_f32(h.ptr, xs[i % 4].data_ptr(), n, out.data_ptr(), None, None, None, st))
    torch.cuda.synchronize()
    h.profile_enable(True)
    h.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        _lib.check(lib.mdc_predict_f32(h.ptr, xs[i % 4].data_ptr(), n, out.data_ptr(), None, None, None, st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kms, kl, kname = h.profile_read()
    print(f"{mode}: {ms:.3f} ms/step = {n / ms * 1e3:.3e} frames/s; {kname} {kms / max(kl, 1):.3f} ms x {kl / steps:.1f} per step "
          f"(rest {ms - kms / steps:.3f} ms); range flags {h.range_flags()}", flush=True)
    m.close()
