"""Race check for the cross-CTA barrier protocol of the tensor-core kernels: the same 65,536 frames, many passes,
every pass must be bit-identical to the first (and the first is checked against the oracle by tests/).
    python tools/stress_determinism.py [mode=bf16] [passes=200]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import _lib, synth  # noqa: E402
from modulationdetectioncnn_b200.model import vt_cnn2  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 200
n = 65536
m = vt_cnn2(11, mode=mode)
m.set_weights(synth.vt_cnn2_weights(11, 1602))
x = torch.randn((n, 2, 128), device="cuda").mul_(2.0 ** -7)
x[: n // 8] *= 64
st = torch.cuda.current_stream().cuda_stream
ref = torch.empty((n, 11), device="cuda")
_lib.check(m._h._lib.mdc_predict_f32(m._h.ptr, x.data_ptr(), n, None, ref.data_ptr(), None, None, st))
torch.cuda.synchronize()
out = torch.empty_like(ref)
bad = 0
for i in range(passes):
    out.fill_(float("nan"))
    _lib.check(m._h._lib.mdc_predict_f32(m._h.ptr, x.data_ptr(), n, None, out.data_ptr(), None, None, st))
    if not torch.equal(out, ref):
        bad += 1
        d = (out != ref).any(-1).nonzero().flatten()
        print(f"pass {i}: {d.numel()} frames differ, first {d[:5].tolist()}")
print(f"{mode}: {passes} passes x {n} frames, {bad} passes differ from the first")
sys.exit(1 if bad else 0)
