"""Timing driver for the SDR ingest kernel (diagnostic): python tools/prof_sdr.py [log2_samples=28]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import _lib  # noqa: E402

ns = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
raw = torch.randint(0, 256, (2 * ns,), device="cuda", dtype=torch.uint8)
f = torch.empty((ns // 128, 2, 128), device="cuda")
q = torch.empty((ns // 128, 256), dtype=torch.int32, device="cuda")
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for outs, b in ((("f32", "q612"), 2304), (("f32",), 1280)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(4):
        e0.record()
        _lib.check(lib.mdc_sdr_ingest_u8(raw.data_ptr(), ns, f.data_ptr(), q.data_ptr() if "q612" in outs else None, None, st))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{'+'.join(outs)}: {ms:.3f} ms  {ns / 128 / ms * 1e3:.4g} frames/s  {b * (ns // 128) / ms / 1e6:.0f} GB/s")
