import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from modulationdetectioncnn_b200 import _lib, synth
from modulationdetectioncnn_b200.model import vt_cnn2
n = 65536
xh = torch.randn((n, 2, 128)).mul_(2.0 ** -7).pin_memory().numpy()
p = torch.empty((n, 11)).pin_memory().numpy()
m = vt_cnn2(11, mode="bf16"); m.set_weights(synth.vt_cnn2_weights(11, 1602))
lib, h = m._h._lib, m._h
for i in range(4):
    sys.stderr.write(f"call {i}\n")
    _lib.check(lib.mdc_predict_f32_host(h.ptr, xh.ctypes.data, n, p.ctypes.data, None, None, None))
