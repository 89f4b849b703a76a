"""Small, ragged invocations of every kernel (sizes 1, 3, 13, 77, ... around the tile boundaries).  Written for
`compute-sanitizer --tool memcheck python tools/sanitize_small.py`; on pools where the sanitizer is closed it
still serves as a crash / hang smoke sweep (results themselves are checked by tests/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modulationdetectioncnn_b200 import export, metrics, sdr, synth  # noqa: E402
from modulationdetectioncnn_b200.fwht import fwht  # noqa: E402
from modulationdetectioncnn_b200.model import tiny_cnn2, vt_cnn2  # noqa: E402
from modulationdetectioncnn_b200.qmodel import FixedPointCNN2  # noqa: E402
from modulationdetectioncnn_b200.svtext import QWeights  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "qweights.npz"))
hw = np.load(os.path.join(ROOT, "tests", "golden", "h5_weights.npz"))
which = sys.argv[1:] or ["q612", "tiny", "fwht", "sdr", "metrics", "vt"]
if "q612" in which:
    q = FixedPointCNN2(3, 3)
    q.set_tables(QWeights(g["A_conv_tab"], g["A_dense_bias"], g["A_dense_tabs"]))
    for n in (1, 77, 1000):
        q.predict(synth.q612_frames(n), output="pre")
        q.predict((synth.q612_frames(n) * 4000).astype(np.int32), output="out")      # exact 36-bit path
    q10 = FixedPointCNN2(10, 3)
    q10.set_tables(export.qweights_from_dense_dump([hw[f"E_f10_{k}"] for k in ("conv_k", "conv_b", "dense_k", "dense_b")], g["E_dense_flat"]))
    q10.predict(synth.q612_frames(333))
    print("q612 ok")
if "tiny" in which:
    for tag in ("A_3conv", "E_f10"):
        w = [hw[f"{tag}_{k}"] for k in ("conv_k", "conv_b", "dense_k", "dense_b")]
        m = tiny_cnn2(w[0].shape[-1], 3)
        m.set_weights(w)
        for n in (1, 13, 15, 1000):
            m.predict(synth.iq_frames(n))
            m.class_histogram(synth.iq_frames(n))
    print("tiny ok")
if "fwht" in which:
    for lg in (5, 10, 13):
        x = synth.q612_frames(64).reshape(-1, 1 << lg)
        fwht(x)
        fwht(x, ordering="sequency")
    print("fwht ok")
if "sdr" in which:
    raw = torch.randint(0, 256, (2 * 1024 * 3,), dtype=torch.uint8, device="cuda")
    sdr.ingest_u8(raw, ("f32", "q612", "fwht"))
    print("sdr ok")
if "metrics" in which:
    t = torch.randint(0, 11, (5000,), device="cuda")
    p = torch.randint(0, 11, (5000,), device="cuda")
    metrics.confusion_matrix(t, p, 11)
    metrics.accuracy_by_snr(t, p, torch.randint(-10, 10, (5000,), device="cuda"), 11)
    print("metrics ok")
if "vt" in which:
    w = synth.vt_cnn2_weights(11, 1602)
    for mode in ("bf16", "tf32x3", "fp32"):
        m = vt_cnn2(11, mode=mode)
        m.set_weights(w)
        for n in (1, 3, 130):
            m.predict(synth.iq_frames(n))
    print("vt ok")
