"""BASELINE configs[2] / SURVEY C3: checkpoint sweep and precision-mode tolerances on the GPU.

For each shipped checkpoint (the four `2/3/4/5conv` retrains, F=3, and the 10-filter `convmodrecnets_CNN2_0.5`)
and for the VT-CNN2 stack in its four arithmetic modes, compare the CUDA path with the fp64 oracle on the same
synthetic frames: max error of the last Dense output relative to the frame's largest |value|, max |softmax| error,
argmax agreement.  Test tool: uses oracle/ as the checker.

    python tests/tools/sweep_c3.py [out.md]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from modulationdetectioncnn_b200 import synth                      # noqa: E402
from modulationdetectioncnn_b200.model import tiny_cnn2, vt_cnn2    # noqa: E402
from oracle import cnn2_float as cf                                 # noqa: E402


def compare(z, p, cls, ref_z, ref_p):
    scale = np.maximum(np.abs(ref_z).max(-1, keepdims=True), 1e-30)
    top2 = np.sort(ref_z, -1)[:, -2:]
    margin = (top2[:, 1] - top2[:, 0]) / scale[:, 0]
    err = np.abs(z - ref_z) / scale
    agree = cls == ref_z.argmax(-1)
    return {"dense_rel_err_max": float(err.max()), "softmax_abs_err_max": float(np.abs(p - ref_p).max()),
            "argmax_agree": float(agree.mean()),
            # disagreements can only be ties within the error: largest oracle top-2 margin among them
            "worst_margin_of_disagreement": float(margin[~agree].max()) if (~agree).any() else 0.0}


def main():
    rows = []
    hw = np.load(os.path.join(ROOT, "tests", "golden", "h5_weights.npz"))
    x = synth.iq_frames(65536, seed=2016)
    for tag, name in (("B_2conv", "2convmodrecnets_CNN2_0.5"), ("A_3conv", "3convmodrecnets_CNN2_0.5"),
                      ("D_4conv", "4convmodrecnets_CNN2_0.5"), ("C_5conv", "5convmodrecnets_CNN2_0.5"),
                      ("E_f10", "convmodrecnets_CNN2_0.5")):
        w = [hw[f"{tag}_{k}"] for k in ("conv_k", "conv_b", "dense_k", "dense_b")]
        m = tiny_cnn2(w[0].shape[-1], 3)
        m.set_weights(w)
        # x64 exercises the ReLU-active regime too (N(0, 2^-7) frames mostly sit below the dense ReLU)
        for label, xx in (("N(0,2^-7)", x), ("x64", x * 64)):
            ref_z = cf.tiny_cnn2_forward(xx, *w, output="dense")
            ref_p = cf.tiny_cnn2_forward(xx, *w, output="softmax")
            r = compare(m.predict(xx, output="dense"), m.predict(xx), m.predict_classes(xx), ref_z, ref_p)
            rows.append({"model": f"TinyCNN2 F={w[0].shape[-1]} {name}", "mode": "fp32", "frames": len(xx), "input": label, **r})
    wv = synth.vt_cnn2_weights(11, 1602)
    kw = cf.vt_cnn2_init(11, 1602)
    for mode, n in (("fp32", 2048), ("f16x3", 8192), ("tf32x3", 8192), ("bf16", 8192)):
        xx = x[:n].copy()
        xx[: n // 8] *= 64
        ref_z = cf.vt_cnn2_forward(xx, **kw, output="logits")
        ref_p = cf.vt_cnn2_forward(xx, **kw, output="softmax")
        m = vt_cnn2(11, mode=mode)
        m.set_weights(wv)
        r = compare(m.predict(xx, output="dense"), m.predict(xx), m.predict_classes(xx), ref_z, ref_p)
        rows.append({"model": "VT-CNN2 11-class (synthetic weights)", "mode": mode, "frames": n, "input": "N(0,2^-7), 1/8 x64", **r})
    lines = ["| model | mode | frames | input | max Dense err / max abs Dense | max abs softmax err | argmax agreement | worst oracle top-2 margin among disagreements |",
             "|---|---|---|---|---|---|---|---|"]
    for r in rows:
        lines.append(f"| {r['model']} | {r['mode']} | {r['frames']} | {r['input']} | {r['dense_rel_err_max']:.3g} | "
                     f"{r['softmax_abs_err_max']:.3g} | {r['argmax_agree']:.6f} | {r['worst_margin_of_disagreement']:.3g} |")
    text = "\n".join(lines)
    print(text)
    print(json.dumps(rows))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as fh:
            fh.write(text + "\n")


if __name__ == "__main__":
    main()
