#!/usr/bin/env python
"""Generate tests/golden/* from the read-only reference checkout.

Run in the build container (``/root/reference`` present):

    python tests/tools/make_golden.py

The GPU box has no ``/root/reference``; GPU parity tests, ``smoke()`` and
``bench.py`` read only the small fixtures written here.  Everything is
*decoded* data (integers / float arrays / recorded numbers), produced through
this repo's own readers; no reference source text is copied.

Outputs
-------
qweights.npz     literal ROM contents of weight sets A-D
                 + set E dense table (F=10, DenseWeights1.txt)
vectors.npz      the 16 test vectors (int32 [16,256], Verilog over-width policy),
                 their names, and the "zero"-policy variant
h5_weights.npz   float32 weights of the five *.wts.h5 checkpoints (+ class count, F)
sv_roms.npz      ROMs parsed straight out of cnn_test_latest1.sv
kat.json         what the reference itself records: Keras Dense+ReLU outputs
                 (12.16.testDataYunyun.txt:2,264; CNN.ipynb cells 18,21), the
                 un-quantised cell-18 input frame, float2fix I/O pairs (cells 21,25)
int_goldens.json integer-oracle outputs for every vector x weight set.  NOT from
                 the reference (it records none): produced by oracle/sv_datapath.py,
                 whose two independent restatements agree; matches SURVEY Appendix D.
"""
from __future__ import annotations

import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from modulationdetectioncnn_b200 import svtext  # noqa: E402
from modulationdetectioncnn_b200.h5lite import H5File  # noqa: E402
from oracle import sv_datapath  # noqa: E402

REF = os.environ.get("MDC_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

VECTOR_FILES = [
    "12.15.testDataClass1.txt", "12.15testDataClass2.txt", "12.15.testDataClass3.txt",
    "12.15.newTestFirst.txt", "12.15.newTestSecond.txt", "12.15.newTestThird.txt",
    "12.15.newTestFourth.txt", "12.15.sixSampleData.txt", "12.15.sixtyfourSamples.txt",
    "12.14.testdata.class2.txt", "12.14.testdata.class3.txt", "newTestData.txt",
    "newTestDataClass2.txt", "newTestDataClass3.txt", "12.16.testDataYunyun.txt",
]
H5_FILES = {
    "E_f10": "convmodrecnets_CNN2_0.5.wts.h5",
    "B_2conv": "2convmodrecnets_CNN2_0.5.wts.h5",
    "A_3conv": "3convmodrecnets_CNN2_0.5.wts.h5",
    "D_4conv": "4convmodrecnets_CNN2_0.5.wts.h5",
    "C_5conv": "5convmodrecnets_CNN2_0.5.wts.h5",
}


def r(p):
    return os.path.join(REF, p)


def load_sets():
    A = svtext.load_qweights(r("12.15.latestWeights.txt"))
    B = svtext.load_qweights(r("12.15.denseWeights.txt"), conv_from=r("12.14.weights.txt"))
    C = svtext.load_qweights(r("am.fm.8psk.txt"))
    # am.fm.qpsk.txt has no dense-bias section: quantise the checkpoint's bias
    from modulationdetectioncnn_b200.fixedpoint import quantize
    h = read_h5(r(H5_FILES["D_4conv"]))
    D = svtext.load_qweights(r("am.fm.qpsk.txt"), dense_bias=quantize(h["dense_b"]).tolist())
    return {"A": A, "B": B, "C": C, "D": D}


def read_h5(path):
    f = H5File(path)
    names = [str(n) for n in f.attrs("/model_weights")["layer_names"]]
    out = {}
    for n in names:
        wn = f.attrs(f"/model_weights/{n}").get("weight_names", [])
        for w in wn:
            w = str(w)
            kind = "conv" if "conv" in w else "dense"
            part = "k" if "kernel" in w else "b"
            out[f"{kind}_{part}"] = f.dataset(f"/model_weights/{n}/{w}")
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    sets = load_sets()
    q = {}
    for k, w in sets.items():
        q[f"{k}_conv_tab"], q[f"{k}_dense_bias"], q[f"{k}_dense_tabs"] = w.conv_tab, w.dense_bias, w.dense_tabs
        q[f"{k}_n_overwide"] = np.int32(w.n_overwide)
    pf = svtext.parse_file(r("DenseWeights1.txt"))
    q["E_dense_flat"] = pf.tables[0].dense(7740)
    np.savez_compressed(os.path.join(OUT, "qweights.npz"), **q)

    names, vecs, vecs_zero = [], [], []
    for fn in VECTOR_FILES:
        v = svtext.load_vectors(r(fn))
        vz = svtext.load_vectors(r(fn), overwidth="zero")
        for i in range(v.shape[0]):
            names.append(fn if v.shape[0] == 1 else f"{fn}#{i}")
            vecs.append(v[i])
            vecs_zero.append(vz[i])
    np.savez_compressed(os.path.join(OUT, "vectors.npz"), names=np.array(names),
                        vectors=np.stack(vecs), vectors_zero=np.stack(vecs_zero))

    h5 = {}
    for tag, fn in H5_FILES.items():
        for k, a in read_h5(r(fn)).items():
            h5[f"{tag}_{k}"] = a
    np.savez_compressed(os.path.join(OUT, "h5_weights.npz"), **h5)

    roms = svtext.parse_sv_roms(r("cnn_test_latest1.sv"))
    np.savez_compressed(os.path.join(OUT, "sv_roms.npz"), **roms)

    # ---- what the reference records
    nb = json.load(open(r("CNN.ipynb")))
    cell18 = "".join(nb["cells"][18]["outputs"][0]["text"])
    body = cell18[: cell18.rindex("[3.47")]
    frame = [float(t) for t in re.findall(r"-?\d+\.\d*(?:e[-+]?\d+)?", body)]
    assert len(frame) == 256
    cell25 = "".join(nb["cells"][25]["outputs"][0]["text"])
    bits25 = re.findall(r"18'b([01]+);", cell25)
    kat = {
        "source": {
            "keras_dense_3samples": "12.16.testDataYunyun.txt:2",
            "keras_dense_64samples": "12.16.testDataYunyun.txt:264 ; CNN.ipynb cell 18",
            "cell18_frame": "CNN.ipynb cell 18 printed newTest1 (X_test[2000], samples 64..127 zeroed)",
            "float2fix_pairs": "CNN.ipynb cells 21 (floats) and 25 (bit strings)",
            "weights": "3convmodrecnets_CNN2_0.5.wts.h5 (CNN.ipynb cell 8)",
        },
        "keras_dense_3samples": [0.0, 3.1391976, 0.3649335],
        "keras_dense_64samples": [3.4700375, 2.4710786, 1.3579643],
        "vector_index_3samples": names.index("12.16.testDataYunyun.txt#0"),
        "vector_index_64samples": names.index("12.16.testDataYunyun.txt#1"),
        "cell18_frame": frame,
        "float2fix_pairs": [[3.510959, bits25[0]], [3.1282985, bits25[1]], [4.310499, bits25[2]]],
        "class_order": ["WBFM", "AM-SSB", "GFSK"],
        "eval_loss_3conv": 0.5455338358879089,
    }
    json.dump(kat, open(os.path.join(OUT, "kat.json"), "w"), indent=1)

    # ---- integer goldens from the oracle (unpinned)
    V = np.stack(vecs)
    g = {"note": "oracle/sv_datapath.py outputs; the reference records no SV output (parity unpinned)",
         "names": names, "pre": {}, "testbench_vector": {}}
    for k, w in sets.items():
        g["pre"][k] = sv_datapath.forward_pre(V, w.conv_tab, w.dense_bias, w.dense_tabs).tolist()
    tb = np.zeros(256, dtype=np.int32)
    tb[: roms["test_table"].shape[0]] = roms["test_table"]
    A = sets["A"]
    out, pre, info = sv_datapath.simulate_rtl(tb, A.conv_tab, A.dense_bias, A.dense_tabs)
    g["testbench_vector"] = {"nonzero": {str(i): int(v) for i, v in enumerate(tb) if v},
                             "pre": pre.tolist(), "out": out.tolist(), "cycles": info["cycles"]}
    json.dump(g, open(os.path.join(OUT, "int_goldens.json"), "w"), indent=1)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
