"""Formats: float2fix restatement, SV-literal text fixtures, minimal HDF5 reader."""
import os

import numpy as np
import pytest

from modulationdetectioncnn_b200 import fixedpoint as fx
from modulationdetectioncnn_b200 import svtext
from modulationdetectioncnn_b200.h5lite import H5File, H5FormatError


def test_float2fix_recorded_pairs(golden):
    # CNN.ipynb cells 21 + 25: the only float2fix I/O pairs the reference records
    for val, bits in golden["kat"]["float2fix_pairs"]:
        assert fx.float2fix(np.float32(val), 18, 12) == bits
        assert fx.bits_to_int(bits) == int(val * 4096)


def test_float2fix_truncates_toward_zero_and_twos_complement():
    assert fx.float2fix(0.0) == "0" * 18
    assert fx.float2fix(1.0) == "000001000000000000"
    assert fx.float2fix(-1.0) == "111111000000000000"
    assert fx.bits_to_int(fx.float2fix(-2.0896616)) == -8559      # conv kernel of set A
    assert fx.bits_to_int(fx.float2fix(-0.00075721)) == -3
    assert fx.bits_to_int(fx.float2fix(0.9999)) == 4095
    assert fx.bits_to_int(fx.float2fix(-0.9999)) == -4095          # not -4096: truncation, not floor


def test_float2fix_tiny_negative_bug():
    s = fx.float2fix(-1e-5)
    assert s == "1100000000000000000" and len(s) == 19             # one bit too long
    assert fx.bits_to_int(s, overwidth="verilog") == -131072       # sized literal keeps low 18 bits
    assert fx.bits_to_int(s, overwidth="zero") == 0
    with pytest.raises(ValueError):
        fx.bits_to_int(s, overwidth="error")


def test_quantize_matches_scalar_path():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(0, 2, 2000), rng.normal(0, 1e-4, 200), [0.0, -0.0, 31.9997, -32.0]]).astype(np.float32)
    ref = np.array([fx.bits_to_int(fx.float2fix(v)) for v in x])
    assert np.array_equal(fx.quantize(x), ref)
    refz = np.array([fx.bits_to_int(fx.float2fix(v), overwidth="zero") for v in x])
    assert np.array_equal(fx.quantize(x, overwidth="zero"), refz)


def test_int_bits_roundtrip():
    for v in (-131072, -65536, -1, 0, 1, 29, 131071):
        assert fx.bits_to_int(fx.int_to_bits(v)) == v
    assert fx.wrap_signed(131072) == -131072
    assert np.array_equal(fx.wrap_signed(np.array([131072, -131073, 5])), [-131072, 131071, 5])


def test_sv_roms_equal_latest_weights(golden):
    # cnn_test_latest1.sv ROMs == 12.15.latestWeights.txt, entry for entry (2,334 integers)
    roms, q = golden["sv_roms"], golden["qweights"]
    assert np.array_equal(roms["rom_cov"], q["A_conv_tab"])
    assert np.array_equal(roms["dense_bias"], q["A_dense_bias"])
    names = ["rom_dense_i_class1", "rom_dense_q_class1", "rom_dense_i_class2",
             "rom_dense_q_class2", "rom_dense_i_class3", "rom_dense_q_class3"]
    for i, n in enumerate(names):
        assert np.array_equal(roms[n], q["A_dense_tabs"][i]), n
    assert q["A_dense_tabs"][5, 373] == -65536      # the hand-edited -16.0 (sv:3114)


@pytest.mark.parametrize("tag,setname", [("A_3conv", "A"), ("B_2conv", "B"), ("C_5conv", "C"), ("D_4conv", "D")])
def test_weight_dumps_are_quantised_checkpoints(golden, tag, setname):
    """table(c,row)[f*129+p] == float2fix(DenseKernel[row*387 + p*3 + f, c])."""
    h, q = golden["h5_weights"], golden["qweights"]
    ck, cb, dk, db = h[f"{tag}_conv_k"], h[f"{tag}_conv_b"], h[f"{tag}_dense_k"], h[f"{tag}_dense_b"]
    F = ck.shape[-1]
    conv = np.stack([fx.quantize(ck[0, 0, 0]), fx.quantize(ck[0, 1, 0]), fx.quantize(cb)], axis=1).reshape(-1)
    mism_conv = int((conv != q[f"{setname}_conv_tab"]).sum())
    tabs = q[f"{setname}_dense_tabs"]
    mism = 0
    for c in range(3):
        for row in range(2):
            want = fx.quantize(dk.reshape(2, 129, F, 3)[row, :, :, c].T.reshape(-1))   # [f*129 + p]
            mism += int((want != tabs[2 * c + row]).sum())
    # the only mismatches are hand edits / over-width-literal bug entries (Appendix A.4)
    assert mism_conv <= 1 and mism <= 4, (mism_conv, mism)
    if setname != "D":
        assert np.array_equal(fx.quantize(db), q[f"{setname}_dense_bias"])


def test_set_e_dense_table_layout(golden):
    """DenseWeights1.txt = [c][r][f][p] dump of the 10-filter checkpoint (Appendix C, set E)."""
    dk = golden["h5_weights"]["E_f10_dense_k"]          # (2580, 3), rows (r, p, f)
    want = fx.quantize(dk.reshape(2, 129, 10, 3).transpose(3, 0, 2, 1).reshape(-1))
    got = golden["qweights"]["E_dense_flat"]
    assert (want != got).sum() <= 8


def test_text_roundtrip(tmp_path, qsets):
    ct, db, dt = qsets["A"]
    qw = svtext.QWeights(ct.copy(), db.copy(), dt.copy())
    p = tmp_path / "w.txt"
    svtext.write_qweights(qw, str(p))
    back = svtext.load_qweights(str(p))
    assert np.array_equal(back.conv_tab, ct) and np.array_equal(back.dense_bias, db)
    assert np.array_equal(back.dense_tabs, dt)
    v = np.arange(-128, 128, dtype=np.int32) * 1000
    svtext.write_vector(v, str(tmp_path / "v.txt"), header="demo")
    assert np.array_equal(svtext.load_vectors(str(tmp_path / "v.txt"))[0], v)


def test_parser_grammar_and_overwidth_policy():
    text = """* Convolution Bias + Weights:

18'd00: data <= 18'b111101111010010001;
18'd01: data <= 18'b1100000000000000000;
18'd02: data <= 18'b000000000000011101;

* Dense Bias:
18'b000000001100100111  // first
18'b111111111100000010
first table
18'd000: data = 18'b000000000000000001;
18'd001: data = 18'b000000000000000010;
18'd000: data = 18'b000000000000000011;
"""
    pf = svtext.parse_text(text)
    assert [len(t.addrs) for t in pf.tables] == [3, 2, 1]
    assert pf.tables[0].values == [-8559, -131072, 29] and pf.n_overwide == 1
    assert pf.bare == [807, -254]
    assert svtext.parse_text(text, overwidth="zero").tables[0].values[1] == 0
    with pytest.raises(ValueError):
        svtext.parse_text(text, overwidth="error")


def test_reference_files_parse_to_goldens(reference_dir, golden):
    """The committed fixtures are exactly what the parsers read from the reference checkout."""
    r = lambda p: os.path.join(reference_dir, p)  # noqa: E731
    A = svtext.load_qweights(r("12.15.latestWeights.txt"))
    q = golden["qweights"]
    assert np.array_equal(A.conv_tab, q["A_conv_tab"]) and np.array_equal(A.dense_tabs, q["A_dense_tabs"])
    names = [str(n) for n in golden["vectors"]["names"]]
    for i, n in enumerate(names):
        fn, _, idx = n.partition("#")
        v = svtext.load_vectors(r(fn))
        assert np.array_equal(v[int(idx or 0)], golden["vectors"]["vectors"][i]), n
    assert len(names) == 16
    roms = svtext.parse_sv_roms(r("cnn_test_latest1.sv"))
    assert np.array_equal(roms["rom_cov"], golden["sv_roms"]["rom_cov"])
    # over-width literal census
    assert svtext.parse_file(r("newTestDataClass2.txt")).n_overwide == 25
    assert svtext.parse_file(r("am.fm.qpsk.txt")).n_overwide == 3


def test_h5_reader_on_reference_checkpoints(reference_dir, golden):
    from modulationdetectioncnn_b200.model import read_keras_weights
    files = {"E_f10": "convmodrecnets_CNN2_0.5.wts.h5", "A_3conv": "3convmodrecnets_CNN2_0.5.wts.h5",
             "B_2conv": "2convmodrecnets_CNN2_0.5.wts.h5", "D_4conv": "4convmodrecnets_CNN2_0.5.wts.h5",
             "C_5conv": "5convmodrecnets_CNN2_0.5.wts.h5"}
    for tag, fn in files.items():
        path = os.path.join(reference_dir, fn)
        f = H5File(path)
        a = f.attrs("/")
        assert a["keras_version"] == "2.4.0" and a["backend"] == "tensorflow"
        ws = read_keras_weights(path)
        F = 10 if tag == "E_f10" else 3
        assert [w.shape for w in ws] == [(1, 2, 1, F), (F,), (258 * F, 3), (3,)]
        h = golden["h5_weights"]
        for w, k in zip(ws, ("conv_k", "conv_b", "dense_k", "dense_b")):
            assert np.array_equal(w, h[f"{tag}_{k}"])
        assert "optimizer_weights" in f.listdir("/")


def test_h5_reader_rejects_garbage(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(H5FormatError):
        H5File(str(p))
