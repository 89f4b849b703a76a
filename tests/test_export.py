"""Quantise/export tool and the 10-filter integer model (8f-1): h5 weights -> Q6.12 tables ->
reference-grammar text -> parser, against the committed goldens of /root/reference's own file pairs."""
import numpy as np
import pytest

from modulationdetectioncnn_b200 import export, svtext
from modulationdetectioncnn_b200 import fixedpoint as fx


@pytest.mark.parametrize("tag,setname", [("A_3conv", "A"), ("B_2conv", "B"), ("C_5conv", "C"), ("D_4conv", "D")])
def test_quantize_checkpoint_reproduces_reference_dumps(h5w, qsets, tag, setname):
    """The exporter regenerates the reference's weight files from its checkpoints; the only differences are the
    documented hand edits / over-width literals."""
    qw = export.quantize_checkpoint(h5w[tag])
    ct, db, dt = qsets[setname]
    assert qw.filters == 3 and qw.classes == 3
    assert int((qw.conv_tab != ct).sum()) <= 1
    assert int((qw.dense_tabs != dt).sum()) <= 4
    if setname != "D":                                   # am.fm.qpsk.txt has no dense-bias section
        assert np.array_equal(qw.dense_bias, db)


def test_exported_text_round_trips(tmp_path, h5w):
    qw = export.quantize_checkpoint(h5w["A_3conv"])
    p = tmp_path / "a.Weights.txt"
    svtext.write_qweights(qw, str(p))
    back = svtext.load_qweights(str(p))
    assert np.array_equal(back.conv_tab, qw.conv_tab)
    assert np.array_equal(back.dense_bias, qw.dense_bias)
    assert np.array_equal(back.dense_tabs, qw.dense_tabs)
    text = p.read_text()
    assert "18'd00: data <= 18'b" in text and text.count("data <=") == 9 + 6 * 387


def test_export_vector_matches_recorded_fixture(tmp_path, golden):
    """A float frame quantised and written by the exporter parses back to the same integers, and quantising the
    de-quantised recorded vectors is the identity (the fixtures are fixed points of float2fix)."""
    V = golden["vectors"]["vectors"]
    ok = V != -(1 << 17)                                  # over-width-literal entries are not representable floats
    frames = (V.astype(np.float64) / 4096.0).reshape(-1, 2, 128)
    q = export.quantize_frame(frames)
    assert np.array_equal(q[ok], V[ok])
    p = tmp_path / "v.txt"
    v = export.export_vector(frames[0], str(p), header="demo")
    assert np.array_equal(svtext.load_vectors(str(p))[0], v)
    with pytest.raises(ValueError):
        export.quantize_frame(np.zeros((3, 128)))


def test_cli(tmp_path, h5w, capsys):
    f = tmp_path / "frame.npy"
    np.save(f, np.linspace(-0.01, 0.01, 256, dtype=np.float32).reshape(2, 128))
    assert export.main(["vector", str(f), str(tmp_path / "o.txt")]) == 0
    assert "256 entries" in capsys.readouterr().out
    assert export.main(["bogus"]) == 2


def test_set_e_model_tables(h5w, golden):
    """10-filter integer model: literal DenseWeights1.txt tables + conv table / dense bias quantised from the h5."""
    from oracle import sv_datapath as sv
    flat = golden["qweights"]["E_dense_flat"]
    qw = export.qweights_from_dense_dump(h5w["E_f10"], flat)
    assert qw.filters == 10 and qw.classes == 3 and qw.dense_tabs.shape == (6, 1290)
    assert np.array_equal(qw.dense_tabs.reshape(-1), flat)
    ref = export.quantize_checkpoint(h5w["E_f10"])
    assert int((ref.dense_tabs != qw.dense_tabs).sum()) <= 8           # same bound as test_formats (bug entries)
    # the integer model tracks the float model on small-signal frames where the ROM address skew matters least
    V = golden["vectors"]["vectors"]
    out = sv.forward(V, qw.conv_tab, qw.dense_bias, qw.dense_tabs)
    assert out.shape == (V.shape[0], 3) and (out >= 0).all()
    with pytest.raises(ValueError):
        export.qweights_from_dense_dump(h5w["E_f10"], flat[:-1])
