"""GPU parity: raw RTL-SDR ingest (mdc_sdr_ingest_u8) vs the numpy oracle - bit-exact (integers, and floats that
are exact multiples of 2^-12) - and its hand-off into predict / FWHT."""
import numpy as np
import pytest

from conftest import philox

pytestmark = pytest.mark.gpu


def _ingest(raw, want):
    import torch
    from modulationdetectioncnn_b200 import sdr
    out = sdr.ingest_u8(torch.from_numpy(raw).cuda(), want)
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("n", [128, 1024, 1024 * 37, 128 * 1001])
def test_ingest_matches_oracle(n):
    from oracle import sdr as osdr
    raw = philox(n).integers(0, 256, 2 * n, dtype=np.uint8)
    raw[:8] = [0, 255, 255, 0, 127, 128, 128, 127]              # extremes and the two mid codes
    f32, q612, fw = osdr.ingest_u8(raw)
    want = ("f32", "q612", "fwht") if n % 1024 == 0 else ("f32", "q612")
    got = _ingest(raw, want)
    assert np.array_equal(got["f32"], f32) and np.array_equal(got["q612"], q612)
    if n % 1024 == 0:
        assert np.array_equal(got["fwht"], fw)
    # subset requests leave the others untouched and agree
    assert np.array_equal(_ingest(raw, ("q612",))["q612"], q612)



def test_ingest_feeds_predict_and_fwht(golden, qsets, h5w):
    import torch
    from modulationdetectioncnn_b200 import sdr
    from modulationdetectioncnn_b200.fwht import fwht
    from modulationdetectioncnn_b200.model import tiny_cnn2
    from modulationdetectioncnn_b200.qmodel import FixedPointCNN2
    from modulationdetectioncnn_b200.svtext import QWeights
    from oracle import cnn2_float as cf, fwht as ofw, sdr as osdr, sv_datapath as sv
    n = 1024 * 16
    raw = (127.5 + 6 * philox(9).normal(size=2 * n)).clip(0, 255).astype(np.uint8)   # weak signal around mid-scale
    f32, q612, fw = osdr.ingest_u8(raw)
    out = sdr.ingest_u8(torch.from_numpy(raw).cuda(), ("f32", "q612", "fwht"))
    qm = FixedPointCNN2(3, 3)
    qm.set_tables(QWeights(*[a.copy() for a in qsets["A"]]))
    assert np.array_equal(qm.predict(out["q612"], output="pre").cpu().numpy(), sv.forward_pre(q612, *qsets["A"]))
    tm = tiny_cnn2(3, 3)
    tm.set_weights(h5w["A_3conv"])
    z = tm.predict(out["f32"], output="dense").cpu().numpy()
    ref = cf.tiny_cnn2_forward(f32, *h5w["A_3conv"], output="dense")
    assert np.max(np.abs(z - ref)) <= 1e-5 * max(1.0, np.abs(ref).max())
    spec = fwht(out["fwht"]).cpu().numpy()
    assert np.array_equal(spec, ofw.fwht_matrix(fw))


def test_ingest_rejects_bad_sizes():
    import torch
    from modulationdetectioncnn_b200 import _lib, sdr
    with pytest.raises(ValueError):
        sdr.ingest_u8(torch.zeros(2 * 100, dtype=torch.uint8, device="cuda"), ("f32",))
    with pytest.raises(ValueError):
        sdr.ingest_u8(torch.zeros(2 * 128, dtype=torch.uint8, device="cuda"), ("fwht",))
    buf = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    lib = _lib.load()
    assert lib.mdc_sdr_ingest_u8(buf.data_ptr(), 100, buf.data_ptr(), None, None, None) != 0        # C ABI checks too
    assert lib.mdc_sdr_ingest_u8(buf.data_ptr(), 128, None, None, buf.data_ptr(), None) != 0
    assert b"1024" in lib.mdc_last_error()
    with pytest.raises(ValueError):
        sdr.ingest_u8(np.zeros(256, np.uint8), ("f32",))
    assert sdr.ingest_u8(torch.zeros(0, dtype=torch.uint8, device="cuda"), ("f32",))["f32"].shape == (0, 2, 128)
