"""Integer oracle: RTL-level model vs closed form, and the recorded goldens."""
import numpy as np
import pytest

from oracle import sv_datapath as sv

# SURVEY.md Appendix D (survey-derived; reproduced here independently by two restatements)
APPENDIX_D = {
    "12.15.testDataClass1.txt": {"A": [-132941, -86162, -42693], "B": [-5060, 141251, -125148],
                                 "C": [-70158, -12297, -9544], "D": [-20737, 20708, -10573]},
    "12.15testDataClass2.txt": {"A": [29414, -183388, -31751], "B": [1113, 735704, -618780],
                                "C": [7390, -71928, -54253], "D": [-27274, 14504, -6933]},
    "12.16.testDataYunyun.txt#1": {"A": [13609, 9518, 2058], "B": [-215, -1549, 3491],
                                   "C": [15148, 9359, 6583], "D": [-25108, -6577, -1255]},
    "12.14.testdata.class2.txt": {"A": [2091853, -927715, 439042], "B": [-8724, 1211630, -1051023],
                                  "C": [1768539, 43507, 181545], "D": [-63284, -316821, -22366]},
}
APPENDIX_D_OUT_A = {
    "12.15.newTestFirst.txt": [0, 12216, 0], "12.15.newTestSecond.txt": [0, 12154, 0],
    "12.15.newTestThird.txt": [0, 7760, 56], "12.15.newTestFourth.txt": [0, 0, 0],
    "12.15.sixSampleData.txt": [0, 12307, 0], "12.15.sixtyfourSamples.txt": [13609, 9518, 2058],
    "newTestData.txt": [0, 0, 0], "newTestDataClass2.txt": [2350972, 0, 625186],
    "newTestDataClass3.txt": [15080, 16514, 12943],
}


def test_mult_slice_semantics():
    # floor, not truncation, when it fits
    assert sv.mult_slice(-1, 1, 0, 0) == -1           # -1/4096 -> floor = -1
    assert sv.mult_slice(4096, 4096, 0, 0) == 4096    # 1.0*1.0
    assert sv.mult_slice(-4096, 4096, 4096, 2048) == -2048
    # 36-bit wrap: (-2^17)^2 * 2 = 2^35 -> bit35 set, low bits 0 -> -2^17
    assert sv.mult_slice(-131072, -131072, -131072, -131072) == -131072
    # forced sign with dropped middle bits: 2^34 -> m[28:12]=0, m[35]=0 -> 0
    assert sv.mult_slice(-131072, -131072, 0, 0) == 0
    # 131071^2*2 = 34359214082 -> bits 28..12 = 0x1FFC0 with sign 0
    assert sv.mult_slice(131071, 131071, 131071, 131071) == ((2 * 131071 ** 2) >> 12) & 0x1FFFF


def test_dense_rom_address_skew():
    A = sv.dense_rom_address(3)
    assert A.shape == (3, 128)
    assert A[0, :4].tolist() == [0, 0, 1, 2] and A[0, 127] == 126
    assert A[1, 0] == 128 and A[1, 1] == 128 and A[2, 127] == 382
    used = set(A.reshape(-1).tolist())
    assert not used & {127, 255, 383, 384, 385, 386}


def test_testbench_vector(golden, qsets):
    """The 2-sample vector embedded in cnn_test_latest1.sv:130,133 through both models."""
    tb = golden["int_goldens"]["testbench_vector"]
    v = np.zeros(256, dtype=np.int64)
    for k, val in tb["nonzero"].items():
        v[int(k)] = val
    assert tb["nonzero"] == {"0": 11, "128": -32}
    out, pre, info = sv.simulate_rtl(v, *qsets["A"])
    assert pre.tolist() == [-1662, 12273, -1814] == tb["pre"]
    assert out.tolist() == [0, 12273, 0] == tb["out"]
    assert info["cycles"] == 649 == tb["cycles"]          # 12.98 us at 50 MHz
    assert sv.forward_pre(v, *qsets["A"])[0].tolist() == tb["pre"]


def test_goldens_all_vectors_all_sets(golden, qsets):
    V = golden["vectors"]["vectors"]
    names = golden["int_goldens"]["names"]
    for k, w in qsets.items():
        pre = sv.forward_pre(V, *w)
        assert pre.tolist() == golden["int_goldens"]["pre"][k]
        for n, want in APPENDIX_D.items():
            assert pre[names.index(n)].tolist() == want[k], (n, k)
    outA = sv.forward(V, *qsets["A"])
    for n, want in APPENDIX_D_OUT_A.items():
        assert outA[names.index(n)].tolist() == want, n


@pytest.mark.parametrize("setname", ["A", "D"])
def test_rtl_model_equals_closed_form_on_fixture_vectors(golden, qsets, setname):
    V = golden["vectors"]["vectors"]
    closed = sv.forward_pre(V, *qsets[setname])
    for i in range(V.shape[0]):
        _, pre, _ = sv.simulate_rtl(V[i], *qsets[setname])
        assert np.array_equal(pre, closed[i]), i


def test_rtl_model_equals_closed_form_fuzz(qsets):
    rng = np.random.default_rng(2016)
    for trial in range(12):
        k = "ABCD"[trial % 4]
        if trial < 8:      # full 18-bit range: exercises 36-bit wrap, forced sign, 18-bit bias wrap
            x = rng.integers(-(1 << 17), 1 << 17, 256)
        else:              # fixture-like magnitudes
            x = np.trunc(rng.normal(0, 32, 256)).astype(np.int64)
        out, pre, _ = sv.simulate_rtl(x, *qsets[k])
        assert np.array_equal(pre, sv.forward_pre(x, *qsets[k])[0])
        assert np.array_equal(out, sv.forward(x, *qsets[k])[0])


def test_rtl_model_random_tables_and_other_geometry():
    """F=4, C=2 with random full-range ROMs: the restatements agree beyond the shipped shape."""
    rng = np.random.default_rng(7)
    F, C = 4, 2
    conv = rng.integers(-(1 << 17), 1 << 17, 3 * F)
    bias = rng.integers(-(1 << 17), 1 << 17, C)
    tabs = rng.integers(-(1 << 17), 1 << 17, (2 * C, 129 * F))
    for _ in range(3):
        x = rng.integers(-(1 << 17), 1 << 17, 256)
        _, pre, info = sv.simulate_rtl(x, conv, bias, tabs)
        assert np.array_equal(pre, sv.forward_pre(x, conv, bias, tabs)[0])
        assert info["cycles"] == 649 + 129


def test_unused_rom_entries_do_not_matter(qsets):
    ct, db, dt = (a.copy() for a in qsets["A"])
    x = np.random.default_rng(3).integers(-2000, 2000, (4, 256))
    base = sv.forward_pre(x, ct, db, dt)
    for a in (127, 255, 383, 384, 385, 386):
        dt[:, a] = 12345
    assert np.array_equal(sv.forward_pre(x, ct, db, dt), base)


def test_deskewed_address_map_is_the_only_difference():
    """forward_pre and forward_deskewed_pre run the same arithmetic (sv._accumulate); only the ROM address map differs."""
    A, D = sv.dense_rom_address(3), sv.deskewed_rom_address(3)
    assert A.shape == (3, 128) and D.shape == (3, 129)
    assert D[0].tolist() == list(range(129)) and D[2, 128] == 386          # every table entry used exactly once
    assert sorted(D.reshape(-1).tolist()) == list(range(387))
    g = np.random.default_rng(1)
    x = g.integers(-(1 << 17), 1 << 17, (8, 256))
    ct = g.integers(-(1 << 17), 1 << 17, 9)
    db = g.integers(-(1 << 17), 1 << 17, 3)
    dt = g.integers(-(1 << 17), 1 << 17, (6, 387))
    assert np.array_equal(sv._accumulate(x, ct, db, dt, A), sv.forward_pre(x, ct, db, dt))
    assert np.array_equal(sv._accumulate(x, ct, db, dt, D), sv.forward_deskewed_pre(x, ct, db, dt))
    assert not np.array_equal(sv.forward_pre(x, ct, db, dt), sv.forward_deskewed_pre(x, ct, db, dt))


def test_deskewed_datapath_reproduces_recorded_keras_output(golden, qsets, h5w):
    """PIN of the integer primitives (slice, 18-bit bias wrap, ReLU, 32-bit accumulate) to a number the reference
    records.  With the dense ROM read at the address the tables were laid out for (129 f + p, all 129 positions) the
    integer datapath must be the Q6.12 quantisation of the Keras net, i.e. reproduce the Dense+ReLU output Keras
    printed for the two recorded frames (12.16.testDataYunyun.txt:2,264; CNN.ipynb cell 18) to within quantisation
    error.  The bound, per class, in output LSBs (2^-12):
        387 floors of the dense slice (each in (-1, 0])          -> 387
      + truncated dense weights, |w*4096 - wq| < 1 per entry    -> sum_p (yI + yQ) / 4096
      + conv-stage error |dy| < 2 + (|x_p| + |x_p+1|) / 4096 (two truncated weights, truncated bias, one floor)
        carried through the dense weights                        -> sum |wq| / 4096 * |dy|
    What is left unpinned after this is the skew itself, which is read directly off cnn_test_latest1.sv:336,351-378.
    One table entry is not a quantisation of the checkpoint: rom_dense_q_class3[373] was hand-edited to -16.0 (-65536;
    float2fix of the tiny negative h5 weight -1.38e-4 emits a 19-bit literal); the pin uses the value
    float2fix's formula gives for it (0) and checks the hand edit's effect separately."""
    from oracle import cnn2_float as cf
    kat = golden["kat"]
    vec = golden["vectors"]["vectors"]
    ct, db, dt = (np.array(a) for a in qsets["A"])
    w = h5w["A_3conv"]
    r, p, f, c = 1, 115, 2, 2                                    # table class3-Q, entry 129*2 + 115 = 373
    assert dt[2 * c + r][129 * f + p] == -65536
    true_w = float(w[2][r * 387 + p * 3 + f, c])
    assert -1 < true_w * 4096 < 0                                 # the float2fix corner case
    dt_q = dt.copy()
    dt_q[2 * c + r][129 * f + p] = 0
    for idx_key, out_key in (("vector_index_3samples", "keras_dense_3samples"),
                             ("vector_index_64samples", "keras_dense_64samples")):
        x = vec[kat[idx_key]]
        got = np.maximum(sv.forward_deskewed_pre(x, ct, db, dt_q)[0], 0).astype(np.float64)
        xf = (x.astype(np.float64) / 4096).reshape(1, 2, 128)
        z = cf.tiny_cnn2_forward(xf, *w, output="dense")[0] * 4096          # fp64 Keras restatement, same inputs
        y = cf.tiny_cnn2_forward(xf, *w, output="conv")[0] * 4096           # [2,129,3] conv outputs in LSBs
        xi = np.abs(x.astype(np.float64)).reshape(2, 128)
        xpad = np.zeros((2, 130))
        xpad[:, 1:129] = xi
        dy = 2 + (xpad[:, 0:129] + xpad[:, 1:130]) / 4096                   # [2,129]
        for cls in range(3):
            wabs = np.stack([np.abs(dt_q[2 * cls + rr].reshape(3, 129)).T for rr in range(2)])   # [2,129,3]
            bound = 387 + y.sum() / 4096 + (wabs / 4096 * dy[:, :, None]).sum() + 1
            assert abs(got[cls] - max(z[cls], 0)) <= bound, (out_key, cls, got[cls], z[cls], bound)
            assert bound < 0.15 * z.max()                                   # the worst case is 13 % of the largest output,
            assert abs(got[cls] - max(z[cls], 0)) < 0.025 * z.max()         # the error measured here is 2 % (-274 LSB)
        # and against the digits Keras printed (its inputs were the unquantised floats: ~3 digits, test_oracle_float)
        keras = np.array(kat[out_key]) * 4096
        assert np.abs(got - keras).max() < 0.05 * keras.max(), (got, keras)
    # the hand edit: frame #1 has zero samples at p = 115, so the conv output there is the ReLU'd conv bias (241) and
    # the deployed ROM moves class 3 by slice(241 * -65536) = -3856
    x = vec[kat["vector_index_64samples"]]
    d_rom = sv.forward_deskewed_pre(x, ct, db, dt)[0] - sv.forward_deskewed_pre(x, ct, db, dt_q)[0]
    assert d_rom.tolist() == [0, 0, (241 * -65536) >> 12]
