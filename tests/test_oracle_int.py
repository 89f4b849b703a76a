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
