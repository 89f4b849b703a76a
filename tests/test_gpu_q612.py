"""GPU parity: the integer (SystemVerilog-exact) kernel through the C ABI vs the oracle.  Bit-exact."""
import numpy as np
import pytest

from conftest import philox

pytestmark = pytest.mark.gpu


def _model(qsets, k):
    from modulationdetectioncnn_b200.qmodel import FixedPointCNN2
    from modulationdetectioncnn_b200.svtext import QWeights
    m = FixedPointCNN2(3, 3)
    m.set_tables(QWeights(*[a.copy() for a in qsets[k]]))
    return m


@pytest.mark.parametrize("k", list("ABCD"))
def test_fixture_vectors_all_weight_sets(golden, qsets, k):
    from oracle import sv_datapath as sv
    V = golden["vectors"]["vectors"]
    m = _model(qsets, k)
    pre = m.predict(V, output="pre")
    out = m.predict(V, output="out")
    assert pre.tolist() == golden["int_goldens"]["pre"][k]
    assert np.array_equal(out, sv.forward(V, *qsets[k]))
    assert np.array_equal(m.predict(V, output="argmax"), out.argmax(-1))


def test_testbench_vector(golden, qsets):
    tb = golden["int_goldens"]["testbench_vector"]
    v = np.zeros((1, 256), dtype=np.int32)
    for a, val in tb["nonzero"].items():
        v[0, int(a)] = val
    m = _model(qsets, "A")
    assert m.predict(v, output="pre")[0].tolist() == tb["pre"]
    assert m.predict(v, output="out")[0].tolist() == tb["out"]


@pytest.mark.parametrize("k", list("ABCD"))
def test_fuzz_full_range_and_fixture_like(qsets, k):
    from oracle import sv_datapath as sv
    g = philox(2016)
    full = g.integers(-(1 << 17), 1 << 17, (3000, 256)).astype(np.int32)      # C1(iii): wrap paths
    small = np.trunc(philox(2015).normal(0, 32, (5000, 256))).astype(np.int32)  # C1(ii)
    edge = np.array([[-(1 << 17)] * 256, [(1 << 17) - 1] * 256, [0] * 256], dtype=np.int32)
    m = _model(qsets, k)
    for x in (full, small, edge):
        assert np.array_equal(m.predict(x, output="pre"), sv.forward_pre(x, *qsets[k]))


def test_random_tables_full_range():
    """Random full-range ROMs (incl. -2^17): every wrap path of slice/bias-add/accumulate."""
    from modulationdetectioncnn_b200.qmodel import FixedPointCNN2
    from modulationdetectioncnn_b200.svtext import QWeights
    from oracle import sv_datapath as sv
    g = philox(5)
    for F, C in ((3, 3), (10, 3), (4, 2), (1, 1), (16, 16)):
        conv = g.integers(-(1 << 17), 1 << 17, 3 * F).astype(np.int32)
        bias = g.integers(-(1 << 17), 1 << 17, C).astype(np.int32)
        tabs = g.integers(-(1 << 17), 1 << 17, (2 * C, 129 * F)).astype(np.int32)
        tabs[0, 0] = -(1 << 17)
        x = g.integers(-(1 << 17), 1 << 17, (700, 256)).astype(np.int32)
        x[0] = -(1 << 17)
        m = FixedPointCNN2(F, C)
        m.set_tables(QWeights(conv, bias, tabs))
        assert np.array_equal(m.predict(x, output="pre"), sv.forward_pre(x, conv, bias, tabs)), (F, C)
        assert np.array_equal(m.predict(x, output="out"), sv.forward(x, conv, bias, tabs)), (F, C)


def test_inputs_are_wrapped_to_18_bits(qsets):
    from oracle import sv_datapath as sv
    g = philox(9)
    x = g.integers(-(1 << 31), 1 << 31, (64, 256)).astype(np.int32)
    w18 = ((x.astype(np.int64) + (1 << 17)) % (1 << 18) - (1 << 17)).astype(np.int32)
    m = _model(qsets, "A")
    assert np.array_equal(m.predict(x, output="pre"), sv.forward_pre(w18, *qsets["A"]))


def test_empty_ragged_and_device_path(qsets):
    import torch
    from oracle import sv_datapath as sv
    m = _model(qsets, "A")
    assert m.predict(np.zeros((0, 256), dtype=np.int32)).shape == (0, 3)
    assert m.class_histogram(np.zeros((0, 256), dtype=np.int32)).tolist() == [0, 0, 0]
    for n in (1, 31, 33, 16385, 40001):          # around warp / host-chunk boundaries
        x = np.trunc(philox(n).normal(0, 300, (n, 256))).astype(np.int32)
        want = sv.forward(x, *qsets["A"])
        assert np.array_equal(m.predict(x), want)
        xt = torch.from_numpy(x).cuda()
        assert np.array_equal(m.predict(xt).cpu().numpy(), want)
        h = m.class_histogram(xt).cpu().numpy()
        assert h.tolist() == np.bincount(want.argmax(-1), minlength=3).tolist()
        assert m.class_histogram(x).tolist() == h.tolist()


def test_full_size_properties(qsets):
    """BASELINE size (N = 2^22 frames, 4 GiB): properties that need no oracle, plus a sampled check."""
    import torch
    from oracle import sv_datapath as sv
    n = 1 << 22
    gen = torch.Generator(device="cuda").manual_seed(2015)
    x = torch.randn((n, 256), generator=gen, device="cuda").mul_(32).trunc_().to(torch.int32)
    m = _model(qsets, "A")
    out = m.predict(x)
    hist = m.class_histogram(x).cpu().numpy()
    assert int(hist.sum()) == n                                           # conservation
    assert torch.equal(out, m.predict(x))                                 # deterministic
    assert hist.tolist() == torch.bincount(m.predict(x, output="argmax").long(), minlength=3).cpu().tolist()
    perm = torch.randperm(n, device="cuda", generator=gen)[: 1 << 16]
    assert torch.equal(m.predict(x[perm]), out[perm])                     # frames are independent
    idx = torch.arange(0, n, n // 4096, device="cuda")[:4096]
    assert np.array_equal(out[idx].cpu().numpy(), sv.forward(x[idx].cpu().numpy(), *qsets["A"]))
    assert (out >= 0).all()


def test_ten_filter_model_set_e(golden, h5w):
    """the 10-filter integer model (DenseWeights1.txt tables + conv table quantised from
    convmodrecnets_CNN2_0.5.wts.h5) through the same kernel family, bit-exact against the oracle."""
    from modulationdetectioncnn_b200 import export
    from modulationdetectioncnn_b200.qmodel import FixedPointCNN2
    from oracle import sv_datapath as sv
    qw = export.qweights_from_dense_dump(h5w["E_f10"], golden["qweights"]["E_dense_flat"])
    m = FixedPointCNN2(10, 3)
    m.set_tables(qw)
    V = golden["vectors"]["vectors"]
    full = philox(77).integers(-(1 << 17), 1 << 17, (2000, 256)).astype(np.int32)
    small = np.trunc(philox(78).normal(0, 32, (4000, 256))).astype(np.int32)
    for x in (V, full, small):
        want = sv.forward_pre(x, qw.conv_tab, qw.dense_bias, qw.dense_tabs)
        assert np.array_equal(m.predict(x, output="pre"), want)
        assert np.array_equal(m.predict(x, output="argmax"), np.maximum(want, 0).argmax(-1))
    assert int(m.class_histogram(small).sum()) == small.shape[0]


def _xfast(conv_tab, dense_tabs):
    """The input bound of the kernel's 32-bit path, derived as in mdc_set_weights_q612 (mdc_api.cu)."""
    w = np.abs(conv_tab.astype(np.int64)).reshape(-1, 3)
    csum, bmax = int((w[:, 0] + w[:, 1]).max()), int(w[:, 2].max())
    d = np.abs(dense_tabs.astype(np.int64)).reshape(dense_tabs.shape[0] // 2, 2, -1)
    dsum = int((d[:, 0] + d[:, 1]).max())
    lim, top = (1 << 28) - 1, (1 << 17) - 1
    xlim = lim // csum if csum else top
    ylim = min(lim // dsum if dsum else top, (1 << 16) - 2)
    return min(xlim, ((ylim - bmax - 1) * 4096) // csum if csum else top) if ylim > bmax + 1 else -1


@pytest.mark.parametrize("k", list("ABC"))
def test_small_signal_path_is_exact_up_to_its_bound(qsets, k):
    """Frames whose largest |x| is within the host-derived bound use 32-bit sums; frames just above it use the
    36-bit slices.  Both sides of the bound, with every sample AT the bound, must equal the oracle."""
    from oracle import sv_datapath as sv
    ct, db, dt = qsets[k]
    xf = _xfast(ct, dt)
    assert xf > 200, xf                                  # the shipped ROMs do have a small-signal regime
    m = _model(qsets, k)
    g = philox(31)
    for bound in (xf, xf + 1, xf // 2):
        x = g.integers(-bound, bound + 1, (3000, 256)).astype(np.int32)
        x[:500] = np.where(g.random((500, 256)) < 0.5, -bound, bound)      # all samples at the extreme
        x[500:1000, ::7] = bound
        assert np.array_equal(m.predict(x, output="pre"), sv.forward_pre(x, ct, db, dt)), (k, bound)


@pytest.mark.parametrize("geom", ["A", "E", "generic"])
def test_raw_u8_and_int16_frames(golden, qsets, h5w, geom):
    """int16 [N,256] (the test_table address map) and raw uint8 I/Q bytes are converted in the frame load: int16 equals
    the int32 call on the same values, uint8 equals the oracle on (2u - 255) * 16 (what sdr.ingest_u8 writes) - every
    kernel family (F=3, F=10, generic), host, device and streaming paths."""
    import torch
    from modulationdetectioncnn_b200 import export
    from modulationdetectioncnn_b200.qmodel import FixedPointCNN2
    from modulationdetectioncnn_b200.svtext import QWeights
    from oracle import sdr as osdr, sv_datapath as sv
    if geom == "A":
        qw = QWeights(*[a.copy() for a in qsets["A"]])
    elif geom == "E":
        qw = export.qweights_from_dense_dump(h5w["E_f10"], golden["qweights"]["E_dense_flat"])
    else:
        g = philox(5)
        qw = QWeights(g.integers(-9000, 9000, 3 * 4).astype(np.int32), g.integers(-3000, 3000, 2).astype(np.int32),
                      g.integers(-6000, 6000, (4, 129 * 4)).astype(np.int32))
    m = FixedPointCNN2(qw.filters, qw.classes)
    m.set_tables(qw)
    n = 16384 + 37
    q = philox(21).integers(-3000, 3000, (n, 256)).astype(np.int16)
    q[0, :4] = (-32768, 32767, -1, 0)                     # beyond the small-signal bound: the 36-bit slices
    want = sv.forward_pre(q.astype(np.int32), qw.conv_tab, qw.dense_bias, qw.dense_tabs)
    assert np.array_equal(m.predict(q, output="pre"), want)
    assert np.array_equal(m.predict(torch.from_numpy(q).cuda(), output="pre").cpu().numpy(), want)
    assert np.array_equal(m.predict_async(q, output="pre").result(), want)
    raw = philox(22).integers(0, 256, (n, 128, 2), dtype=np.uint8)
    _, q612, _ = osdr.ingest_u8(raw.reshape(-1))
    wantu = sv.forward_pre(q612, qw.conv_tab, qw.dense_bias, qw.dense_tabs)
    assert np.array_equal(m.predict(raw, output="pre"), wantu)
    assert np.array_equal(m.predict(torch.from_numpy(raw).cuda(), output="pre").cpu().numpy(), wantu)
    assert np.array_equal(m.predict(raw, output="argmax"), np.maximum(wantu, 0).argmax(-1))
    for k in (1, 31, 33):
        assert np.array_equal(m.predict(raw[:k], output="pre"), wantu[:k]), k
    assert m.class_histogram(raw).tolist() == np.bincount(np.maximum(wantu, 0).argmax(-1), minlength=qw.classes).tolist()
    from modulationdetectioncnn_b200 import _lib
    assert m._h._lib.mdc_predict_q612_raw_host(m._h.ptr, raw.ctypes.data, _lib.IN_F32, 1, None, None, None, None) == -1
