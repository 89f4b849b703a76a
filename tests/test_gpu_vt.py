"""GPU parity: VT-CNN2 through the facade vs the fp64 oracle (synthetic seeded weights -
the reference ships none; parity unpinned)."""
import numpy as np
import pytest

from conftest import philox

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vt():
    from oracle import cnn2_float as cf
    w = cf.vt_cnn2_init(classes=11, seed=1602)
    x = philox(2016).normal(0, 2 ** -7, (300, 2, 128)).astype(np.float32)
    x[:8] *= 64            # some frames with O(1) samples
    ref = {k: cf.vt_cnn2_forward(x, **w, output=k) for k in ("logits", "softmax")}
    return w, x, ref


def _wlist(w):
    return [w[k] for k in ("w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4")]


def test_fp32_mode_within_1e5(vt):
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, ref = vt
    m = vt_cnn2(11, mode="fp32")
    m.set_weights(_wlist(w))
    z = m.predict(x, output="dense")
    scale = np.abs(ref["logits"]).max(axis=-1, keepdims=True)
    assert np.max(np.abs(z - ref["logits"]) / scale) < 1e-5
    p = m.predict(x)
    np.testing.assert_allclose(p, ref["softmax"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(p.sum(-1), 1, atol=1e-6)
    assert np.array_equal(m.predict_classes(x), z.argmax(-1))
    assert m.class_histogram(x).sum() == x.shape[0]


def test_fp32_channels_first_flatten(vt):
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, ref = vt
    w3cf = w["w3"].reshape(132, 80, 256).transpose(1, 0, 2).reshape(10560, 256).copy()
    m = vt_cnn2(11, mode="fp32", flatten="channels_first")
    ws = _wlist(w)
    ws[4] = w3cf
    m.set_weights(ws)
    z = m.predict(x[:64], output="dense")
    scale = np.abs(ref["logits"][:64]).max(axis=-1, keepdims=True)
    assert np.max(np.abs(z - ref["logits"][:64]) / scale) < 1e-5


def test_tf32x3_mode_within_1e5(vt):
    """3xTF32 on the tensor cores: every fp32 operand split into tf32 hi + lo, three MMAs per product.
    Same bar as the fp32 CUDA-core mode: logits within 1e-5 of the largest logit of the fp64 oracle."""
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, ref = vt
    m = vt_cnn2(11, mode="tf32x3")
    m.set_weights(_wlist(w))
    z = m.predict(x, output="dense")
    scale = np.abs(ref["logits"]).max(axis=-1, keepdims=True)
    err = np.abs(z - ref["logits"]) / scale
    assert err.max() < 1e-5, err.max()
    p = m.predict(x)
    np.testing.assert_allclose(p, ref["softmax"], rtol=1e-4, atol=1e-7)
    assert np.array_equal(m.predict_classes(x), ref["logits"].argmax(-1))
    for n in (1, 3, 129):
        assert np.array_equal(m.predict(x[:n], output="dense"), z[:n]), n
    assert m.class_histogram(x).sum() == x.shape[0]


def test_bf16_mode_tolerance(vt):
    """bf16 operands, fp32 accumulate: logits within 2e-2 of the largest logit; argmax agrees
    wherever the oracle's top-2 margin exceeds that error."""
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, ref = vt
    m = vt_cnn2(11, mode="bf16")
    m.set_weights(_wlist(w))
    z = m.predict(x, output="dense")
    scale = np.abs(ref["logits"]).max(axis=-1, keepdims=True)
    err = np.abs(z - ref["logits"]) / scale
    assert err.max() < 2e-2, err.max()
    p = m.predict(x)
    assert np.abs(p - ref["softmax"]).max() < 2e-2
    srt = np.sort(ref["logits"], axis=-1)
    clear = (srt[:, -1] - srt[:, -2]) > 4e-2 * scale[:, 0]
    assert np.array_equal(m.predict_classes(x)[clear], ref["logits"].argmax(-1)[clear])


def test_bf16_ragged_sizes_and_determinism(vt):
    import torch
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, ref = vt
    m = vt_cnn2(11, mode="bf16")
    m.set_weights(_wlist(w))
    full = m.predict(x, output="dense")
    for n in (1, 2, 5, 127, 129, 300):
        part = m.predict(x[:n], output="dense")
        assert np.array_equal(part, full[:n]), n          # frames are independent: same bits
    xt = torch.from_numpy(x).cuda()
    assert np.array_equal(m.predict(xt, output="dense").cpu().numpy(), full)
    assert m.predict(np.zeros((0, 2, 128), np.float32)).shape == (0, 11)


@pytest.mark.parametrize("classes", [1, 3, 16])
@pytest.mark.parametrize("mode", ["bf16", "tf32x3"])
def test_other_class_counts(classes, mode):
    """The fused dense1 epilogue (bf16) and the head kernel (3xTF32) are specialised per class count (1..16)."""
    from modulationdetectioncnn_b200.model import vt_cnn2
    from oracle import cnn2_float as cf
    w = cf.vt_cnn2_init(classes=classes, seed=7)
    x = philox(3).normal(0, 2 ** -7, (130, 2, 128)).astype(np.float32)
    ref = cf.vt_cnn2_forward(x, **w, output="logits")
    m = vt_cnn2(classes, mode=mode)
    m.set_weights(_wlist(w))
    z = m.predict(x, output="dense")
    scale = np.abs(ref).max()
    if classes == 1:
        # a lone logit is a sum with heavy cancellation and has no larger neighbour to be measured against:
        # use the size of the terms it is made of
        h = cf.vt_cnn2_forward(x, **w, output="dense1")
        scale = (np.abs(h) @ np.abs(w["w4"])).max()
    err = np.abs(z - ref).max() / scale
    assert err < (2e-2 if mode == "bf16" else 1e-5), err
    p = m.predict(x)
    np.testing.assert_allclose(p.sum(-1), 1, atol=1e-5)
    assert int(m.class_histogram(x).sum()) == 130
    assert np.array_equal(m.predict_classes(x), z.argmax(-1))


def test_host_and_device_paths_agree_across_pass_boundaries(vt):
    """Host buffers go through the chunked copy/conv pipeline with 32,768-frame dense passes; device tensors through
    65,536-frame passes: same numbers, any batch size."""
    import torch
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, _ = vt
    m = vt_cnn2(11, mode="bf16")
    m.set_weights(_wlist(w))
    n = 32768 + 8192 + 77
    xx = np.tile(x, (n // x.shape[0] + 1, 1, 1))[:n].copy()
    zh = m.predict(xx, output="dense")
    zd = m.predict(torch.from_numpy(xx).cuda(), output="dense").cpu().numpy()
    assert np.array_equal(zh, zd)
    assert np.array_equal(zh[:300], zh[300:600])          # the same frames give the same rows wherever they sit
    assert int(m.class_histogram(xx).sum()) == n


@pytest.mark.parametrize("mode", ["bf16", "tf32x3"])
def test_predict_async_stream_matches_sync(vt, mode):
    """Several host batches in flight (the next one's copies run under this one's kernels): every result equals the
    synchronous call's, in submission order, for ragged batch sizes around the chunk / pass boundaries."""
    import torch
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, _ = vt
    m = vt_cnn2(11, mode=mode)
    m.set_weights(_wlist(w))
    sizes = [300, 8192 + 5, 1, 2048, 33000] if mode == "bf16" else [300, 1, 19000]
    batches = []
    for i, n in enumerate(sizes):
        b = np.tile(x, (n // x.shape[0] + 1, 1, 1))[:n] * (1.0 + 0.25 * i)
        batches.append(torch.from_numpy(b.astype(np.float32)).pin_memory().numpy())
    want = [m.predict(b, output="dense") for b in batches]
    pend = [m.predict_async(b, output="dense") for b in batches]          # all submitted before any wait
    for p, wnt in zip(pend, want):
        assert np.array_equal(p.result(), wnt)
    # interleaved submit / wait, other outputs, and a second result() call
    p0 = m.predict_async(batches[0])
    p1 = m.predict_async(batches[1], output="argmax")
    assert np.array_equal(p0.result(), m.predict(batches[0])) and np.array_equal(p0.result(), m.predict(batches[0]))
    assert np.array_equal(p1.result(), want[1].argmax(-1))


def test_predict_async_tiny_and_integer_models(h5w, qsets):
    """The chunked host pipeline of the TinyCNN2 / integer paths streams too: several batches in flight, same results."""
    import torch
    from modulationdetectioncnn_b200 import synth
    from modulationdetectioncnn_b200.model import tiny_cnn2
    from modulationdetectioncnn_b200.qmodel import FixedPointCNN2
    from modulationdetectioncnn_b200.svtext import QWeights
    m = tiny_cnn2(3, 3)
    m.set_weights(h5w["A_3conv"])
    q = FixedPointCNN2(3, 3)
    q.set_tables(QWeights(*[a.copy() for a in qsets["A"]]))
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    xs = [pin(synth.iq_frames(n, seed=n)) for n in (1000, 16384 * 2 + 3, 1, 40000)]
    qs = [pin(synth.q612_frames(n, seed=n)) for n in (500, 16384 + 1, 70000)]
    want = [m.predict(x) for x in xs]
    wantq = [q.predict(x, output="pre") for x in qs]
    pend = [m.predict_async(x) for x in xs]
    pendq = [q.predict_async(x, output="pre") for x in qs]
    for p, w in zip(pend + pendq, want + wantq):
        assert np.array_equal(p.result(), w)
    assert np.array_equal(q.predict_async(qs[0], output="argmax").result(), q.predict(qs[0], output="argmax"))


def test_device_predict_is_cuda_graph_capturable(vt):
    """The device-pointer call only enqueues kernels on the caller's stream (no allocation, copy or synchronisation
    after the first call has sized the work space), so a caller can capture it in a CUDA graph and replay it."""
    import torch
    from modulationdetectioncnn_b200 import _lib
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, _ = vt
    m = vt_cnn2(11, mode="bf16")
    m.set_weights(_wlist(w))
    xd = torch.from_numpy(x).cuda()
    out = torch.empty((x.shape[0], 11), device="cuda")
    hist = torch.zeros(11, dtype=torch.int64, device="cuda")
    lib, h = m._h._lib, m._h

    def call(stream):
        _lib.check(lib.mdc_predict_f32(h.ptr, xd.data_ptr(), x.shape[0], None, out.data_ptr(), None, hist.data_ptr(), stream))

    call(torch.cuda.current_stream().cuda_stream)          # sizes the work space, sets kernel attributes
    torch.cuda.synchronize()
    want = out.clone()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            call(s.cuda_stream)
    out.zero_()
    hist.zero_()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    assert int(hist.sum()) == 3 * x.shape[0]
