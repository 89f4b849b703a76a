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


def test_f16x3_mode_within_1e5(vt):
    """fp16 hi/lo split on the tensor cores (three kind::f16 MMAs per product at the full 16-bit rate): the same bar
    as the fp32 CUDA-core mode and 3xTF32 - logits within 1e-5 of the largest logit of the fp64 oracle (north_star:
    "fp32 mode within 1e-5 of Keras")."""
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, ref = vt
    m = vt_cnn2(11, mode="f16x3")
    m.set_weights(_wlist(w))
    z = m.predict(x, output="dense")
    scale = np.abs(ref["logits"]).max(axis=-1, keepdims=True)
    err = np.abs(z - ref["logits"]) / scale
    assert err.max() < 1e-5, err.max()
    assert m._fallback is None                       # no range fallback was needed (8 frames have O(1) samples)
    p = m.predict(x)
    np.testing.assert_allclose(p, ref["softmax"], rtol=1e-4, atol=1e-7)
    assert np.array_equal(m.predict_classes(x), ref["logits"].argmax(-1))
    for n in (1, 3, 127, 129):
        assert np.array_equal(m.predict(x[:n], output="dense"), z[:n]), n
    assert m.class_histogram(x).sum() == x.shape[0]
    assert m.predict(np.zeros((0, 2, 128), np.float32)).shape == (0, 11)


def test_f16x3_range_fallback(vt):
    """Values outside the fp16 range: the C ABI reports MDC_ERR_RANGE, the facade reruns the batch in 3xTF32."""
    import ctypes as C
    from modulationdetectioncnn_b200 import _lib
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, _ = vt
    big = (x[:64] * np.float32(1e8)).astype(np.float32)          # samples ~1e6: conv1 activations leave the fp16 range
    m = vt_cnn2(11, mode="f16x3")
    m.set_weights(_wlist(w))
    t = vt_cnn2(11, mode="tf32x3")
    t.set_weights(_wlist(w))
    out = np.empty((64, 11), np.float32)
    rc = m._h._lib.mdc_predict_f32_host(m._h.ptr, big.ctypes.data, 64, None, out.ctypes.data, None, None)
    assert rc == _lib.ERR_RANGE
    assert b"fp16 range" in m._h._lib.mdc_last_error()
    want = t.predict(big, output="dense")
    assert np.array_equal(m.predict(big, output="dense"), want)              # host path, automatic rerun
    import torch
    assert np.array_equal(m.predict(torch.from_numpy(big).cuda(), output="dense").cpu().numpy(), want)   # device path
    assert np.array_equal(m.predict_async(big, output="dense").result(), want)
    # and the flag does not stick: the next in-range batch runs in f16x3 again
    z = m.predict(x[:64], output="dense")
    assert m._h.range_flags() == 0
    assert np.abs(z - t.predict(x[:64], output="dense")).max() < 1e-5 * np.abs(z).max()


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
@pytest.mark.parametrize("mode", ["bf16", "tf32x3", "f16x3"])
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


@pytest.mark.parametrize("mode", ["bf16", "tf32x3", "f16x3"])
def test_predict_async_stream_matches_sync(vt, mode):
    """Several host batches in flight (the next one's copies run under this one's kernels): every result equals the
    synchronous call's, in submission order, for ragged batch sizes around the chunk / pass boundaries."""
    import torch
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, _ = vt
    m = vt_cnn2(11, mode=mode)
    m.set_weights(_wlist(w))
    sizes = [300, 1, 19000] if mode == "tf32x3" else [300, 8192 + 5, 1, 2048, 33000]
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


# ---------------------------------------------------------------------------------------------------------------
# The persistent loops at the BASELINE batch size: 65,536 frames are ~22,700 (bf16) / ~68,700 (split modes) conv
# super-tiles and 256 / 512 dense tiles on 148 CTAs, so every CTA wraps its operand ring, flips its TMEM buffers and
# reuses its frame buffers many times.
@pytest.fixture(scope="module")
def full_batch(vt):
    from modulationdetectioncnn_b200 import synth
    from oracle import cnn2_float as cf
    w, _, _ = vt
    n = 65536
    x = synth.iq_frames(n, seed=99)
    x[::1000] *= 48                       # some frames with O(1) samples
    idx = np.sort(philox(5).choice(n, 4096, replace=False))
    idx[:3] = (0, 1, 2)
    idx[-3:] = (n - 3, n - 2, n - 1)
    ref = cf.vt_cnn2_forward(x[idx], **w, output="logits")
    return x, idx, ref


@pytest.mark.parametrize("mode,tol", [("bf16", 2e-2), ("f16x3", 1e-5), ("tf32x3", 1e-5)])
def test_full_batch_sampled_oracle(vt, full_batch, mode, tol):
    """65,536 frames in one call; 4,096 sampled frames against the fp64 oracle, and frame independence: the same
    frames give the same bits wherever they sit in the batch (device path, host path)."""
    import torch
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, _, _ = vt
    x, idx, ref = full_batch
    m = vt_cnn2(11, mode=mode)
    m.set_weights(_wlist(w))
    xd = torch.from_numpy(x).cuda()
    z = m.predict(xd, output="dense").cpu().numpy()
    scale = np.abs(ref).max(axis=-1, keepdims=True)
    err = np.abs(z[idx] - ref) / scale
    assert err.max() < tol, (mode, err.max())
    assert np.array_equal(m.predict(x[idx], output="dense"), z[idx])          # independent of position, host == device
    hist = m.class_histogram(xd).cpu().numpy()
    assert np.array_equal(hist, np.bincount(z.argmax(-1), minlength=11))
    assert m._fallback is None


@pytest.mark.parametrize("mode,passes", [("bf16", 60), ("f16x3", 50), ("tf32x3", 12)])
def test_repeated_passes_are_bit_identical(vt, mode, passes):
    """Race check for the cross-CTA barrier protocol of the tensor-core kernels (cta-scope arrives on the peer's
    mbarriers, multicast commits): the same 65,536 frames, many passes, every pass bit-identical to the first (which
    test_full_batch_sampled_oracle ties to the oracle)."""
    import torch
    from modulationdetectioncnn_b200 import _lib
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, _, _ = vt
    n = 65536
    m = vt_cnn2(11, mode=mode)
    m.set_weights(_wlist(w))
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((n, 2, 128), device="cuda", generator=g).mul_(2.0 ** -7)
    x[: n // 8] *= 64
    st = torch.cuda.current_stream().cuda_stream
    ref = torch.empty((n, 11), device="cuda")
    out = torch.empty_like(ref)
    _lib.check(m._h._lib.mdc_predict_f32(m._h.ptr, x.data_ptr(), n, None, ref.data_ptr(), None, None, st))
    torch.cuda.synchronize()
    assert torch.isfinite(ref).all()
    bad = 0
    for _ in range(passes):
        out.fill_(float("nan"))
        _lib.check(m._h._lib.mdc_predict_f32(m._h.ptr, x.data_ptr(), n, None, out.data_ptr(), None, None, st))
        bad += int(not torch.equal(out, ref))
    assert bad == 0, f"{bad} of {passes} passes differ from the first"


def _bf16_round(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def test_bf16_conv2_activations_match_bf16_emulation(vt):
    """Layer tap (`model3`-style, CNN.ipynb cell 17): the bf16 kernel's conv2 activations against a numpy model that
    rounds where the kernel rounds (conv1 -> bf16, bf16 weights, wide accumulate, -> bf16): 1 bf16 ulp, 2,048 frames."""
    import ctypes as C
    from modulationdetectioncnn_b200 import _lib, synth
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, _, _ = vt
    ws = _wlist(w)
    n = 2048
    x = synth.iq_frames(n, seed=7)
    x[: n // 8] *= 64
    m = vt_cnn2(11, mode="bf16")
    m.set_weights(ws)
    m.predict(x, output="dense")
    bits = np.empty((n * 132, 80), np.uint16)
    got = C.c_size_t()
    _lib.check(m._h._lib.mdc_debug_read(m._h.ptr, 0, bits.ctypes.data, bits.nbytes, C.byref(got)))
    assert got.value == bits.nbytes
    act = (bits.astype(np.uint32) << 16).view(np.float32).reshape(n, 132, 80)
    w1 = ws[0].reshape(3, 256).astype(np.float32)
    w2 = _bf16_round(ws[2]).reshape(1536, 80).astype(np.float64)
    wants = []
    for s in range(0, n, 256):
        xs = x[s:s + 256]
        k = xs.shape[0]
        xp = np.zeros((k, 2, 132), np.float32)
        xp[:, :, 2:130] = xs
        a = ws[1].astype(np.float32) + sum(xp[:, :, j:j + 130, None] * w1[j] for j in range(3))
        a = _bf16_round(np.maximum(a, 0))
        ap = np.zeros((k, 2, 134, 256), np.float32)
        ap[:, :, 2:132] = a
        cols = np.concatenate([ap[:, r, j:j + 132, :] for r in range(2) for j in range(3)], axis=-1).astype(np.float64)
        wants.append(_bf16_round(np.maximum(cols @ w2 + ws[3].astype(np.float64), 0).astype(np.float32)))
    want = np.concatenate(wants)
    d = np.abs(act - want)
    tol = 2.0 ** -7 * np.abs(want) + 1e-4 * np.abs(want).max()       # 1 bf16 ulp + cancellation slack
    bad = d > tol
    assert not bad.any(), (int(bad.sum()), np.argwhere(bad)[:5].tolist(), float(d.max()))


# ---------------------------------------------------------------------------------------------------------------
# Narrow input formats and the pageable-memory call the reference makes (cnn.py:198,237 pass ordinary ndarrays)
@pytest.mark.parametrize("mode", ["bf16", "f16x3"])
def test_raw_u8_and_int16_frames_equal_ingest_then_predict(vt, mode):
    """uint8 interleaved I/Q (RTL-SDR bytes) and int16 Q6.12 frames are converted inside the conv kernel's frame load:
    bit-identical to sdr.ingest_u8 -> predict(f32) / to predict(q / 4096), host, device and streaming paths."""
    import torch
    from modulationdetectioncnn_b200 import sdr
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, _, _ = vt
    m = vt_cnn2(11, mode=mode)
    m.set_weights(_wlist(w))
    n = 8192 + 300
    raw = philox(11).integers(0, 256, (n, 128, 2), dtype=np.uint8)
    f32 = sdr.ingest_u8(torch.from_numpy(raw.reshape(-1)).cuda(), ("f32",))["f32"]
    want = m.predict(f32, output="dense").cpu().numpy()
    assert np.array_equal(m.predict(raw, output="dense"), want)                                  # host u8
    assert np.array_equal(m.predict(torch.from_numpy(raw).cuda(), output="dense").cpu().numpy(), want)
    assert np.array_equal(m.predict_async(raw, output="dense").result(), want)
    q = philox(12).integers(-2000, 2000, (n, 2, 128)).astype(np.int16)
    wantq = m.predict((q.astype(np.float32) / np.float32(4096)), output="dense")
    assert np.array_equal(m.predict(q, output="dense"), wantq)
    assert np.array_equal(m.predict(torch.from_numpy(q).cuda(), output="dense").cpu().numpy(), wantq)
    assert np.array_equal(m.predict_classes(raw), want.argmax(-1))


def test_raw_formats_rejected_where_no_kernel_reads_them(h5w, vt):
    """Raw frames are read by the tensor-core VT kernels and the specialised TinyCNN2 kernels; the generic TinyCNN2
    kernel, the fp32 CUDA-core VT mode and wrong format ids refuse them before anything is launched."""
    from modulationdetectioncnn_b200 import _lib
    from modulationdetectioncnn_b200.model import tiny_cnn2, vt_cnn2
    m = tiny_cnn2(4, 2)                                           # no specialisation for this shape
    g = philox(3)
    m.set_weights([g.normal(0, 1, (1, 2, 1, 4)).astype(np.float32), np.zeros(4, np.float32),
                   g.normal(0, 1, (2 * 129 * 4, 2)).astype(np.float32), np.zeros(2, np.float32)])
    with pytest.raises(_lib.MdcError) as e:
        m.predict(np.zeros((4, 128, 2), np.uint8))
    assert e.value.code == -4
    w, _, _ = vt
    v = vt_cnn2(11, mode="fp32")
    v.set_weights(_wlist(w))
    with pytest.raises(_lib.MdcError) as e:
        v.predict(np.zeros((4, 256), np.int16))
    assert e.value.code == -4
    lib, h = v._h._lib, v._h
    assert lib.mdc_predict_raw_host(h.ptr, np.zeros(1024, np.uint8).ctypes.data, 3, 1, None, None, None, None) == -1


def test_pageable_and_pinned_host_buffers_agree(vt):
    """model.predict(ndarray) as cnn.py:198 calls it - ordinary pageable memory, staged through the library's pinned
    ring - gives the same bits as page-locked buffers, across chunk boundaries."""
    import torch
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, _ = vt
    for mode in ("bf16", "f16x3"):
        m = vt_cnn2(11, mode=mode)
        m.set_weights(_wlist(w))
        n = 40000 + 13
        xx = np.tile(x, (n // x.shape[0] + 1, 1, 1))[:n].copy()                  # pageable
        xp = torch.from_numpy(xx).pin_memory().numpy()                           # page-locked copy
        a = m.predict(xx, output="dense")
        b = m.predict(xp, output="dense")
        assert np.array_equal(a, b), mode
        assert np.array_equal(a[:300], a[300:600])
        assert np.array_equal(m.predict(xx), m.predict(xp))


def test_predict_async_returns_before_the_batch_is_done(vt):
    """The streaming call only enqueues: with pinned input and (library-allocated) pinned output it returns while
    the GPU is still working, so the next batch's copies can run under this batch's kernels."""
    import time
    import torch
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, _ = vt
    m = vt_cnn2(11, mode="f16x3")
    m.set_weights(_wlist(w))
    n = 4 * 65536                                                                # ~20 ms of kernels
    xp = torch.from_numpy(np.tile(x, (n // x.shape[0] + 1, 1, 1))[:n].copy()).pin_memory().numpy()
    for _ in range(2):
        m.predict_async(xp).result()                   # warm-up: work space, device slots, the pinned result block
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p = m.predict_async(xp)
    t_submit = time.perf_counter() - t0
    out = p.result()
    t_total = time.perf_counter() - t0
    assert out.shape == (n, 11)
    assert t_submit < 0.5 * t_total, (t_submit, t_total)


def test_reserve_makes_the_first_call_capturable(vt):
    """mdc_reserve sizes the work space and packs the weights, so even the FIRST mdc_predict_f32 only enqueues kernels
    and can be captured into a CUDA graph; without it a capturing first call is refused, not silently broken."""
    import torch
    from modulationdetectioncnn_b200 import _lib
    from modulationdetectioncnn_b200.model import vt_cnn2
    w, x, _ = vt
    xd = torch.from_numpy(x).cuda()
    n = x.shape[0]
    for mode in ("bf16", "f16x3"):
        m = vt_cnn2(11, mode=mode)
        m.set_weights(_wlist(w))
        want = m.predict(xd, output="dense").clone()
        m2 = vt_cnn2(11, mode=mode)
        m2.set_weights(_wlist(w))
        m2.reserve(n)
        out = torch.zeros((n, 11), device="cuda")
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                _lib.check(m2._h._lib.mdc_predict_f32(m2._h.ptr, xd.data_ptr(), n, None, out.data_ptr(), None, None,
                                                      s.cuda_stream))
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, want), mode


def test_model_refuses_a_tensor_on_another_device(vt):
    import torch
    from modulationdetectioncnn_b200.model import vt_cnn2
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    w, x, _ = vt
    m = vt_cnn2(11, mode="bf16", device=0)
    m.set_weights(_wlist(w))
    with pytest.raises(ValueError):
        m.predict(torch.from_numpy(x[:4]).to("cuda:1"))
