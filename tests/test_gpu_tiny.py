"""GPU parity: TinyCNN2 fp32 through the Keras-shaped facade vs the fp64 oracle and the Keras KATs.
Tolerance (north_star): 1e-5 relative."""
import numpy as np
import pytest

from conftest import philox

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _model(w):
    from modulationdetectioncnn_b200.model import tiny_cnn2
    m = tiny_cnn2(filters=w[0].shape[-1], classes=w[3].shape[0])
    m.set_weights(w)
    return m


def test_keras_recorded_output(golden, h5w):
    kat = golden["kat"]
    x = np.array(kat["cell18_frame"], dtype=np.float32).reshape(1, 2, 128)
    m = _model(h5w["A_3conv"])
    z = m.predict(x, output="dense")[0]                      # model2 of CNN.ipynb cell 17
    np.testing.assert_allclose(z, kat["keras_dense_64samples"], rtol=RTOL)
    p = m.predict(x)[0]
    e = np.exp(np.array(kat["keras_dense_64samples"]) - max(kat["keras_dense_64samples"]))
    np.testing.assert_allclose(p, e / e.sum(), rtol=RTOL)


@pytest.mark.parametrize("tag", ["A_3conv", "B_2conv", "C_5conv", "D_4conv", "E_f10"])
def test_checkpoints_vs_oracle(h5w, tag):
    from oracle import cnn2_float as cf
    w = h5w[tag]
    x = philox(2016).normal(0, 2 ** -7, (4099, 2, 128)).astype(np.float32)      # C2a input law
    x[:3] *= 300                                                                  # large activations too
    m = _model(w)
    z = m.predict(x, output="dense")
    zo = cf.tiny_cnn2_forward(x, *w, output="dense")
    np.testing.assert_allclose(z, zo, rtol=RTOL, atol=1e-6)
    p = m.predict(x, batch_size=1024)
    po = cf.tiny_cnn2_forward(x, *w)
    np.testing.assert_allclose(p, po, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(p.sum(-1), 1, atol=1e-6)
    cls = m.predict_classes(x)
    margin = np.sort(zo, axis=-1)
    clear = (margin[:, -1] - margin[:, -2]) > 1e-4
    assert np.array_equal(cls[clear], zo.argmax(-1)[clear])
    assert np.array_equal(cls, z.argmax(-1))                  # fused argmax == argmax of its own output
    assert m.class_histogram(x).tolist() == np.bincount(cls, minlength=3).tolist()


def test_load_model_from_golden_like_geometries():
    """Generic (non-specialised) geometries go through the runtime-shape kernel."""
    from modulationdetectioncnn_b200.model import tiny_cnn2
    from oracle import cnn2_float as cf
    g = philox(11)
    for F, C in ((1, 1), (4, 2), (7, 11), (16, 16)):
        w = [g.normal(0, 1, (1, 2, 1, F)).astype(np.float32), g.normal(0, 0.1, F).astype(np.float32),
             g.normal(0, 0.05, (258 * F, C)).astype(np.float32), g.normal(0, 0.1, C).astype(np.float32)]
        x = g.normal(0, 1, (257, 2, 128)).astype(np.float32)
        m = tiny_cnn2(F, C)
        m.set_weights(w)
        np.testing.assert_allclose(m.predict(x, output="dense"), cf.tiny_cnn2_forward(x, *w, output="dense"),
                                   rtol=RTOL, atol=1e-5)
        np.testing.assert_allclose(m.predict(x), cf.tiny_cnn2_forward(x, *w), rtol=1e-4, atol=1e-6)


def test_evaluate_and_device_path(h5w):
    import torch
    from oracle import cnn2_float as cf
    w = h5w["A_3conv"]
    g = philox(3)
    x = g.normal(0, 2 ** -7, (2048, 2, 128)).astype(np.float32)
    y = np.eye(3)[g.integers(0, 3, 2048)]
    m = _model(w)
    loss = m.evaluate(x, y, batch_size=1024, verbose=0)
    assert abs(loss - cf.categorical_crossentropy(cf.tiny_cnn2_forward(x, *w), y)) < 1e-5
    xt = torch.from_numpy(x).cuda()
    pt = m.predict(xt)
    assert pt.is_cuda and np.allclose(pt.cpu().numpy(), m.predict(x), rtol=0, atol=0)   # same kernel, same bits
    assert m.predict(np.zeros((0, 2, 128), np.float32)).shape == (0, 3)
    with pytest.raises(ValueError):
        m.set_weights(w[:3])


def test_reads_reference_h5_when_present(reference_dir, h5w):
    import os
    from modulationdetectioncnn_b200.model import load_model
    m = load_model(os.path.join(reference_dir, "3convmodrecnets_CNN2_0.5.wts.h5"))
    assert (m.kind, m.filters, m.classes) == ("tiny", 3, 3)
    x = philox(1).normal(0, 2 ** -7, (64, 2, 128)).astype(np.float32)
    assert np.array_equal(m.predict(x), _model(h5w["A_3conv"]).predict(x))


@pytest.mark.parametrize("tag", ["A_3conv", "E_f10"])
def test_large_batch_wraps_the_frame_ring(h5w, tag):
    """2^19 + 5 frames: every warp of the 148 x 14 grid runs ~63 four-frame passes, so its three-buffer bulk-copy ring
    wraps ~21 times, the mbarrier phases flip, and the last pass is ragged (one frame).  The batch is a 4,099-frame block
    repeated: every copy must give the bits of the first, which is checked against the fp64 oracle; plus ragged sizes
    around the pass width, and the class histogram (one atomic per class and warp)."""
    import torch
    from oracle import cnn2_float as cf
    w = h5w[tag]
    base = philox(2016).normal(0, 2 ** -7, (4099, 2, 128)).astype(np.float32)
    base[:3] *= 300
    n = (1 << 19) + 5
    x = torch.from_numpy(np.tile(base, (n // 4099 + 1, 1, 1))[:n].copy()).cuda()
    m = _model(w)
    z = m.predict(x, output="dense").cpu().numpy()
    np.testing.assert_allclose(z[:4099], cf.tiny_cnn2_forward(base, *w, output="dense"), rtol=RTOL, atol=1e-6)
    full = n // 4099
    assert np.array_equal(z[:full * 4099].reshape(full, 4099, 3), np.broadcast_to(z[:4099], (full, 4099, 3)))
    assert np.array_equal(z[full * 4099:], z[:n - full * 4099])
    cls = m.predict_classes(x).cpu().numpy()
    assert np.array_equal(cls, z.argmax(-1))
    assert m.class_histogram(x).cpu().numpy().tolist() == np.bincount(cls, minlength=3).tolist()
    for k in (1, 2, 3, 4, 5, 7, 57):
        assert np.array_equal(m.predict(x[:k], output="dense").cpu().numpy(), z[:k]), k


@pytest.mark.parametrize("tag", ["A_3conv", "E_f10"])
def test_back_to_back_launches_and_graph_replay(h5w, tag):
    """The kernel is launched with programmatic stream serialisation (the next launch's prologue runs under this
    launch's tail).  Dependent launches chained on one stream - each reading what another kernel has just written, all
    writing the same output - must see ordered data, and the launch must stay capturable in a CUDA graph."""
    import torch
    from modulationdetectioncnn_b200 import _lib
    from oracle import cnn2_float as cf
    w = h5w[tag]
    m = _model(w)
    lib, h = m._h._lib, m._h
    n = 20000
    base = philox(7).normal(0, 2 ** -7, (n, 2, 128)).astype(np.float32)
    want = cf.tiny_cnn2_forward(base, *w, output="dense")
    x = torch.empty((n, 2, 128), device="cuda")
    src = torch.from_numpy(base).cuda()
    out = torch.empty((n, 3), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(20):
        x.copy_(src * float(rep + 1))                     # a producer kernel right in front of the dependent launch
        _lib.check(lib.mdc_predict_f32(h.ptr, x.data_ptr(), n, None, out.data_ptr(), None, None, st))
        x.zero_()                                         # and a writer right behind it
    torch.cuda.synchronize()
    want20 = cf.tiny_cnn2_forward(base * np.float32(20), *w, output="dense")
    np.testing.assert_allclose(out.cpu().numpy(), want20, rtol=RTOL, atol=1e-5 * np.abs(want20).max(axis=1, keepdims=True).max())
    x.copy_(src)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(3):
                _lib.check(lib.mdc_predict_f32(h.ptr, x.data_ptr(), n, None, out.data_ptr(), None, None, s.cuda_stream))
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("tag", ["A_3conv", "E_f10"])
def test_raw_u8_and_int16_frames_equal_ingest_then_predict(h5w, tag):
    """uint8 interleaved I/Q (RTL-SDR bytes) and int16 Q6.12 frames are converted when a lane picks its samples out of
    the ring: bit-identical to sdr.ingest_u8 -> predict(f32) / to predict(q / 4096) on the host, device and streaming
    paths, incl. ragged sizes (the ring's last pass) and the sample left of a lane's four (lane 0: zero padding)."""
    import torch
    from modulationdetectioncnn_b200 import sdr
    w = h5w[tag]
    m = _model(w)
    n = 16384 + 4099
    raw = philox(11).integers(0, 256, (n, 128, 2), dtype=np.uint8)
    f32 = sdr.ingest_u8(torch.from_numpy(raw.reshape(-1)).cuda(), ("f32",))["f32"]
    want = m.predict(f32, output="dense").cpu().numpy()
    assert np.abs(want).max() > 0
    assert np.array_equal(m.predict(raw, output="dense"), want)                                  # host u8
    assert np.array_equal(m.predict(torch.from_numpy(raw).cuda(), output="dense").cpu().numpy(), want)
    assert np.array_equal(m.predict_async(raw, output="dense").result(), want)
    for k in (1, 2, 3, 5, 4099):
        assert np.array_equal(m.predict(raw[:k], output="dense"), want[:k]), k
    q = philox(12).integers(-2000, 2000, (n, 2, 128)).astype(np.int16)
    q[0, :, 0] = (-32768, 32767)
    wantq = m.predict((q.astype(np.float32) / np.float32(4096)), output="dense")
    assert np.array_equal(m.predict(q, output="dense"), wantq)
    assert np.array_equal(m.predict(torch.from_numpy(q).cuda(), output="dense").cpu().numpy(), wantq)
    assert np.array_equal(m.predict_classes(raw), want.argmax(-1))
    assert m.class_histogram(raw).tolist() == np.bincount(want.argmax(-1), minlength=3).tolist()
