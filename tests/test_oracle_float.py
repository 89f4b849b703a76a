"""Float oracle pinned on the Keras outputs the reference recorded."""
import numpy as np

from oracle import cnn2_float as cf


def test_keras_recorded_output_from_unquantised_frame(golden, h5w):
    """CNN.ipynb cell 18: model2.predict(newTest1) -> [3.4700375 2.4710786 1.3579643].
    The printed frame has ~6 significant digits, so agreement is ~1e-6, not exact."""
    kat = golden["kat"]
    x = np.array(kat["cell18_frame"], dtype=np.float64).reshape(1, 2, 128)
    assert np.all(x[0, :, 64:] == 0)
    z = cf.tiny_cnn2_forward(x, *h5w["A_3conv"], output="dense")[0]
    np.testing.assert_allclose(z, kat["keras_dense_64samples"], rtol=2e-6, atol=0)
    z32 = cf.tiny_cnn2_forward(x, *h5w["A_3conv"], output="dense", dtype=np.float32)[0]
    np.testing.assert_allclose(z32, kat["keras_dense_64samples"], rtol=5e-6, atol=0)


def test_keras_recorded_outputs_from_quantised_vectors(golden, h5w):
    """12.16.testDataYunyun.txt:2,264: inputs there are Q6.12-quantised, so only ~3 digits agree."""
    kat = golden["kat"]
    V = golden["vectors"]["vectors"]
    x3 = V[kat["vector_index_3samples"]].reshape(1, 2, 128) / 4096.0
    x64 = V[kat["vector_index_64samples"]].reshape(1, 2, 128) / 4096.0
    z3 = cf.tiny_cnn2_forward(x3, *h5w["A_3conv"], output="dense")[0]
    z64 = cf.tiny_cnn2_forward(x64, *h5w["A_3conv"], output="dense")[0]
    np.testing.assert_allclose(z3, kat["keras_dense_3samples"], atol=3e-3)
    np.testing.assert_allclose(z64, kat["keras_dense_64samples"], atol=0.1)
    # survey Appendix D float KATs (numpy restatement)
    np.testing.assert_allclose(z3, [0, 3.13992102, 0.36301438], atol=1e-7)
    np.testing.assert_allclose(cf.softmax(z64), [0.65046078, 0.26595495, 0.08358427], atol=1e-7)


def test_tiny_layers_and_relu_before_softmax(h5w):
    rng = np.random.default_rng(0)
    x = rng.normal(0, 2 ** -7, (16, 2, 128)).astype(np.float32)
    for tag, w in h5w.items():
        F = w[0].shape[-1]
        y = cf.tiny_cnn2_forward(x, *w, output="conv")
        assert y.shape == (16, 2, 129, F) and (y >= 0).all()
        z = cf.tiny_cnn2_forward(x, *w, output="dense")
        assert (z >= 0).all()
        p = cf.tiny_cnn2_forward(x, *w)
        np.testing.assert_allclose(p.sum(-1), 1, atol=1e-12)
        # direct definition, one frame
        ck, cb, dk, db = (a.astype(np.float64) for a in w)
        xp = np.zeros((2, 130)); xp[:, 1:129] = x[0]
        yy = np.maximum(xp[:, :129, None] * ck[0, 0, 0] + xp[:, 1:, None] * ck[0, 1, 0] + cb, 0)
        np.testing.assert_allclose(np.maximum(yy.reshape(-1) @ dk + db, 0), z[0], rtol=1e-12, atol=1e-12)


def test_vt_shapes_flatten_orders_and_direct_definition():
    w = cf.vt_cnn2_init(classes=11, seed=1602)
    assert sum(v.size for v in w.values()) == 2830427
    rng = np.random.default_rng(1)
    x = rng.normal(0, 2 ** -7, (3, 2, 128)).astype(np.float32)
    a = cf.vt_cnn2_forward(x, **w, output="conv1")
    c = cf.vt_cnn2_forward(x, **w, output="conv2")
    assert a.shape == (3, 2, 130, 256) and c.shape == (3, 132, 80)
    # conv2 from its definition at a few points
    w2 = w["w2"].astype(np.float64)
    ap = np.zeros((2, 134, 256)); ap[:, 2:132] = a[0]
    for u, o in ((0, 0), (5, 17), (131, 79), (64, 40)):
        s = sum(ap[r, u + j, :] @ w2[r, j, :, o] for r in range(2) for j in range(3)) + w["b2"][o]
        np.testing.assert_allclose(max(s, 0), c[0, u, o], rtol=1e-10, atol=1e-12)
    p = cf.vt_cnn2_forward(x, **w)
    assert p.shape == (3, 11)
    np.testing.assert_allclose(p.sum(-1), 1, atol=1e-12)
    # channels_first flatten == channels_last with permuted dense1 rows
    w3cf = w["w3"].reshape(132, 80, 256).transpose(1, 0, 2).reshape(10560, 256)
    p2 = cf.vt_cnn2_forward(x, **{**w, "w3": w3cf}, flatten="channels_first")
    np.testing.assert_allclose(p, p2, rtol=1e-12)


def test_crossentropy_matches_definition():
    p = np.array([[0.7, 0.2, 0.1], [1.0, 0.0, 0.0]])
    y = np.array([[1, 0, 0], [0, 1, 0]])
    want = -(np.log(0.7) + np.log(1e-7)) / 2
    assert abs(cf.categorical_crossentropy(p, y) - want) < 1e-12


def test_torch_cpu_standin_computes_the_same_nets(h5w):
    """The CPU baseline / reference arm of bench.py (oracle/cnn2_torch_cpu.py: torch conv2d + matmul, a second,
    independent statement of both layer stacks) agrees with the numpy restatement - for every shipped checkpoint and for
    VT-CNN2 with the benchmark's synthetic weights - so the timed stand-in is the function the GPU path is checked
    against."""
    from modulationdetectioncnn_b200 import synth
    from oracle import cnn2_float as cf
    from oracle.cnn2_torch_cpu import TinyCNN2Cpu, VTCNN2Cpu
    x = synth.iq_frames(96, seed=5)
    x[:8] *= 64
    for tag, w in h5w.items():
        m = TinyCNN2Cpu(*w)
        want = cf.tiny_cnn2_forward(x, *w, output="dense")
        np.testing.assert_allclose(m.predict(x, batch_size=32, output="dense"), want, rtol=1e-5,
                                   atol=1e-5 * np.abs(want).max(), err_msg=tag)
        np.testing.assert_allclose(m.predict(x), cf.tiny_cnn2_forward(x, *w), atol=2e-6, err_msg=tag)
    wv = synth.vt_cnn2_weights(11, 1602)
    v = VTCNN2Cpu(*wv)
    ref = cf.vt_cnn2_forward(x, **cf.vt_cnn2_init(11, 1602), output="logits")
    got = v.predict(x, batch_size=32, output="logits")
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    np.testing.assert_allclose(v.predict(x), cf.vt_cnn2_forward(x, **cf.vt_cnn2_init(11, 1602)), atol=1e-6)
