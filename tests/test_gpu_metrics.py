"""GPU parity: confusion matrix / per-SNR accuracy kernels against the reference's own Python loops
(cnn.py:200-218, 227-255 restated verbatim in numpy here - pure counting, bit-exact)."""
import numpy as np
import pytest

from conftest import philox

pytestmark = pytest.mark.gpu


def _loops(y_onehot, y_hat_cls, nclass):
    conf = np.zeros([nclass, nclass])
    for i in range(y_onehot.shape[0]):                 # cnn.py:205-211
        j = list(y_onehot[i, :]).index(1)
        k = int(y_hat_cls[i])
        conf[j, k] = conf[j, k] + 1
    return conf


@pytest.mark.parametrize("nclass,n", [(3, 1000), (11, 20000), (16, 5)])
def test_confusion_matches_reference_loop(nclass, n):
    import torch
    from modulationdetectioncnn_b200 import metrics
    g = philox(nclass)
    t = g.integers(0, nclass, n)
    p = g.integers(0, nclass, n)
    onehot = np.eye(nclass, dtype=np.float32)[t]
    want = _loops(onehot, p, nclass)
    got = metrics.confusion_matrix(torch.from_numpy(onehot).cuda(), torch.from_numpy(p).cuda(), nclass)
    assert np.array_equal(got.cpu().numpy(), want.astype(np.int64))
    got2 = metrics.confusion_matrix(torch.from_numpy(t).cuda(), torch.from_numpy(p).cuda(), nclass)
    assert np.array_equal(got2.cpu().numpy(), want.astype(np.int64))
    cn = metrics.normalize_rows(got)
    rows = want.sum(1) > 0
    np.testing.assert_allclose(cn[rows], want[rows] / want[rows].sum(1, keepdims=True))


def test_accuracy_by_snr_matches_reference_loop():
    import torch
    from modulationdetectioncnn_b200 import metrics
    g = philox(7)
    n, nclass = 30000, 11
    snrs = list(range(-20, 20, 2))
    snr = g.choice(snrs, n)
    t = g.integers(0, nclass, n)
    p = np.where(g.random(n) < (snr + 22) / 45.0, t, g.integers(0, nclass, n))      # better at high SNR
    acc = {}
    for s in snrs:                                        # cnn.py:231-254
        sel = np.where(snr == s)
        conf = _loops(np.eye(nclass)[t[sel]], p[sel], nclass)
        cor = np.sum(np.diag(conf))
        ncor = np.sum(conf) - cor
        acc[s] = 1.0 * cor / (cor + ncor)
    got = metrics.accuracy_by_snr(torch.from_numpy(t).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(snr).cuda(), nclass)
    assert sorted(got) == snrs
    for s in snrs:
        assert got[s] == pytest.approx(acc[s], abs=1e-12)
    some = metrics.accuracy_by_snr(torch.from_numpy(t).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(snr).cuda(), nclass,
                                   snrs=[0, 18])
    assert some == {0: pytest.approx(acc[0]), 18: pytest.approx(acc[18])}


def test_many_groups_and_bad_labels():
    import torch
    from modulationdetectioncnn_b200 import metrics
    g = philox(9)
    n, nclass, ng = 50000, 16, 300                        # 76,800 cells: the global-atomics path
    t = g.integers(-1, nclass + 1, n)                     # some labels out of range: skipped
    p = g.integers(0, nclass, n)
    grp = g.integers(0, ng, n)
    ok = (t >= 0) & (t < nclass)
    want = np.zeros((ng, nclass, nclass), np.int64)
    np.add.at(want, (grp[ok], t[ok], p[ok]), 1)
    got = metrics.confusion_matrix(torch.from_numpy(t).cuda(), torch.from_numpy(p).cuda(), nclass,
                                   groups=torch.from_numpy(grp).cuda(), n_groups=ng)
    assert np.array_equal(got.cpu().numpy(), want)
    with pytest.raises(ValueError):
        metrics.confusion_matrix(torch.from_numpy(t[:5]).cuda(), torch.from_numpy(p).cuda(), nclass)


def test_model_to_metrics_pipeline(h5w):
    """predict(argmax) on the device -> confusion matrix, as cnn.py does it for the whole test set."""
    import torch
    from modulationdetectioncnn_b200 import metrics, synth
    from modulationdetectioncnn_b200.model import tiny_cnn2
    from oracle import cnn2_float as cf
    x = synth.iq_frames(4096, seed=3) * 64
    m = tiny_cnn2(3, 3)
    m.set_weights(h5w["A_3conv"])
    cls = m.predict(torch.from_numpy(x).cuda(), output="argmax")
    ref = cf.tiny_cnn2_forward(x, *h5w["A_3conv"], output="dense").argmax(-1)
    y = philox(1).integers(0, 3, 4096)
    conf = metrics.confusion_matrix(torch.from_numpy(y).cuda(), cls, 3).cpu().numpy()
    want = np.zeros((3, 3), np.int64)
    np.add.at(want, (y, ref), 1)
    assert np.array_equal(conf, want) and conf.sum() == 4096
