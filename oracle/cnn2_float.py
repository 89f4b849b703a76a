"""TEST INFRASTRUCTURE - CPU restatement of the reference's float CNN2 nets.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product path
(``modulationdetectioncnn_b200``) never does.

The arithmetic of the reference lives in un-vendored Keras 2.4.0 / TensorFlow
2.x (h5 root attrs ``keras_version=2.4.0``, ``backend=tensorflow``), which is
not installable here, so this file restates the layer stacks the reference
builds, with Keras-2/TF semantics (cross-correlation, channels_last):

* TinyCNN2(F, C): /root/reference/CNN.ipynb:1 (cell 6) and the ``model_config``
  embedded in each ``*.wts.h5``:
  Reshape(2,128,1) -> ZeroPadding2D((0,0),(1,1)) -> Conv2D(F,(1,2),relu,valid)
  -> Flatten -> Dense(C, relu) -> softmax           (note the ReLU before softmax)
* VT-CNN2(C): /root/reference/examples-master/modulation_recognition/
  RML2016.10a_VTCNN2_example.ipynb:231-243 (shapes :194-216):
  pad(0,2) -> Conv(256,1x3,relu) -> pad(0,2) -> Conv(80,2x3,relu) -> Flatten
  -> Dense(256,relu) -> Dense(C) -> softmax          (Dropout = identity)

Pinning: the TinyCNN2 restatement is checked in
``tests/test_oracle_float.py`` against the two Keras outputs the reference
recorded (12.16.testDataYunyun.txt:2,264; CNN.ipynb cells 18, 21).
VT-CNN2: **parity unpinned** - the reference ships no weights or outputs for it.

Weights are plain numpy arrays in Keras layouts: conv kernel (kh,kw,cin,cout),
dense kernel (in,out).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "softmax", "tiny_cnn2_forward", "vt_cnn2_forward", "vt_cnn2_init",
    "categorical_crossentropy",
]


def softmax(z: np.ndarray) -> np.ndarray:
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


def tiny_cnn2_forward(x, conv_k, conv_b, dense_k, dense_b, dtype=np.float64, output="softmax"):
    """x [N,2,128] -> [N,C].  output in {"softmax","dense","conv","argmax"}.

    conv_k (1,2,1,F), conv_b (F), dense_k (2*129*F, C), dense_b (C).
    Flatten is channels_last: index = r*129*F + p*F + f.
    """
    x = np.asarray(x, dtype=dtype)
    N = x.shape[0]
    k = np.asarray(conv_k, dtype=dtype).reshape(2, -1)       # [tap, F]
    F = k.shape[1]
    xp = np.zeros((N, 2, 130), dtype=dtype)
    xp[:, :, 1:129] = x
    y = (xp[:, :, 0:129, None] * k[0] + xp[:, :, 1:130, None] * k[1]
         + np.asarray(conv_b, dtype=dtype))
    y = np.maximum(y, 0)                                     # [N,2,129,F]
    if output == "conv":
        return y
    z = y.reshape(N, 2 * 129 * F) @ np.asarray(dense_k, dtype=dtype) + np.asarray(dense_b, dtype=dtype)
    z = np.maximum(z, 0)
    if output == "dense":
        return z
    if output == "argmax":
        return z.argmax(axis=-1).astype(np.int32)
    return softmax(z)


def vt_cnn2_forward(x, w1, b1, w2, b2, w3, b3, w4, b4, dtype=np.float64, output="softmax",
                    flatten="channels_last", chunk=256):
    """x [N,2,128] -> [N,C].

    w1 (1,3,1,256) b1 (256) | w2 (2,3,256,80) b2 (80) | w3 (10560,256) b3 (256)
    | w4 (256,C) b4 (C).  ``flatten``: "channels_last" (row = pos*80 + ch, the
    Keras-2/TF order of every checkpoint in the reference) or "channels_first"
    (row = ch*132 + pos, the Keras-1/Theano order of the example notebook).
    output in {"softmax","logits","dense1","conv2","conv1","argmax"}.
    """
    x = np.asarray(x, dtype=dtype)
    N = x.shape[0]
    w1 = np.asarray(w1, dtype=dtype).reshape(3, 256)
    w2 = np.asarray(w2, dtype=dtype).reshape(2 * 3 * 256, 80)   # K index = (r*3+j)*256 + ch
    b1, b2, b3, b4 = (np.asarray(b, dtype=dtype) for b in (b1, b2, b3, b4))
    w3 = np.asarray(w3, dtype=dtype)
    w4 = np.asarray(w4, dtype=dtype)
    outs = []
    for s in range(0, N, chunk):
        xs = x[s:s + chunk]
        n = xs.shape[0]
        xp = np.zeros((n, 2, 132), dtype=dtype)
        xp[:, :, 2:130] = xs
        a = b1 + sum(xp[:, :, j:j + 130, None] * w1[j] for j in range(3))
        a = np.maximum(a, 0)                                 # [n,2,130,256]
        if output == "conv1":
            outs.append(a)
            continue
        ap = np.zeros((n, 2, 134, 256), dtype=dtype)
        ap[:, :, 2:132] = a
        # im2col: [n,132,(r,j,ch)]
        cols = np.concatenate([ap[:, r, j:j + 132, :] for r in range(2) for j in range(3)], axis=-1)
        c = np.maximum(cols @ w2 + b2, 0)                    # [n,132,80]
        if output == "conv2":
            outs.append(c)
            continue
        if flatten == "channels_last":
            flat = c.reshape(n, 132 * 80)
        elif flatten == "channels_first":
            flat = c.transpose(0, 2, 1).reshape(n, 80 * 132)
        else:
            raise ValueError(flatten)
        h = np.maximum(flat @ w3 + b3, 0)
        if output == "dense1":
            outs.append(h)
            continue
        logits = h @ w4 + b4
        if output == "logits":
            outs.append(logits)
        elif output == "argmax":
            outs.append(logits.argmax(axis=-1).astype(np.int32))
        else:
            outs.append(softmax(logits))
    return np.concatenate(outs, axis=0)


def vt_cnn2_init(classes: int = 11, seed: int = 1602):
    """Synthetic seeded VT-CNN2 weights.

    conv: Glorot-uniform, dense: He-normal (the inits named in the example
    notebook :233-241), biases N(0, 0.01); numpy Philox(seed).  float32.
    """
    g = np.random.Generator(np.random.Philox(seed))

    def glorot(shape):
        kh, kw, cin, cout = shape
        lim = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
        return g.uniform(-lim, lim, size=shape).astype(np.float32)

    def he(shape):
        return (g.standard_normal(shape) * np.sqrt(2.0 / shape[0])).astype(np.float32)

    def bias(n):
        return (g.standard_normal(n) * 0.01).astype(np.float32)

    return dict(w1=glorot((1, 3, 1, 256)), b1=bias(256), w2=glorot((2, 3, 256, 80)), b2=bias(80),
                w3=he((10560, 256)), b3=bias(256), w4=he((256, classes)), b4=bias(classes))


def categorical_crossentropy(probs: np.ndarray, y_onehot: np.ndarray) -> float:
    """Mean categorical cross-entropy as Keras computes it for ``evaluate``
    (/root/reference/cnn.py:154): probabilities clipped to [1e-7, 1-1e-7]."""
    p = np.clip(np.asarray(probs, dtype=np.float64), 1e-7, 1 - 1e-7)
    return float(-(np.asarray(y_onehot, dtype=np.float64) * np.log(p)).sum(axis=-1).mean())
