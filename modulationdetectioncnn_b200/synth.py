"""Synthetic weights and inputs for the configurations the reference has no data for.

VT-CNN2 has no checkpoint in the reference, so benchmark/parity runs use seeded
random-init weights of that architecture: conv
Glorot-uniform, dense He-normal - the initialisers named in
/root/reference/examples-master/modulation_recognition/RML2016.10a_VTCNN2_example.ipynb:233-241
- and biases N(0, 0.01), drawn from numpy Philox(seed).
"""
from __future__ import annotations

from typing import List

import numpy as np

__all__ = ["vt_cnn2_weights", "iq_frames", "q612_frames"]


def vt_cnn2_weights(classes: int = 11, seed: int = 1602) -> List[np.ndarray]:
    """[w1,b1,w2,b2,w3,b3,w4,b4] in Keras layouts, float32."""
    g = np.random.Generator(np.random.Philox(seed))

    def glorot(shape):
        kh, kw, cin, cout = shape
        lim = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
        return g.uniform(-lim, lim, size=shape).astype(np.float32)

    def he(shape):
        return (g.standard_normal(shape) * np.sqrt(2.0 / shape[0])).astype(np.float32)

    def bias(n):
        return (g.standard_normal(n) * 0.01).astype(np.float32)

    w1, b1 = glorot((1, 3, 1, 256)), bias(256)
    w2, b2 = glorot((2, 3, 256, 80)), bias(80)
    w3, b3 = he((10560, 256)), bias(256)
    w4, b4 = he((256, classes)), bias(classes)
    return [w1, b1, w2, b2, w3, b3, w4, b4]


def iq_frames(n: int, seed: int = 2016, sigma: float = 2.0 ** -7) -> np.ndarray:
    """float32 [n,2,128] i.i.d. N(0, sigma): RML2016.10a-like magnitudes (CNN.ipynb cell 16)."""
    g = np.random.Generator(np.random.Philox(seed))
    return (g.standard_normal((n, 2, 128), dtype=np.float32) * np.float32(sigma)).astype(np.float32)


def q612_frames(n: int, seed: int = 2015, sigma: float = 32.0) -> np.ndarray:
    """int32 [n,256] trunc(N(0, sigma)): the magnitude of the reference's test vectors (~+-30)."""
    g = np.random.Generator(np.random.Philox(seed))
    return np.trunc(g.standard_normal((n, 256)) * sigma).astype(np.int32)
