"""TEST INFRASTRUCTURE / CPU BASELINE - the float nets on host cores with torch-CPU.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.

The reference runs ``model.predict`` on Keras 2.4 / TensorFlow-CPU (Eigen/oneDNN kernels,
/root/reference/cnn.py:198,237).  Keras and TensorFlow are not installable here, so the
closest available analogue - the same layer stack on torch's oneDNN/MKL CPU kernels with
all host threads - is the timed stand-in (kind "port").  It is validated against the numpy
restatement (oracle/cnn2_float.py) in tests/test_oracle_float.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

__all__ = ["TinyCNN2Cpu", "VTCNN2Cpu"]


class TinyCNN2Cpu:
    """Reshape -> ZeroPad(1) -> Conv2D(F,(1,2),relu) -> Flatten(channels_last) -> Dense(C,relu) -> softmax
    (/root/reference/CNN.ipynb:1 cell 6)."""

    def __init__(self, conv_k, conv_b, dense_k, dense_b):
        Fn = conv_k.shape[-1]
        self.k = torch.from_numpy(np.ascontiguousarray(conv_k.reshape(2, Fn).T)).reshape(Fn, 1, 1, 2).float()
        self.b = torch.from_numpy(np.asarray(conv_b)).float()
        # Keras flatten order (r,p,f) -> torch conv output order (f,r,p)
        d = np.asarray(dense_k).reshape(2, 129, Fn, -1).transpose(2, 0, 1, 3).reshape(2 * 129 * Fn, -1)
        self.d = torch.from_numpy(np.ascontiguousarray(d)).float()
        self.db = torch.from_numpy(np.asarray(dense_b)).float()

    @torch.no_grad()
    def predict(self, x: np.ndarray, batch_size: int = 1024, output: str = "softmax") -> np.ndarray:
        outs = []
        xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).reshape(-1, 1, 2, 128)
        for s in range(0, xt.shape[0], batch_size):
            y = F.relu(F.conv2d(F.pad(xt[s:s + batch_size], (1, 1)), self.k, self.b))
            z = F.relu(y.flatten(1) @ self.d + self.db)
            outs.append(z if output == "dense" else torch.softmax(z, -1))
        return torch.cat(outs).numpy()


class VTCNN2Cpu:
    """pad2 -> Conv(256,1x3,relu) -> pad2 -> Conv(80,2x3,relu) -> Flatten -> Dense(256,relu) -> Dense(C) -> softmax
    (example notebook :231-243).  Weights in Keras layouts, flatten channels_last."""

    def __init__(self, w1, b1, w2, b2, w3, b3, w4, b4):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()  # noqa: E731
        self.w1 = t(np.asarray(w1).transpose(3, 2, 0, 1))        # (256,1,1,3)
        self.w2 = t(np.asarray(w2).transpose(3, 2, 0, 1))        # (80,256,2,3)
        self.b1, self.b2, self.b3, self.b4 = t(b1), t(b2), t(b3), t(b4)
        # Keras flatten (pos, ch) -> torch conv output (ch, pos)
        self.w3 = t(np.asarray(w3).reshape(132, 80, 256).transpose(1, 0, 2).reshape(10560, 256))
        self.w4 = t(w4)

    @torch.no_grad()
    def predict(self, x: np.ndarray, batch_size: int = 1024, output: str = "softmax") -> np.ndarray:
        outs = []
        xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).reshape(-1, 1, 2, 128)
        for s in range(0, xt.shape[0], batch_size):
            a = F.relu(F.conv2d(F.pad(xt[s:s + batch_size], (2, 2)), self.w1, self.b1))
            c = F.relu(F.conv2d(F.pad(a, (2, 2)), self.w2, self.b2))
            h = F.relu(c.flatten(1) @ self.w3 + self.b3)
            z = h @ self.w4 + self.b4
            outs.append(z if output == "logits" else torch.softmax(z, -1))
        return torch.cat(outs).numpy()
