"""TEST INFRASTRUCTURE - CPU oracle for the raw RTL-SDR ingest.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

**Parity unpinned**: the reference has no ingest code; /root/reference/README.md:5 only names the
RTL-SDR as the sample source.  The definition (include/mdc.h) is the usual rtl_sdr byte format:
interleaved unsigned 8-bit I/Q centred on 127.5, value = (u - 127.5) / 128, whose Q6.12 integer
(2u - 255) * 16 is exact - so the integer outputs also equal ``float2fix(value, 18, 12)`` of
/root/reference/CNN.ipynb:1 cell 23.
"""
from __future__ import annotations

import numpy as np

__all__ = ["ingest_u8"]


def ingest_u8(iq: np.ndarray):
    """uint8 [2n] -> (f32 [n/128,2,128], q612 int32 [n/128,256], fwht int32 [n/1024,2,1024] or None)."""
    u = np.asarray(iq, dtype=np.uint8).reshape(-1, 2).astype(np.int64)
    n = u.shape[0]
    if n % 128:
        raise ValueError("n_samples must be a multiple of 128")
    q = (2 * u - 255) * 16                                   # [n, (I,Q)]
    frames = q.reshape(n // 128, 128, 2).transpose(0, 2, 1)  # [frame, row, t]
    f32 = (frames.astype(np.float64) / 4096.0).astype(np.float32)
    q612 = frames.reshape(n // 128, 256).astype(np.int32)
    fwht = q.reshape(n // 1024, 1024, 2).transpose(0, 2, 1).astype(np.int32).copy() if n % 1024 == 0 else None
    return f32, q612, fwht
