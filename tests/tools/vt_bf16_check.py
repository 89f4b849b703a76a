"""GPU debugging aid: VT-CNN2 bf16 path, layer by layer, against a bf16-emulating numpy model.

    python tests/tools/vt_bf16_check.py [n_frames]

Uses oracle/ only as the checker (this is a test tool, not a product path).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from modulationdetectioncnn_b200 import _lib, synth          # noqa: E402
from modulationdetectioncnn_b200.model import vt_cnn2          # noqa: E402
from oracle import cnn2_float as cf                            # noqa: E402


def bf16_round(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def bf16_bits_to_f32(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def emulate(x, w):
    """conv1 fp32 -> bf16; conv2 (bf16 weights, wide accumulate) -> +b2, relu -> bf16; dense1 likewise."""
    n = x.shape[0]
    w1 = w[0].reshape(3, 256).astype(np.float32)
    xp = np.zeros((n, 2, 132), np.float32)
    xp[:, :, 2:130] = x
    a = w[1].astype(np.float32) + sum(xp[:, :, j:j + 130, None] * w1[j] for j in range(3))
    a = bf16_round(np.maximum(a, 0))
    ap = np.zeros((n, 2, 134, 256), np.float32)
    ap[:, :, 2:132] = a
    cols = np.concatenate([ap[:, r, j:j + 132, :] for r in range(2) for j in range(3)], axis=-1).astype(np.float64)
    w2 = bf16_round(w[2]).reshape(1536, 80).astype(np.float64)
    c = np.maximum(cols @ w2 + w[3].astype(np.float64), 0)
    act = bf16_round(c.astype(np.float32))
    h = np.maximum(act.reshape(n, 10560).astype(np.float64) @ bf16_round(w[4]).astype(np.float64) + w[5], 0)
    logits = h @ w[6].astype(np.float64) + w[7]
    return act, h, logits


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 37
    w = synth.vt_cnn2_weights(11, 1602)
    x = synth.iq_frames(n, seed=7)
    x[: max(1, n // 8)] *= 64
    m = vt_cnn2(11, mode="bf16")
    m.set_weights(w)
    z = m.predict(x, output="dense")
    lib = m._h._lib
    act_bits = np.empty((n * 132, 80), np.uint16)
    got = C.c_size_t()
    _lib.check(lib.mdc_debug_read(m._h.ptr, 0, act_bits.ctypes.data, act_bits.nbytes, C.byref(got)))
    hb = np.empty((n, 256), np.float32)
    _lib.check(lib.mdc_debug_read(m._h.ptr, 1, hb.ctypes.data, hb.nbytes, C.byref(got)))
    act = bf16_bits_to_f32(act_bits).reshape(n, 132, 80)
    ract, rh, rlog = emulate(x, w)
    d = np.abs(act - ract)
    tol = 2.0 ** -7 * np.abs(ract) + 1e-4 * np.abs(ract).max()   # 1 bf16 ulp + cancellation slack
    bad = d > tol
    print(f"n={n} conv2: max abs err {d.max():.4g}  mismatches(>1 bf16 ulp) {bad.sum()} / {bad.size}  ref max {np.abs(ract).max():.4g}")
    if bad.any():
        idx = np.argwhere(bad)
        print(" first bad (frame,pos,ch):", idx[:10].tolist())
        print(" bad per frame:", np.bincount(idx[:, 0], minlength=n)[:16].tolist())
        print(" bad per pos  :", np.bincount(idx[:, 1], minlength=132).tolist())
        f, p, c = idx[0]
        print(" got", act[f, p, c], "want", ract[f, p, c])
    dh = np.abs(hb - rh)
    print(f"dense1: max abs err {dh.max():.4g} (ref max {np.abs(rh).max():.4g})")
    dl = np.abs(z - rlog) / np.abs(rlog).max(axis=-1, keepdims=True)
    print(f"logits: max rel err vs emulation {dl.max():.4g}")
    ref = cf.vt_cnn2_forward(x, **cf.vt_cnn2_init(11, 1602), output="logits")
    dl64 = np.abs(z - ref) / np.abs(ref).max(axis=-1, keepdims=True)
    print(f"logits: max rel err vs fp64 oracle {dl64.max():.4g}")
    ok = (not bad.any()) and dh.max() < 2e-2 * max(1.0, np.abs(rh).max()) and dl.max() < 2e-3
    print("PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
