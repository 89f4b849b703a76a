"""TEST INFRASTRUCTURE - CPU restatement of the reference's fixed-point datapath.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.

Restates /root/reference/cnn_test_latest1.sv (no Verilog simulator exists in
this image, so the file cannot be executed) twice, independently:

* ``simulate_rtl``     - a register-transfer, clock-by-clock model of the modules
  ``test_input`` (:71-117), ``layers_top`` (:144-209), ``dense_layer``
  (:240-399), ``conv_top``/``conv_layer``/``conv_compute`` (:401-637) and the
  synchronous ``case`` ROMs (:685-3132), with non-blocking-assignment
  semantics (all next-state values computed from the current state, then
  committed).  Pure Python, one frame at a time.
* ``forward``          - the closed form those modules reduce to, vectorised
  over frames with numpy int64.

``tests/test_oracle_int.py`` fuzzes one against the other, including full-range
18-bit inputs that exercise the 36-bit product wrap, the forced-sign bit slice
and the 18-bit bias-add wrap.

Pinning: **integer outputs are parity-unpinned** - the reference records no SV
output anywhere (the testbench has no $display / expected values).  What *is*
pinned by reference artefacts: the ROM contents (SV ROMs == 12.15.latestWeights.txt)
and the embedded 2-sample testbench vector (:130,133), whose result under this
restatement is recorded in tests/golden/.

Arithmetic:

    slice(a,b,c,d) = s18({m[35], m[28:12]}),  m = (a*b + c*d) mod 2**36   (:653-655,:672-674)
    conv(a,b,c,d,e)= relu18(s18(slice + e))                                (:656-657)
    y[f][p][r]     = conv(xp[r][p], CT[3f], xp[r][p+1], CT[3f+1], CT[3f+2]),  p=0..128
    acc[c]         = sext(DB[c]) + sum_{f,s<128} slice(y[f][s][I], WI_c[A], y[f][s][Q], WQ_c[A])
                     with A = 128*f + max(s-1, 0)   (1-cycle ROM latency, :336,:351-378)
    out[c]         = acc[c] if acc[c] >= 0 else 0   (32-bit wrap, :175-177)
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

__all__ = ["mult_slice", "forward", "forward_pre", "forward_deskewed_pre", "simulate_rtl", "dense_rom_address",
           "deskewed_rom_address"]

_M36 = (1 << 36) - 1


def _s(v: np.ndarray, width: int) -> np.ndarray:
    m = np.int64(1) << width
    v = v & (m - 1)
    return np.where(v >= (m >> 1), v - m, v)


def mult_slice(a, b, c, d) -> np.ndarray:
    """``{mult_out[35], mult_out[28:12]}`` of the 36-bit ``a*b + c*d`` (signed_mult, sv:664-675)."""
    a, b, c, d = (np.asarray(v, dtype=np.int64) for v in (a, b, c, d))
    u = (a * b + c * d) & _M36                       # |a*b + c*d| < 2**35 in int64: no overflow
    bits = (((u >> 35) & 1) << 17) | ((u >> 12) & 0x1FFFF)
    return _s(bits, 18)


def dense_rom_address(F: int) -> np.ndarray:
    """ROM address used for (filter f, sample s<128): ``128*f + max(s-1,0)`` -> int64 [F,128]."""
    s = np.arange(128)
    return 128 * np.arange(F)[:, None] + np.maximum(s - 1, 0)[None, :]


def _accumulate(x: np.ndarray, conv_tab, dense_bias, dense_tabs, address: np.ndarray) -> np.ndarray:
    """The datapath with the dense ROM address map as a parameter: ``address`` int [F,P] gives, for filter f and conv
    position p < P, the ROM entry that multiplies conv output (f, p).  Everything else - input map, zero padding,
    slice / bias wrap / ReLU of the conv stage, slice and 32-bit accumulate of the dense stage - is common."""
    x = np.asarray(x, dtype=np.int64).reshape(-1, 256)
    conv_tab = np.asarray(conv_tab, dtype=np.int64)
    dense_bias = np.asarray(dense_bias, dtype=np.int64)
    dense_tabs = np.asarray(dense_tabs, dtype=np.int64)
    N = x.shape[0]
    F = conv_tab.shape[0] // 3
    C = dense_tabs.shape[0] // 2
    P = address.shape[1]
    xp = np.zeros((N, 2, 130), dtype=np.int64)
    xp[:, 0, 1:129] = x[:, 0:128]        # I  (test_input, sv:88-89,102)
    xp[:, 1, 1:129] = x[:, 128:256]      # Q
    acc = np.broadcast_to(dense_bias, (N, C)).copy()
    for f in range(F):
        w0, w1, bias = conv_tab[3 * f], conv_tab[3 * f + 1], conv_tab[3 * f + 2]
        y = _s(mult_slice(xp[:, :, 0:P], w0, xp[:, :, 1:P + 1], w1) + bias, 18)   # conv positions 0..P-1
        y = np.where(y < 0, 0, y)        # [N,2,P]
        for c in range(C):
            wi = dense_tabs[2 * c][address[f]]
            wq = dense_tabs[2 * c + 1][address[f]]
            acc[:, c] += mult_slice(y[:, 0, :], wi, y[:, 1, :], wq).sum(axis=1)
    return _s(acc, 32).astype(np.int32)


def forward_pre(x: np.ndarray, conv_tab, dense_bias, dense_tabs) -> np.ndarray:
    """32-bit accumulators before the final ReLU, as the hardware computes them: conv positions 0..127 only, ROM
    address ``128*f + max(s-1,0)``.  x int [N,256] -> int32 [N,C]."""
    F = np.asarray(conv_tab).shape[0] // 3
    return _accumulate(x, conv_tab, dense_bias, dense_tabs, dense_rom_address(F))


def deskewed_rom_address(F: int) -> np.ndarray:
    """The address map the weight tables were LAID OUT for (``table[f*129 + p]`` =
    ``float2fix(DenseKernel[r*387 + p*3 + f, c])``): entry ``129*f + p`` for every conv position p = 0..128."""
    return 129 * np.arange(F)[:, None] + np.arange(129)[None, :]


def forward_deskewed_pre(x: np.ndarray, conv_tab, dense_bias, dense_tabs) -> np.ndarray:
    """The quantisation of the Keras net the authors intended: the SAME arithmetic primitives as ``forward_pre``
    (``_accumulate``), with the dense ROM read at ``129*f + p`` over all 129 positions instead of the skewed
    ``128*f + max(s-1,0)`` over 128.  Not what the hardware computes - it exists to pin slice / bias wrap / ReLU /
    accumulate against the Keras output the reference records (12.16.testDataYunyun.txt:264, CNN.ipynb cell 18):
    ``tests/test_oracle_int.py::test_deskewed_datapath_reproduces_recorded_keras_output``."""
    F = np.asarray(conv_tab).shape[0] // 3
    return _accumulate(x, conv_tab, dense_bias, dense_tabs, deskewed_rom_address(F))


def forward(x: np.ndarray, conv_tab, dense_bias, dense_tabs) -> np.ndarray:
    """``out_data`` of ``layers_top``: ReLU of the accumulators.  int32 [N,C] (Q.12)."""
    pre = forward_pre(x, conv_tab, dense_bias, dense_tabs)
    return np.where(pre < 0, 0, pre).astype(np.int32)


# --------------------------------------------------------------------------- RTL model
def _s_int(v: int, width: int) -> int:
    m = 1 << width
    v &= m - 1
    return v - m if v >= (m >> 1) else v


def _slice_int(a: int, b: int, c: int, d: int) -> int:
    u = (a * b + c * d) & _M36
    return _s_int((((u >> 35) & 1) << 17) | ((u >> 12) & 0x1FFFF), 18)


def _rom(tab, addr: int, hold: int) -> int:
    """Synchronous ``case`` ROM without default: unmatched address holds the old value."""
    return int(tab[addr]) if 0 <= addr < len(tab) else hold


def simulate_rtl(vec, conv_tab, dense_bias, dense_tabs, reset_cycles: int = 2,
                 max_cycles: int = 2000) -> Tuple[np.ndarray, np.ndarray, Dict[str, int]]:
    """Clock-by-clock model of ``testbench_iteration1`` for one 256-entry vector.

    Returns (out_data int32[C], pre_out_data int32[C], info) where
    ``info["cycles"]`` is the number of clk_50 edges after reset until ``nn_done``.
    Works for F filters / C classes (the SV file instantiates F=3, C=3).
    """
    vec = [int(v) for v in np.asarray(vec).reshape(-1)]
    conv_tab = [int(v) for v in conv_tab]
    dense_bias = [int(v) for v in dense_bias]
    dense_tabs = [[int(v) for v in t] for t in dense_tabs]
    F, C = len(conv_tab) // 3, len(dense_bias)

    def table(addr):  # test_table: combinational, default 0
        return vec[addr] if 0 <= addr < 256 else 0

    # ---- state (registers); X treated as 0
    ti = dict(i_addr=0, q_addr=0, sidx=0, done=0, data=[(0, 0)] * 128)
    top = dict(done=0, started=0, out=[0] * C)
    cl = [dict(ad=[(0, 0)] * 130, started=0, done=0, out_cov=[(0, 0)] * 129, sidx=0,
               waddr=[0, 0], baddr=0, w=[0, 0], b=0) for _ in range(F)]
    dl = dict(sidx=0, kidx=0, war=0, acc=[0] * C, done=0, started=0, w=[[0, 0] for _ in range(C)])

    cycles_after_reset = 0
    for cyc in range(max_cycles):
        reset = 1 if cyc < reset_cycles else 0
        # ---------- combinational views of the *current* state
        input_done = ti["done"]
        conv_done = cl[0]["done"]
        i_data, q_data = table(ti["i_addr"]), table(ti["q_addr"])
        # conv_output[k][p] = out_cov[k][p]  (double reversal cancels, sv:456-468,507)
        # ---------- next state
        # test_input (sv:86-108)
        nti = dict(ti, data=list(ti["data"]))
        if reset:
            nti.update(i_addr=0, q_addr=128, sidx=0, done=0)
        elif ti["sidx"] == 128:
            nti.update(done=1, sidx=129)
        elif ti["sidx"] == 129:
            nti.update(done=0)
        else:
            nti["data"][ti["sidx"]] = (i_data, q_data)
            nti.update(sidx=(ti["sidx"] + 1) & 0xFF, i_addr=ti["i_addr"] + 1, q_addr=ti["q_addr"] + 1)

        # conv layers (sv:476-509) + conv_compute (sv:548-633) + rom_cov (sv:692-705)
        ncl = []
        for k, c in enumerate(cl):
            n = dict(c, ad=list(c["ad"]), out_cov=list(c["out_cov"]), waddr=list(c["waddr"]), w=list(c["w"]))
            # address registers
            if reset:
                n["waddr"], n["baddr"] = [0, 1], 2
            elif 1 <= k <= 9:
                n["waddr"], n["baddr"] = [3 * k, 3 * k + 1], 3 * k + 2
            # ROM data registers sample the *current* address
            n["w"] = [_rom(conv_tab, c["waddr"][0], c["w"][0]), _rom(conv_tab, c["waddr"][1], c["w"][1])]
            n["b"] = _rom(conv_tab, c["baddr"], c["b"])
            # datapath (combinational from current regs)
            s = c["sidx"]
            if input_done or reset or not c["started"]:
                n["sidx"], n["done"] = 0, 0
                for i in range(1, 129):
                    n["ad"][i] = ti["data"][i - 1]
                n["ad"][0] = (0, 0)
                n["ad"][129] = (0, 0)
                n["started"] = 1 if input_done else 0
            elif s == 129:
                n["done"] = 1
            else:
                d1, d2 = c["ad"][s], c["ad"][s + 1]
                res = []
                for r in (0, 1):
                    o = _s_int(_slice_int(d1[r], c["w"][0], d2[r], c["w"][1]) + c["b"], 18)
                    res.append(o if o >= 0 else 0)
                n["out_cov"][128 - s] = (res[0], res[1])
                n["sidx"] = s + 1
            ncl.append(n)

        # dense layer (sv:292-348) + six ROMs
        ndl = dict(dl, acc=list(dl["acc"]))
        ndl["w"] = [[_rom(dense_tabs[2 * c], dl["war"], dl["w"][c][0]),
                     _rom(dense_tabs[2 * c + 1], dl["war"], dl["w"][c][1])] for c in range(C)]
        if reset or (conv_done and not dl["started"]):
            ndl.update(sidx=0, kidx=0, war=0, done=0, acc=[_s_int(b, 32) for b in dense_bias],
                       started=1 if conv_done else 0)
        elif dl["sidx"] == 128:
            if dl["kidx"] == F - 1:
                ndl["done"] = 1
            else:
                ndl.update(sidx=0, kidx=dl["kidx"] + 1)
        elif dl["started"]:
            # cov_out[k] packed index p  ==  out_cov[128-p]
            iq = cl[dl["kidx"]]["out_cov"][128 - dl["sidx"]]
            for c in range(C):
                cur = _slice_int(iq[0], dl["w"][c][0], iq[1], dl["w"][c][1])
                ndl["acc"][c] = _s_int(dl["acc"][c] + cur, 32)
            ndl.update(sidx=dl["sidx"] + 1, war=(dl["war"] + 1) & 0x3FFFF)

        # layers_top (sv:164-187)
        ntop = dict(top, out=list(top["out"]))
        if reset:
            ntop.update(done=0, started=0)
        elif input_done and not top["started"]:
            ntop.update(done=0, started=1)
        elif dl["done"]:
            ntop["done"] = 1
            ntop["out"] = [a if a >= 0 else 0 for a in dl["acc"]]

        ti, cl, dl, top = nti, ncl, ndl, ntop
        if not reset:
            cycles_after_reset += 1
        if top["done"]:
            return (np.array(top["out"], dtype=np.int32), np.array(dl["acc"], dtype=np.int32),
                    {"cycles": cycles_after_reset})
    raise RuntimeError("nn_done never asserted")
