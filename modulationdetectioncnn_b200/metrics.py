"""Caller-side bookkeeping of the reference on the GPU: confusion matrices and per-SNR accuracy.

Replaces the O(N) interpreter loops of /root/reference/cnn.py:200-218 (confusion matrix of the whole test
set) and :227-255 (for every SNR: select that SNR's frames, predict, count, ``acc[snr] = trace / sum``) with
one pass of ``mdc_confusion_grouped_i32`` over class ids that are already on the device
(``CNN2Model.predict(x, output="argmax")`` with a CUDA tensor).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib

__all__ = ["confusion_matrix", "normalize_rows", "accuracy_by_group", "accuracy_by_snr"]


def confusion_matrix(y_true, y_pred, classes: int, groups=None, n_groups: int = 1):
    """Counts ``conf[g, true, pred]`` as a CUDA int64 tensor [n_groups, classes, classes] ([classes, classes]
    when ``groups`` is None).  ``y_true`` / ``y_pred`` / ``groups``: CUDA integer tensors [N] (one-hot
    ``y_true`` [N, classes] is accepted, like ``Y_test`` in cnn.py); labels outside their range are skipped."""
    import torch
    if not (isinstance(y_pred, torch.Tensor) and y_pred.is_cuda):
        raise ValueError("y_pred must be a CUDA tensor (there is no CPU path)")
    dev = y_pred.device
    y_true = torch.as_tensor(y_true, device=dev)
    if y_true.ndim == 2:
        y_true = y_true.argmax(-1)                      # list(Y_test[i, :]).index(1)
    t = y_true.to(torch.int32).contiguous()
    p = y_pred.to(torch.int32).contiguous()
    if t.shape != p.shape or t.ndim != 1:
        raise ValueError(f"y_true {tuple(t.shape)} and y_pred {tuple(p.shape)} must be equal-length vectors")
    g = None
    if groups is not None:
        g = torch.as_tensor(groups, device=dev).to(torch.int32).contiguous()
        if g.shape != t.shape:
            raise ValueError("groups must have one entry per frame")
    conf = torch.zeros((n_groups, classes, classes), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().mdc_confusion_grouped_i32(t.data_ptr(), p.data_ptr(), g.data_ptr() if g is not None else None,
                                                         t.numel(), classes, n_groups, conf.data_ptr(),
                                                         torch.cuda.current_stream(dev).cuda_stream))
    return conf if groups is not None else conf[0]


def normalize_rows(conf) -> np.ndarray:
    """``confnorm[i, :] = conf[i, :] / sum(conf[i, :])`` (cnn.py:214-215); empty rows stay zero."""
    c = np.asarray(conf.cpu() if hasattr(conf, "cpu") else conf, dtype=np.float64)
    s = c.sum(-1, keepdims=True)
    return np.divide(c, s, out=np.zeros_like(c), where=s > 0)


def accuracy_by_group(conf) -> np.ndarray:
    """``cor / (cor + ncor)`` per group from [G, C, C] counts (cnn.py:252-254); NaN for empty groups."""
    c = np.asarray(conf.cpu() if hasattr(conf, "cpu") else conf, dtype=np.float64)
    tot = c.sum((-1, -2))
    cor = np.trace(c, axis1=-2, axis2=-1)
    return np.divide(cor, tot, out=np.full_like(tot, np.nan), where=tot > 0)


def accuracy_by_snr(y_true, y_pred, snr, classes: int, snrs: Optional[Sequence[int]] = None) -> Dict[int, float]:
    """The ``acc`` dictionary of cnn.py:227-255 in one pass: ``snr`` is the per-frame SNR label
    (``lbl[x][1]`` there), ``snrs`` the values to report (default: the distinct values, sorted)."""
    import torch
    snr_t = torch.as_tensor(snr, device=y_pred.device)
    values = sorted(int(v) for v in (snrs if snrs is not None else torch.unique(snr_t).tolist()))
    lut = {v: i for i, v in enumerate(values)}
    if values:
        lo, hi = values[0], values[-1]
        table = torch.full((hi - lo + 1,), -1, dtype=torch.int32, device=y_pred.device)
        table[torch.tensor([v - lo for v in values], device=y_pred.device)] = torch.arange(len(values), dtype=torch.int32,
                                                                                          device=y_pred.device)
        idx = (snr_t.to(torch.int64) - lo).clamp_(0, hi - lo)
        grp = torch.where((snr_t >= lo) & (snr_t <= hi), table[idx], torch.full_like(table[idx], -1))
    else:
        grp = torch.full(snr_t.shape, -1, dtype=torch.int32, device=y_pred.device)
    conf = confusion_matrix(y_true, y_pred, classes, groups=grp, n_groups=max(len(values), 1))
    acc = accuracy_by_group(conf)
    return {v: float(acc[lut[v]]) for v in values}
