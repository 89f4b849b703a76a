"""Raw RTL-SDR ingest: interleaved unsigned 8-bit I/Q -> the input formats of the hot path.

The reference only states where its samples come from (/root/reference/README.md:5: an RTL-SDR
through the ARM/HPS side of the DE1-SoC); this is the step before ``predict`` / the FWHT, done on
the GPU in one pass over the byte stream (``mdc_sdr_ingest_u8``, include/mdc.h).

    value = (u - 127.5) / 128          Q6.12 integer = (2 u - 255) * 16   (exact)
"""
from __future__ import annotations

from typing import Dict, Sequence

from . import _lib

__all__ = ["ingest_u8"]


def ingest_u8(iq, want: Sequence[str] = ("f32",)) -> Dict[str, object]:
    """iq: CUDA uint8 tensor [2*n] (I0 Q0 I1 Q1 ...), n a multiple of 128 (1024 for "fwht").

    Returns CUDA tensors for the names in ``want``:
      "f32"  float32 [n/128, 2, 128]   -> CNN2Model.predict
      "q612" int32   [n/128, 256]      -> FixedPointCNN2.predict
      "fwht" int32   [n/1024, 2, 1024] -> fwht.fwht (one I and one Q spectrum per block)
    """
    import torch
    if not (isinstance(iq, torch.Tensor) and iq.is_cuda and iq.dtype == torch.uint8):
        raise ValueError("iq must be a CUDA uint8 tensor (there is no CPU path)")
    bad = set(want) - {"f32", "q612", "fwht"}
    if bad or not want:
        raise ValueError(f"want must name some of f32/q612/fwht, got {sorted(want)}")
    iq = iq.contiguous().reshape(-1)
    if iq.numel() % 2:
        raise ValueError("odd number of bytes: I/Q pairs expected")
    n = iq.numel() // 2
    if n % 128 or ("fwht" in want and n % 1024):
        raise ValueError(f"{n} samples: frames need a multiple of 128, FWHT blocks a multiple of 1024")
    out: Dict[str, object] = {}
    with torch.cuda.device(iq.device):
        if "f32" in want:
            out["f32"] = torch.empty((n // 128, 2, 128), dtype=torch.float32, device=iq.device)
        if "q612" in want:
            out["q612"] = torch.empty((n // 128, 256), dtype=torch.int32, device=iq.device)
        if "fwht" in want:
            out["fwht"] = torch.empty((n // 1024, 2, 1024), dtype=torch.int32, device=iq.device)
        ptr = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
        _lib.check(_lib.load().mdc_sdr_ingest_u8(iq.data_ptr(), n, ptr("f32"), ptr("q612"), ptr("fwht"),
                                                 torch.cuda.current_stream(iq.device).cuda_stream))
    return out
