"""Walsh-Hadamard spectrogram stage (README.md:5 of the reference; no code there)."""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["fwht"]


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def fwht(x, ordering: str = "natural", device: int = 0):
    """Unnormalised WHT along the last axis.  int32, length 2**k (32..8192), wrap mod 2**32.

    numpy in -> numpy out (host path, chunked H2D/D2H inside the library);
    CUDA torch tensor in -> torch tensor out on the current stream.
    """
    order = {"natural": _lib.FWHT_NATURAL, "sequency": _lib.FWHT_SEQUENCY}.get(ordering)
    if order is None:
        raise ValueError("ordering must be 'natural' or 'sequency'")
    n = int(x.shape[-1])
    log2n = n.bit_length() - 1
    if n != 1 << log2n:
        raise ValueError("transform length must be a power of two")
    lib = _lib.load()
    if _is_torch(x):
        import torch
        if not x.is_cuda:
            raise ValueError("torch inputs must be CUDA tensors (pass numpy for the host path)")
        xt = x.to(torch.int32).contiguous()
        out = torch.empty_like(xt)
        with torch.cuda.device(xt.device):
            _lib.check(lib.mdc_fwht_i32(xt.data_ptr(), out.data_ptr(), xt.numel() // n, log2n, order,
                                        torch.cuda.current_stream(xt.device).cuda_stream))
        return out
    xa = np.ascontiguousarray(x, dtype=np.int32)
    out = np.empty_like(xa)
    _lib.check(lib.mdc_fwht_i32_host(xa.ctypes.data, out.ctypes.data, xa.size // n, log2n, order, device))
    return out
