"""CPU checks of the raw-ingest oracle (definition in include/mdc.h; the reference has no ingest code)."""
import numpy as np
import pytest

from modulationdetectioncnn_b200 import fixedpoint as fx
from oracle import sdr as osdr


def test_q612_is_the_reference_quantiser_of_the_value():
    """(2u - 255) * 16 == float2fix((u - 127.5) / 128, 18, 12) for every byte (CNN.ipynb cell 23)."""
    u = np.arange(256)
    assert np.array_equal(fx.quantize((u - 127.5) / 128.0), (2 * u - 255) * 16)


def test_layouts():
    n = 2048
    raw = np.arange(2 * n, dtype=np.int64).astype(np.uint8)
    f32, q612, fw = osdr.ingest_u8(raw)
    assert f32.shape == (16, 2, 128) and q612.shape == (16, 256) and fw.shape == (2, 2, 1024)
    i, q = raw[0::2].astype(np.int64), raw[1::2].astype(np.int64)
    assert np.array_equal(q612[3, :128], (2 * i[384:512] - 255) * 16)        # frame 3, I row
    assert np.array_equal(q612[3, 128:], (2 * q[384:512] - 255) * 16)        # frame 3, Q row
    assert np.array_equal(fw[1, 1], (2 * q[1024:2048] - 255) * 16)           # block 1, Q block
    assert np.array_equal(f32.reshape(16, 256) * 4096, q612)                 # floats are exact
    assert osdr.ingest_u8(raw[: 2 * 128])[2] is None
    with pytest.raises(ValueError):
        osdr.ingest_u8(raw[:100])
