"""GPU parity: FWHT kernels vs the dense-Hadamard oracle.  Bit-exact (int32, wrap mod 2^32)."""
import numpy as np
import pytest

from conftest import philox

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log2n", [5, 6, 7, 8, 9, 10, 11, 12, 13])
@pytest.mark.parametrize("ordering", ["natural", "sequency"])
def test_vs_matrix_oracle(log2n, ordering):
    from modulationdetectioncnn_b200.fwht import fwht
    from oracle import fwht as of
    n = 1 << log2n
    x = philox(log2n).integers(-(1 << 17), 1 << 17, (37, n)).astype(np.int32)
    assert np.array_equal(fwht(x, ordering), of.fwht_matrix(x, ordering))


def test_wraparound_out_of_contract():
    from modulationdetectioncnn_b200.fwht import fwht
    from oracle import fwht as of
    x = philox(1).integers(-(1 << 31), 1 << 31, (9, 1024)).astype(np.int32)
    assert np.array_equal(fwht(x), of.fwht_matrix(x))


def test_shapes_empty_inplace_and_device_path():
    import torch
    from modulationdetectioncnn_b200.fwht import fwht
    from oracle import fwht as of
    assert fwht(np.zeros((0, 1024), np.int32)).shape == (0, 1024)
    x = philox(2).integers(-1000, 1000, (4, 3, 2, 1024)).astype(np.int32)   # [channels, T, I/Q, 1024]
    want = of.fwht_matrix(x.reshape(-1, 1024)).reshape(x.shape)
    assert np.array_equal(fwht(x), want)
    xt = torch.from_numpy(x).cuda()
    assert np.array_equal(fwht(xt).cpu().numpy(), want)
    with pytest.raises(ValueError):
        fwht(np.zeros((2, 1000), np.int32))
    for n in (1, 7, 1025, 20000):             # ragged counts around grid / host-chunk boundaries
        y = philox(n).integers(-(1 << 17), 1 << 17, (n, 1024)).astype(np.int32)
        assert np.array_equal(fwht(y), of.fwht_butterfly(y))


def test_full_size_involution():
    """C4 size (>= 1 GiB): H(H(x)) = N x, a size-independent property; plus a sampled oracle check."""
    import torch
    from modulationdetectioncnn_b200.fwht import fwht
    from oracle import fwht as of
    s = 1 << 18                                # 2^18 spectra x 1024 x 4 B = 1 GiB
    gen = torch.Generator(device="cuda").manual_seed(2015)
    x = torch.randn((s, 1024), generator=gen, device="cuda").mul_(32).trunc_().to(torch.int32)
    y = fwht(x)
    assert torch.equal(fwht(y), x * 1024)
    assert torch.equal(y[:, 0].to(torch.int64), x.sum(-1))                  # DC bin = sum
    idx = torch.arange(0, s, s // 512, device="cuda")
    assert np.array_equal(y[idx].cpu().numpy(), of.fwht_matrix(x[idx].cpu().numpy()))


def test_in_place_allowed_partial_overlap_rejected():
    """in_dev == out_dev transforms in place (include/mdc.h); any other overlap would let one spectrum's stores land in
    another's unread input and is refused before anything is launched."""
    import torch
    from modulationdetectioncnn_b200 import _lib
    from oracle import fwht as of
    lib = _lib.load()
    x = philox(3).integers(-(1 << 17), 1 << 17, (64, 1024)).astype(np.int32)
    buf = torch.zeros((65, 1024), dtype=torch.int32, device="cuda")
    buf[:64] = torch.from_numpy(x).cuda()
    st = torch.cuda.current_stream().cuda_stream
    assert lib.mdc_fwht_i32(buf.data_ptr(), buf[1:].data_ptr(), 64, 10, 0, st) == -1          # shifted by one spectrum
    assert b"overlap" in lib.mdc_last_error()
    torch.cuda.synchronize()
    assert np.array_equal(buf[:64].cpu().numpy(), x)                                           # nothing was launched
    _lib.check(lib.mdc_fwht_i32(buf.data_ptr(), buf.data_ptr(), 64, 10, 0, st))                # in place
    torch.cuda.synchronize()
    assert np.array_equal(buf[:64].cpu().numpy(), of.fwht_matrix(x))
