"""Host staging of index tensors (crfconv_b200/host_io.py, crfconv_pack_index_host / crfconv_unpack_index): bit-exact by
construction, so the checks are equality against numpy casts; an index that does not fit must raise, never truncate.
The host half runs without a GPU; the round trip through the device is a `-m gpu` test."""
import numpy as np
import pytest
import torch

from crfconv_b200 import _lib


@pytest.fixture(scope="module")
def L():
    from crfconv_b200 import build
    build.build()
    return _lib.lib()


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 65535, 65536, 3 * 65536 + 5, 6 * 40960 * 16])
@pytest.mark.parametrize("bits", [16, 32])
@pytest.mark.parametrize("threads", [1, 5])
def test_pack_matches_a_numpy_cast(L, n, bits, threads):
    rng = np.random.default_rng(n + bits)
    hi = (1 << 16) if bits == 16 else (1 << 31)
    idx = rng.integers(0, hi, size=n, dtype=np.int64)
    if n:
        idx[-1] = hi - 1                                   # the largest representable value in the last (ragged) slot
        idx[0] = 0
    out = np.full(n + 3, 0x5A5A, dtype=np.uint16 if bits == 16 else np.uint32)     # 3 guard elements behind the output
    rc = L.crfconv_pack_index_host(idx.ctypes.data, n, bits, out.ctypes.data, threads)
    assert rc == 0
    assert np.array_equal(out[:n].astype(np.int64), idx)
    assert np.all(out[n:] == 0x5A5A)


@pytest.mark.parametrize("bits,bad", [(16, 65536), (16, -1), (32, 1 << 32), (32, -5), (16, 1 << 40)])
def test_pack_refuses_indices_that_do_not_fit(L, bits, bad):
    idx = np.arange(200000, dtype=np.int64) % 4096
    idx[123457] = bad
    out = np.zeros(idx.size, dtype=np.uint32)
    assert L.crfconv_pack_index_host(idx.ctypes.data, idx.size, bits, out.ctypes.data, 4) == -1
    assert L.crfconv_pack_index_host(idx.ctypes.data, idx.size, 24, out.ctypes.data, 4) == -1        # unsupported width
    with pytest.raises(RuntimeError):
        _lib.check(-1, "pack_index_host")


def test_index_bits():
    from crfconv_b200.host_io import index_bits
    assert index_bits(40960) == 16 and index_bits(65536) == 16 and index_bits(65537) == 32 and index_bits(1 << 31) == 32


@pytest.mark.gpu
@pytest.mark.parametrize("pack", [True, False])
def test_stager_round_trip_is_bit_exact(pack):
    from crfconv_b200.host_io import HostStager
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(7)
    B, N, Nc, K = 2, 5000, 1250, 16
    host = {"unary": torch.randn(B, Nc, 128, generator=g).pin_memory(),
            "pairwise": torch.randn(B, N, 64, generator=g).pin_memory(),
            "up_idx": torch.randint(0, Nc, (B, N, 1), generator=g).pin_memory(),
            "neighbor_idx": torch.randint(0, N, (B, N, K), generator=g).pin_memory(),
            "wide_idx": torch.randint(0, 1 << 20, (B, 777), generator=g).pin_memory()}
    st = HostStager(host, dev, index_limits={"up_idx": Nc, "neighbor_idx": N, "wide_idx": 1 << 20}, pack=pack)
    assert (st.bits == {"up_idx": 16, "neighbor_idx": 16, "wide_idx": 32}) if pack else (st.bits == {})
    side = torch.cuda.Stream()
    for rep in range(3):                                   # the same device tensors are refilled on every call
        if rep:
            host["neighbor_idx"].copy_(torch.randint(0, N, (B, N, K), generator=g))
            host["pairwise"].add_(1.0)
        out = st.upload(host, side)
        side.synchronize()
        for k, v in host.items():
            assert out[k].dtype == v.dtype and torch.equal(out[k].cpu(), v), k
    host["neighbor_idx"].copy_(torch.randint(0, N, (B, N, K), generator=g))     # packing ahead of the upload (helper-thread form)
    st.prepare(host)
    out = st.upload(host, side, prepared=True)
    side.synchronize()
    assert torch.equal(out["neighbor_idx"].cpu(), host["neighbor_idx"]) and torch.equal(out["up_idx"].cpu(), host["up_idx"])
    host["neighbor_idx"].copy_(torch.randint(0, N, (B, N, K), generator=g))     # widening deferred to the consumer's stream
    out = st.upload(host, side, defer_unpack=True)
    main = torch.cuda.current_stream()
    main.wait_stream(side)
    st.unpack(main)
    main.synchronize()
    assert torch.equal(out["neighbor_idx"].cpu(), host["neighbor_idx"]) and torch.equal(out["pairwise"].cpu(), host["pairwise"])
    ref = sum(v.numel() * v.element_size() for v in host.values())
    assert st.h2d_bytes(host) == (ref - 6 * (B * N + B * N * K) - 4 * B * 777 if pack else ref)
    if pack:
        host["neighbor_idx"][1, 17, 3] = 70000             # does not fit 16 bits: loud failure
        with pytest.raises(RuntimeError):
            st.upload(host, side)


@pytest.mark.gpu
def test_debug_index_check_raises_like_torch_gather():
    """ops.CHECK_INDICES: an out-of-range neighbour / up-sampling index raises IndexError (the reference's torch.gather raises,
    models/continuous_crf_conv_big.py:40-44) instead of reading out of bounds; off by default."""
    from crfconv_b200 import ops
    from crfconv_b200.continuous_crf_conv_big import ContinuousGaussianCRFConv
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    B, N, Nc, K = 1, 512, 128, 16
    layer = ContinuousGaussianCRFConv(128, 64, 64, steps=1).to(dev).train()
    unary, pairwise = torch.randn(B, Nc, 128, generator=g).to(dev), torch.randn(B, N, 64, generator=g).to(dev)
    up = torch.randint(0, Nc, (B, N, 1), generator=g).to(dev)
    nbr = torch.randint(0, N, (B, N, K), generator=g).to(dev)
    assert ops.CHECK_INDICES is False
    ops.CHECK_INDICES = True
    try:
        layer(unary, pairwise, up, nbr)                                  # valid indices pass
        bad = nbr.clone(); bad[0, 7, 5] = N
        with pytest.raises(IndexError):
            layer(unary, pairwise, up, bad)
        bad_up = up.clone(); bad_up[0, 3, 0] = -1
        with pytest.raises(IndexError):
            layer(unary, pairwise, bad_up, nbr)
    finally:
        ops.CHECK_INDICES = False
