"""Device random numbers behind Tensor.randn / Tensor.uniform on device="cuda" (reference: src/tensor/tensor.py:402-434
calls cupy.random there).  The library's stream is specified - Philox4x32-10, see include/microtorch_b200.h - so it is
restated here in NumPy and checked bit for bit (uniform) / to 1e-6 (normal: logf / sincospi differ in the last ulp)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def philox4x32_10(counter_lo, counter_hi, seed):
    """Blocks `counter` of the stream keyed by `seed`: uint32 array (n, 4)."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c = [counter_lo.astype(np.uint64), counter_hi.astype(np.uint64), np.zeros_like(counter_lo, np.uint64), np.zeros_like(counter_lo, np.uint64)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=1).astype(np.uint32)


def u01(x):
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


@pytest.fixture(scope="module")
def mt():
    from microtorch_b200 import _capi, cuda
    from microtorch_b200.tensor import Device, Tensor
    _capi.require_device(0)
    return type("NS", (), {"cuda": cuda, "Tensor": Tensor, "Device": Device})


@pytest.mark.parametrize("n", [1, 7, 4096, 1_000_003])
def test_uniform_stream_is_the_specified_philox_stream_bit_for_bit(mt, n):
    seed = 0x1234_5678_9ABC_DEF0
    mt.cuda.random.seed(seed)
    first = mt.cuda.random.uniform(-2.0, 3.0, (n,)).get()
    second = mt.cuda.random.uniform(0.0, 1.0, (5,)).get()           # continues the stream after ceil(n / 4) blocks
    blocks = (n + 3) // 4
    ctr = np.arange(blocks + 2, dtype=np.uint64)
    words = philox4x32_10(ctr & np.uint64(0xFFFFFFFF), ctr >> np.uint64(32), seed)
    want = (np.float32(-2.0) + np.float32(5.0) * u01(words[:blocks].reshape(-1)))[:n]
    np.testing.assert_array_equal(first, want.astype(np.float32))
    np.testing.assert_array_equal(second, u01(words[blocks:].reshape(-1))[:5])
    assert first.min() >= -2.0 and first.max() <= 3.0


def test_normal_stream_matches_box_muller_of_the_same_words(mt):
    seed, n = 77, 100_001
    mt.cuda.random.seed(seed)
    got = mt.cuda.random.randn(n).get()
    blocks = (n + 3) // 4
    ctr = np.arange(blocks, dtype=np.uint64)
    w = philox4x32_10(ctr & np.uint64(0xFFFFFFFF), ctr >> np.uint64(32), seed)
    u = u01(w).astype(np.float64)
    r0, r1 = np.sqrt(-2 * np.log(u[:, 0])), np.sqrt(-2 * np.log(u[:, 2]))
    z = np.stack([r0 * np.cos(2 * np.pi * u[:, 1]), r0 * np.sin(2 * np.pi * u[:, 1]),
                  r1 * np.cos(2 * np.pi * u[:, 3]), r1 * np.sin(2 * np.pi * u[:, 3])], axis=1).reshape(-1)[:n]
    np.testing.assert_allclose(got, z, rtol=0, atol=2e-6)


def test_tensor_factories_draw_on_the_device_and_have_the_right_moments(mt):
    np.random.seed(5)
    mt.cuda.random.seed(None)                       # unseeded: the seed comes from NumPy's global generator
    a = mt.Tensor.randn((2048, 2048), device=mt.Device.CUDA)
    np.random.seed(5)
    mt.cuda.random.seed(None)
    b = mt.Tensor.randn((2048, 2048), device=mt.Device.CUDA)
    ah = a.data.get()
    np.testing.assert_array_equal(ah, b.data.get())                 # np.random.seed fixes the device stream too
    assert a.device == mt.Device.CUDA and ah.dtype == np.float32 and ah.shape == (2048, 2048)
    assert abs(ah.mean()) < 3e-3 and abs(ah.std() - 1) < 3e-3
    assert abs((ah ** 3).mean()) < 1e-2 and abs((ah ** 4).mean() - 3) < 3e-2
    u = mt.Tensor.uniform(-1.0, 1.0, (1 << 22,), device=mt.Device.CUDA).data.get()
    assert u.min() >= -1 and u.max() <= 1 and abs(u.mean()) < 2e-3 and abs(u.var() - 1 / 3) < 2e-3
    hist = np.histogram(u, bins=64, range=(-1, 1))[0]
    assert np.abs(hist / hist.mean() - 1).max() < 0.03
    z = mt.Tensor.uniform(0.0, 1.0, (), device=mt.Device.CUDA)
    assert z.shape == () and 0 < float(z.item()) < 1
