"""The digit extraction of the sliced FP64 GEMM (csrc/ozaki.cu digits8) on the CPU: the magic-number formulation the kernels
use since the end of round 2 (one rounding of x * 2^(7(t+1)) + 1.5 * 2^52 on the full-rate FP64 pipe, integer read from the low word,
exact remainder) yields bit-for-bit the digits of the rint() formulation it replaced, including ties and negative zeros, and the
FP32-count -> FP64 trick of the covariance epilogue (csrc/covtc.cu count_to_double) is exact for every count below 2^23."""
import numpy as np

MAGIC = 6755399441055744.0  # 1.5 * 2^52


def digits_rint(x):
    d = np.empty((8,) + x.shape, dtype=np.int64)
    x = x.copy()
    scale = 128.0
    for t in range(8):
        q = np.rint(x * scale)          # x * scale is exact (power of two)
        x = x - q / scale               # exact
        d[t] = q.astype(np.int64)
        scale *= 128.0
    return d, x


def digits_magic(x):
    d = np.empty((8,) + x.shape, dtype=np.int64)
    x = x.copy()
    scale, inv = 128.0, 1.0 / 128.0
    for t in range(8):
        s = x * scale + MAGIC           # the product is exact, so this is the ONE rounding of fma(x, scale, MAGIC)
        lo = s.view(np.uint64) & np.uint64(0xFFFFFFFF)
        d[t] = lo.astype(np.uint32).view(np.int32).astype(np.int64)   # __double2loint
        x = (MAGIC - s) * inv + x       # both steps exact: the fma of the kernel
        scale *= 128.0
        inv *= 1.0 / 128.0
    return d, x


def test_magic_digits_equal_rint_digits():
    rng = np.random.default_rng(0)
    x = np.concatenate([
        rng.uniform(-0.5, 0.5, 200000),
        rng.uniform(-0.5, 0.5, 50000) * 2.0 ** rng.integers(-60, 0, 50000),
        (rng.integers(-2 ** 20, 2 ** 20, 50000) + 0.5) / 2.0 ** rng.integers(21, 58, 50000),   # exact ties at some digit
        np.array([0.0, -0.0, 0.49999999999999994, -0.49999999999999994, 2.0 ** -57, -(2.0 ** -57), 3 * 2.0 ** -58, 63.5 / 128, -63.5 / 128]),
    ])
    x = x[np.abs(x) < 0.5]
    a, ra = digits_rint(x)
    b, rb = digits_magic(x)
    assert np.array_equal(a, b)
    assert np.array_equal(ra, rb)
    assert np.abs(a).max() <= 64 and np.abs(ra).max() <= 2.0 ** -57
    # the digits reproduce the value up to the dropped remainder
    val = sum(a[t].astype(np.float64) * 128.0 ** -(t + 1) for t in range(8))
    assert np.abs(val + ra - x).max() == 0.0 or np.abs(val + ra - x).max() <= 2.0 ** -60


def test_count_to_double_is_exact_below_2_23():
    c = np.concatenate([np.arange(0, 70000), np.array([2 ** 23 - 257, 2 ** 23 - 256, 2 ** 23 - 1]),
                        np.random.default_rng(1).integers(0, 2 ** 23, 100000)]).astype(np.int64)
    f = c.astype(np.float32)                                   # the FP32 accumulator: an exact integer
    u = (f + np.float32(8388608.0)).view(np.uint32) & np.uint32(0x007FFFFF)
    d = ((np.uint64(0x43300000) << np.uint64(32)) | u.astype(np.uint64)).view(np.float64) - 4503599627370496.0
    assert np.array_equal(d, c.astype(np.float64))
