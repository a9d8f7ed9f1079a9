"""Host mirror (numpy, bit-exact) of the device synthetic-row generator (csrc/aux_kernels.cuh synth_value +
the normalize_avx2 arithmetic), so parity tests can rebuild any row the GPU generated."""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def synth_raw(seed: int, rows: np.ndarray, d: int) -> np.ndarray:
    rows = np.asarray(rows, np.uint64)[:, None]
    cols = np.arange(d, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        z = np.uint64(seed) ^ (rows * np.uint64(0x9E3779B97F4A7C15)) ^ (cols * np.uint64(0xC2B2AE3D27D4EB4F))
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    s = ((z & np.uint64(0xFFFF)).astype(np.int64) + ((z >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
         + ((z >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64) + ((z >> np.uint64(48)) & np.uint64(0xFFFF)).astype(np.int64)
         - 131070)
    return (s.astype(np.float32) * np.float32(1.0 / 65536.0)).astype(np.float32)


def _sqnorm_simd_order(x: np.ndarray) -> np.ndarray:
    """dot_product_avx2(v, v) order (simd_ops.rs:149-183) for each row of x, in float32."""
    n, d = x.shape
    chunks = d // 8
    acc = np.zeros((n, 8), np.float32)
    x64 = x.astype(np.float64)
    for i in range(chunks):
        v = x64[:, 8 * i:8 * i + 8]
        acc = (v * v + acc.astype(np.float64)).astype(np.float32)     # fmaf: exact product+add, one rounding
    a = acc
    h = ((a[:, 0] + a[:, 4]) + (a[:, 1] + a[:, 5])) + ((a[:, 2] + a[:, 6]) + (a[:, 3] + a[:, 7]))
    t = np.zeros(n, np.float32)
    for i in range(chunks * 8, d):
        t = t + x[:, i] * x[:, i]
    return (h + t).astype(np.float32)


def synth_rows(seed: int, rows, d: int, unit_norm: bool = True, f16: bool = False) -> np.ndarray:
    x = synth_raw(seed, rows, d)
    if unit_norm:
        nsq = _sqnorm_simd_order(x)
        inv = (np.float32(1.0) / np.sqrt(nsq, dtype=np.float32)).astype(np.float32)
        inv = np.where(nsq == 0, np.float32(1.0), inv).astype(np.float32)
        x = (x * inv[:, None]).astype(np.float32)
    if f16:
        x = x.astype(np.float16).astype(np.float32)
    return x
