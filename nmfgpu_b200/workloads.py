"""Deterministic synthetic inputs shared by tests, bench.py and the reference driver.

One counter-based generator (a splitmix64 finaliser of seed and linear column-major index) is defined
twice with identical arithmetic: here in numpy and in csrc/session.cu for device-side generation of the
4 GB benchmark matrix (no H2D of V).  Values
are k * 2^-24 for k in [1, 2^24], i.e. strictly positive and exactly representable in fp32.
"""
import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_SEEDMUL = np.uint64(0xD1B54A32D192ED03)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def uniform_block(seed, rows, cols, total_rows=None, row0=0, col0=0, dtype=np.float32):
    """rows x cols block (column-major) of the infinite matrix u[i, j] = f(seed, j * total_rows + i)."""
    total_rows = rows if total_rows is None else total_rows
    with np.errstate(over="ignore"):
        i = np.arange(row0, row0 + rows, dtype=np.uint64)[:, None]
        j = np.arange(col0, col0 + cols, dtype=np.uint64)[None, :]
        idx = j * np.uint64(total_rows) + i
        z = (idx + np.uint64(1)) * _GOLD + np.uint64(seed) * _SEEDMUL
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    k = (z >> np.uint64(40)).astype(np.float64) + 1.0
    return np.asfortranarray((k * 2.0 ** -24).astype(dtype))


def cfg1_inputs(dtype=np.float32):
    """BASELINE.json configs[0]: dense 1000x500, k=10, seeds 42/43/44 for V/W0/H0."""
    return dense_inputs(1000, 500, 10, dtype=dtype)


def dense_inputs(m, n, k, seed=42, dtype=np.float32):
    V = uniform_block(seed, m, n, dtype=dtype)
    W0 = uniform_block(seed + 1, m, k, dtype=dtype)
    H0 = uniform_block(seed + 2, k, n, dtype=dtype)
    return V, W0, H0


def planted_inputs(m, n, k, seed=7, noise=0.01, dtype=np.float32):
    """V = W* H* + noise * U: a low-rank problem on which the algorithms visibly converge."""
    Ws = uniform_block(seed + 10, m, k, dtype=np.float64)
    Hs = uniform_block(seed + 11, k, n, dtype=np.float64)
    V = Ws @ Hs + noise * uniform_block(seed + 12, m, n, dtype=np.float64)
    W0 = uniform_block(seed + 1, m, k, dtype=dtype)
    H0 = uniform_block(seed + 2, k, n, dtype=dtype)
    return np.asfortranarray(V.astype(dtype)), W0, H0


def shard_columns(n, world, rank):
    """Column range [c0, c1) of rank `rank` when n columns are split over `world` ranks."""
    base, rem = divmod(n, world)
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)
