"""Deterministic synthetic inputs (seeded splitmix64 streams) with identical numpy and torch forms: the torch form fills
multi-GiB inputs on the device in milliseconds, the numpy form feeds the CPU oracle that computes the committed known
answers (tests/golden/make_bench_roots.py).  tests/test_bench_fixture.py checks the two forms against each other.

Element k of stream `seed` is mix((k + 1) * 0x9E3779B97F4A7C15 + seed) with splitmix64's finaliser.
"""
from __future__ import annotations

import numpy as np

_GOLD = 0x9E3779B97F4A7C15
_M1, _M2 = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB
P63 = 5102708120182849537


def _i64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def splitmix_np(seed: int, n: int, start: int = 0) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + n + 1, dtype=np.uint64) * np.uint64(_GOLD) + np.uint64(seed)
        z = (idx ^ (idx >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        return z ^ (z >> np.uint64(31))


def splitmix_torch(seed: int, n: int, device, start: int = 0):
    """The same stream as int64 bit patterns (two's complement wrap-around is the uint64 arithmetic)."""
    import torch

    def lsr(x, k):  # logical shift right on int64
        return (x >> k) & ((1 << (64 - k)) - 1)

    idx = torch.arange(start + 1, start + n + 1, dtype=torch.int64, device=device) * _i64(_GOLD) + _i64(seed)
    z = (idx ^ lsr(idx, 30)) * _i64(_M1)
    z = (z ^ lsr(z, 27)) * _i64(_M2)
    return z ^ lsr(z, 31)


def ft63_np(seed: int, n: int, start: int = 0) -> np.ndarray:
    """Reduced Ft63 elements (Montgomery limbs), shape (n, 1): the stream masked to 63 bits, minus p when >= p."""
    z = splitmix_np(seed, n, start) & np.uint64((1 << 63) - 1)
    return np.where(z >= np.uint64(P63), z - np.uint64(P63), z).reshape(n, 1)


def ft63_torch(seed: int, n: int, device, start: int = 0):
    import torch

    z = splitmix_torch(seed, n, device, start) & ((1 << 63) - 1)
    return torch.where(z >= P63, z - P63, z)


def ft255_np(seed: int, n: int) -> np.ndarray:
    """Reduced Ft255 / Ft253_192-sized elements, shape (n, 4): limbs 0..2 are stream words 4k..4k+2, the top limb is word
    4k+3 shifted right by 3 (below 2^61, hence below the top limb of both four-limb moduli)."""
    z = splitmix_np(seed, 4 * n).reshape(n, 4).copy()
    z[:, 3] >>= np.uint64(3)
    return z


def ft255_torch(seed: int, n: int, device):
    z = splitmix_torch(seed, 4 * n, device).view(n, 4)
    z[:, 3] = (z[:, 3] >> 3) & ((1 << 61) - 1)
    return z.reshape(-1)


def bytes_np(seed: int, n: int) -> np.ndarray:
    """n file bytes: the little-endian bytes of the stream."""
    return splitmix_np(seed, (n + 7) // 8).view(np.uint8)[:n].copy()


def bytes_torch(seed: int, n: int, device, start_byte: int = 0):
    """Bytes [start_byte, start_byte + n) of the same file (start_byte a multiple of 8)."""
    import torch

    assert start_byte % 8 == 0
    return splitmix_torch(seed, (n + 7) // 8, device, start_byte // 8).view(torch.uint8)[:n]
