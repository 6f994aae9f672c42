"""Host-side restatement of the part of ``jax.random`` the stream generators use.

``jr.key(seed)`` / ``jr.split`` / ``jr.normal`` with jax's default ``threefry2x32`` implementation in
*partitionable* mode (the default since jax 0.5; the reference pins jax 0.8.0, ``/root/reference/uv.lock:1210``), for
float64 draws.  With it, ``FardalStreamDF`` / ``Fardal2015DF`` seeded with an integer reproduce the reference's
``jr.key(seed)`` release conditions (``df/fardal15.py:61,81-84``, ``experimental/df.py:146-163``) instead of numpy's
stream.  Threefry is exact integer arithmetic; the normal transform is ``sqrt(2) * erfinv(u)`` where XLA uses Giles'
polynomial and this module uses ``scipy.special.erfinv`` -- equal to ~1e-16 relative, not bit for bit.

Pinned by the reference's own doctest of ``Fardal2015DF.sample(jr.key(0), ...)`` (``experimental/df.py:110-123``),
see ``tests/test_jaxrandom.py``.
"""

from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = np.uint64(0xFFFFFFFF)


def _rotl(x, r):
    return ((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & _M32


def threefry2x32(k0, k1, c0, c1):
    """Threefry-2x32, 20 rounds (Salmon et al. 2011), vectorised over the counters; uint32 in/out."""
    k0, k1 = np.asarray(k0, dtype=np.uint64), np.asarray(k1, dtype=np.uint64)  # scalars or arrays (broadcast)
    x0 = np.asarray(c0, dtype=np.uint64).copy()
    x1 = np.asarray(c1, dtype=np.uint64).copy()
    ks = (k0, k1, (k0 ^ k1 ^ np.uint64(0x1BD11BDA)) & _M32)
    x0 = (x0 + ks[0]) & _M32
    x1 = (x1 + ks[1]) & _M32
    for g in range(5):
        for r in _ROT[g % 2]:
            x0 = (x0 + x1) & _M32
            x1 = _rotl(x1, r)
            x1 = x1 ^ x0
        x0 = (x0 + ks[(g + 1) % 3]) & _M32
        x1 = (x1 + ks[(g + 2) % 3] + np.uint64(g + 1)) & _M32
    return x0.astype(np.uint32), x1.astype(np.uint32)


def key(seed: int) -> np.ndarray:
    """``jax.random.key(seed)`` -> raw key data [hi32, lo32] (``threefry_seed``)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


def _iota_2x32(shape):
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    idx = np.arange(n, dtype=np.uint64).reshape(shape)
    return (idx >> np.uint64(32)) & _M32, idx & _M32


def split(k: np.ndarray, num: int = 2) -> np.ndarray:
    """``jax.random.split`` (partitionable / "fold-like" split): key i = threefry(key, counter = i)."""
    c1, c2 = _iota_2x32((num,))
    b1, b2 = threefry2x32(k[0], k[1], c1, c2)
    return np.stack([b1, b2], axis=-1)


def random_bits64(k: np.ndarray, shape) -> np.ndarray:
    c1, c2 = _iota_2x32(tuple(shape))
    b1, b2 = threefry2x32(k[0], k[1], c1, c2)
    return (b1.astype(np.uint64) << np.uint64(32)) | b2.astype(np.uint64)


def uniform(k: np.ndarray, shape=(), minval=0.0, maxval=1.0) -> np.ndarray:
    """``jax.random.uniform`` for float64: 52 random mantissa bits -> [1, 2) - 1, then affine map and clamp."""
    bits = random_bits64(k, shape)
    fl = ((bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)).view(np.float64) - 1.0
    return np.maximum(minval, fl * (maxval - minval) + minval)


def normal(k: np.ndarray, shape=()) -> np.ndarray:
    """``jax.random.normal`` for float64: sqrt(2) erfinv(U(nextafter(-1, 0), 1))."""
    from scipy.special import erfinv

    lo = np.nextafter(-1.0, 0.0)
    return np.sqrt(2.0) * erfinv(uniform(k, shape, lo, 1.0))


def _threefry_int(k0: int, k1: int, c0: int, c1: int) -> tuple[int, int]:
    """Scalar threefry2x32 on Python ints (for the sequential key chain below)."""
    M = 0xFFFFFFFF
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0, x1 = (c0 + ks[0]) & M, (c1 + ks[1]) & M
    for g in range(5):
        for r in _ROT[g % 2]:
            x0 = (x0 + x1) & M
            x1 = ((x1 << r) | (x1 >> (32 - r))) & M
            x1 ^= x0
        x0 = (x0 + ks[(g + 1) % 3]) & M
        x1 = (x1 + ks[(g + 2) % 3] + g + 1) & M
    return x0, x1


def split_chain(k: np.ndarray, M: int) -> np.ndarray:
    """The sub-keys of ``for i in range(M): key, subkey = jr.split(key)`` (a ``lax.scan`` carrying the key, as in
    ``experimental/stream.py:212-218``) -> uint32 ``(M, 2)``."""
    a, b = int(k[0]), int(k[1])
    out = np.empty((M, 2), dtype=np.uint32)
    for i in range(M):
        n0 = _threefry_int(a, b, 0, 0)  # split(key)[0] -> next carry
        s0 = _threefry_int(a, b, 0, 1)  # split(key)[1] -> sub-key
        out[i] = s0
        a, b = n0
    return out


def fardal_draws_per_key(subkeys: np.ndarray) -> np.ndarray:
    """One scalar-shaped Fardal sample per sub-key (``df.sample(subkey, ...)`` with ``x`` of shape (3,)) -> (4, M)."""
    subkeys = np.asarray(subkeys, dtype=np.uint32)
    M = subkeys.shape[0]
    out = np.empty((4, M))
    from scipy.special import erfinv

    lo = np.nextafter(-1.0, 0.0)
    for i in range(4):  # key_i = split(subkey, 4)[i] = threefry(subkey, counter i); normal(key_i, ()) uses counter 0
        ki0, ki1 = threefry2x32(subkeys[:, 0], subkeys[:, 1], np.zeros(M), np.full(M, i))
        b1, b2 = threefry2x32(ki0, ki1, np.zeros(M), np.zeros(M))
        bits = (b1.astype(np.uint64) << np.uint64(32)) | b2.astype(np.uint64)
        fl = ((bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)).view(np.float64) - 1.0
        out[i] = np.sqrt(2.0) * erfinv(np.maximum(lo, fl * (1.0 - lo) + lo))
    return out


def fardal_draws(seed_or_key, M: int) -> np.ndarray:
    """The four normal draws of ``FardalStreamDF._sample`` (``jr.split(key, 4)``, each ``jr.normal(key_i, (M, 1))``)
    as the ``(4, M)`` array the release kernel takes."""
    k = key(seed_or_key) if np.ndim(seed_or_key) == 0 else np.asarray(seed_or_key, dtype=np.uint32)
    ks = split(k, 4)
    return np.stack([normal(ks[i], (M,)) for i in range(4)])


def multivariate_normal_svd(seed_or_key, mean, cov, shape) -> np.ndarray:
    """``jax.random.multivariate_normal(key, mean, cov, shape, method="svd")``: ``(u, s, _) = svd(cov)``,
    ``factor = u * sqrt(s)``, ``mean + factor @ normal(key, shape + (d,))``.

    The singular vectors' signs and order are the SVD routine's choice: XLA:CPU and numpy both call LAPACK's
    ``gesdd``, so this reproduces the reference on CPU as far as the two LAPACK builds agree (XLA:GPU uses a Jacobi
    solver with its own conventions -- the reference's own draws differ between its backends).  Unverified offline.
    """
    k = key(seed_or_key) if np.ndim(seed_or_key) == 0 else np.asarray(seed_or_key, dtype=np.uint32)
    mean = np.asarray(mean, dtype=np.float64)
    u, s, _ = np.linalg.svd(np.asarray(cov, dtype=np.float64))
    factor = u * np.sqrt(s)[None, :]
    z = normal(k, tuple(shape) + mean.shape[-1:])
    return mean + np.einsum("ij,...j->...i", factor, z)


def chen_draws(seed_or_key, M: int) -> np.ndarray:
    """The ``(M, 6)`` samples of ``ChenStreamDF._sample`` (df/chen24.py:89-91) for ``jr.key(seed)``."""
    from .dynamics import ChenStreamDF

    return multivariate_normal_svd(seed_or_key, ChenStreamDF.mean, ChenStreamDF.cov, (M,))
