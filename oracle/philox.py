"""TEST INFRASTRUCTURE: numpy Philox4x32-10 and the stream convention of the CUDA env.

The reference draws from numpy's process-global MT19937 (``np.random.rand()``,
drone.py:57, :73), unseeded, so its random numbers cannot be reproduced by a
counter-based generator.  Parity for resets is therefore defined by *injecting our
uniforms into the reference* (oracle/ref_import.py: patched_rand) -- five per reset in the
reference's order pos.x, pos.y, tgt.x, tgt.y, tgt.z.

Algorithm: Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11; the Random123 library).  Pinned by the Random123 known-answer vectors in
tests/test_philox.py.

Stream convention (must match drone_rl_b200/csrc/philox.cuh):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (env_id & 0xffffffff, env_id >> 32, index, stream)
    stream 0 (RESET)   index = ep_num of the new episode -> the top 24 bits of the four words give
                       pos.x, pos.y, tgt.x, tgt.y; the fifth uniform (tgt.z) is assembled from the
                       otherwise unused low bytes of words 0..2: (w0&255) | (w1&255)<<8 | (w2&255)<<16
                       (one Philox call per reset: 120 of its 128 random bits are used)
    stream 2 (ACTION)  index = global step t             -> 4 motor uniforms (random policy)
    stream 3 (NOISE)   index = global step t             -> 4 uniforms -> 4 Box-Muller normals
    uniform u = (word >> 8) * 2**-24  in [0, 1)   (exact in both float32 and float64)
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85

STREAM_RESET = 0
STREAM_ACTION = 2
STREAM_NOISE = 3

_MASK32 = np.uint64(0xFFFFFFFF)
_SH32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All arguments broadcastable uint32 arrays/ints.

    Returns a tuple of four uint32 arrays.
    """
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0  # 32x32 -> 64 bit product, exact in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> _SH32, p0 & _MASK32
        hi1, lo1 = p1 >> _SH32, p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(words):
    """uint32 -> float64 uniform in [0,1) with 24 random bits (exactly float32-representable)."""
    return (np.asarray(words, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) * (2.0 ** -24)


def env_stream(seed: int, env_ids, index, stream: int):
    """Four uniforms per env for (seed, env_id, index, stream).  Returns float64 [4, ...]."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    idx = np.asarray(index, dtype=np.uint64)
    c3 = (np.uint64(stream) | ((idx >> _SH32) << np.uint64(8))) & _MASK32
    w = philox4x32_10(env_ids & _MASK32, env_ids >> _SH32, idx & _MASK32, c3,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack([u01(x) for x in w])


def reset_uniforms(seed: int, env_ids, ep_num):
    """The five reset uniforms [5, n] in the reference's draw order (drone.py:57, :73)."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    idx = np.asarray(ep_num, dtype=np.uint64)
    c3 = (np.uint64(STREAM_RESET) | ((idx >> _SH32) << np.uint64(8))) & _MASK32
    w = philox4x32_10(env_ids & _MASK32, env_ids >> _SH32, idx & _MASK32, c3,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lo = (w[0] & np.uint32(255)) | ((w[1] & np.uint32(255)) << np.uint32(8)) | ((w[2] & np.uint32(255)) << np.uint32(16))
    fifth = lo.astype(np.float64) * (2.0 ** -24)
    return np.stack([u01(w[0]), u01(w[1]), u01(w[2]), u01(w[3]), fifth])


def action_uniforms(seed: int, env_ids, t):
    """Random-policy motor uniforms [n, 4] for global step ``t`` (float64, f32-exact)."""
    return env_stream(seed, env_ids, t, STREAM_ACTION).T.copy()


def noise_normals(seed: int, env_ids, t):
    """Four standard normals per env [n, 4] (float64 Box-Muller of the NOISE stream).

    z0 = r(u0) cos(2 pi u1), z1 = r(u0) sin(2 pi u1), z2 = r(u2) cos(2 pi u3), z3 = r(u2) sin(2 pi u3)
    with r(u) = sqrt(-2 ln(u + 2**-24))  (u + 2**-24 is in (0, 1]).
    """
    u = env_stream(seed, env_ids, t, STREAM_NOISE)
    r0 = np.sqrt(-2.0 * np.log(u[0] + 2.0 ** -24))
    r1 = np.sqrt(-2.0 * np.log(u[2] + 2.0 ** -24))
    a0 = 2.0 * np.pi * u[1]
    a1 = 2.0 * np.pi * u[3]
    return np.stack([r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1), r1 * np.sin(a1)], axis=1)


def minibatch_permutation(n: int, seed: int, epoch: int) -> np.ndarray:
    """The keyed permutation of 0..n-1 behind ``dronecu_minibatch_permutation`` (csrc/ppo_update.cuh
    ``perm_apply``): on k = ceil(log2 n) bits, four rounds of x = ((x ^ (x >> s)) * odd + add) mod 2^k with
    s = (k + 1) // 2 and the constants from Philox4x32-10(key = seed, counter = (epoch lo, epoch hi, r, "PERM")),
    cycle-walked until the value is < n.  Stands in for SB3's ``np.random.permutation`` (PARITY UNPINNED: any
    uniform-looking order satisfies SB3's contract)."""
    bits = 1
    while (1 << bits) < n:
        bits += 1
    mask = (1 << bits) - 1
    shift = (bits + 1) // 2
    mul, add = [0] * 4, [0] * 4
    for r in (0, 2):
        c = philox4x32_10(epoch & 0xFFFFFFFF, (epoch >> 32) & 0xFFFFFFFF, r, 0x5045524D,
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        mul[r], mul[r + 1], add[r], add[r + 1] = int(c[0]) | 1, int(c[1]) | 1, int(c[2]), int(c[3])
    x = np.arange(n, dtype=np.uint64)
    todo = np.ones(n, dtype=bool)
    while todo.any():
        y = x[todo]
        for r in range(4):
            y ^= y >> np.uint64(shift)
            y = (y * np.uint64(mul[r]) + np.uint64(add[r])) & np.uint64(mask)
        x[todo] = y
        todo[todo] = y >= n
    return x.astype(np.int64)


def minibatch_partition(n: int, batch: int, seed: int, epoch: int) -> np.ndarray:
    """``dronecu_minibatch_partition``: row r belongs to minibatch ``minibatch_permutation(n, seed, epoch)[r] // batch``; the
    result lists minibatch 0's rows in ascending order, then minibatch 1's, ... (a stable sort by minibatch id)."""
    key = minibatch_permutation(n, seed, epoch) // batch
    return np.argsort(key, kind="stable").astype(np.int64)
