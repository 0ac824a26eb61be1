"""NumPy statement of the counter-based RNG scheme of the kernels (qbold_vi_b200/csrc/rng.cuh,
generate.cu).  TEST INFRASTRUCTURE ONLY.

TensorFlow's own random streams cannot be reproduced offline, so "identical
seeds" is realised as this documented scheme: Philox4x32-10 (Salmon et al., SC'11; the same
generator TF/cuRAND use) keyed by the 64-bit seed, counter = (index_lo, index_hi, stream, 0)
with index = GLOBAL voxel index -- results are invariant to sharding and chunking.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
STREAM_REPARAM, STREAM_KL, STREAM_SNR, STREAM_NOISE, STREAM_MISALIGN = 0, 0x100, 0x10000, 0x10001, 0x20000
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over the counter words (uint32 arrays); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def u01(r):
    """(top 24 bits + 0.5) * 2^-24, float32, in (0, 1)."""
    return ((np.asarray(r, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def box_muller(r0, r1):
    u1, u2 = u01(r0), u01(r1)
    rad = np.sqrt((np.float32(-2.0) * np.log(u1.astype(np.float64)).astype(np.float32)).astype(np.float32))
    ang = (np.float32(2.0) * u2).astype(np.float64)
    return (rad * np.cos(np.pi * ang).astype(np.float32)).astype(np.float32), \
           (rad * np.sin(np.pi * ang).astype(np.float32)).astype(np.float32)


def _words(index, seed):
    index = np.asarray(index, dtype=np.uint64)
    return (index & _MASK32).astype(np.uint32), (index >> np.uint64(32)).astype(np.uint32), \
        np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)


def normal_pair(seed, index, stream):
    lo, hi, k0, k1 = _words(index, seed)
    r = philox4x32_10(lo, hi, np.uint32(stream), np.uint32(0), k0, k1)
    return box_muller(r[0], r[1])


def reparam_eps(seed, index):
    n0, n1 = normal_pair(seed, index, STREAM_REPARAM)
    return np.stack([n0, n1], -1)


def mc_box_muller(r0, r1):
    """Box-Muller as the kernels' Monte-Carlo draws evaluate it (rng.cuh mc_box_muller): radius from log2, angle
    shifted by pi.  The kernel uses the SFU approximations, this is the exact value of the same expressions
    (difference <~ 2e-6 absolute per draw; up to 6e-4 for radii < 0.04)."""
    u1, u2 = u01(r0).astype(np.float64), u01(r1).astype(np.float64)
    rad = np.sqrt(np.maximum(np.log2(u1) * -1.3862943611198906, 0.0))
    x = (u2.astype(np.float32) * np.float32(6.283185307179586) + np.float32(-3.141592653589793)).astype(np.float64)
    return (rad * -np.cos(x)).astype(np.float32), (rad * -np.sin(x)).astype(np.float32)


def kl_eps(seed, index, n_samples):
    """Monte-Carlo sample s of a voxel: Philox call (index, STREAM_KL + (s >> 1)), words (x, y) for even s and
    (z, w) for odd s."""
    out = np.empty((len(index), n_samples, 2), dtype=np.float32)
    lo, hi, k0, k1 = _words(index, seed)
    for c in range((n_samples + 1) // 2):
        r = philox4x32_10(lo, hi, np.uint32(STREAM_KL + c), np.uint32(0), k0, k1)
        out[:, 2 * c, 0], out[:, 2 * c, 1] = mc_box_muller(r[0], r[1])
        if 2 * c + 1 < n_samples:
            out[:, 2 * c + 1, 0], out[:, 2 * c + 1, 1] = mc_box_muller(r[2], r[3])
    return out


def snr_u01(seed, index):
    lo, hi, k0, k1 = _words(index, seed)
    return u01(philox4x32_10(lo, hi, np.uint32(STREAM_SNR), np.uint32(0), k0, k1)[0])


def noise_eps(seed, index, n_tau):
    out = np.zeros((len(index), n_tau + 1), dtype=np.float32)
    for t in range(0, n_tau, 2):
        out[:, t], out[:, t + 1] = normal_pair(seed, index, STREAM_NOISE + (t >> 1))
    return out[:, :n_tau]


def misalign_draws(seed, index, n_tau):
    """Misalignment draws of a voxel (forward.cu k_misalign): one Philox call (index, STREAM_MISALIGN); x -> selection
    uniform, y -> first misaligned image in [4, n_tau - 1), (z, w) -> the OEF / DBV normals."""
    lo, hi, k0, k1 = _words(index, seed)
    r = philox4x32_10(lo, hi, np.uint32(STREAM_MISALIGN), np.uint32(0), k0, k1)
    span = n_tau - 1 - 4
    idx = 4 + np.minimum((u01(r[1]) * np.float32(span)).astype(np.int32), span - 1)
    e0, e1 = box_muller(r[2], r[3])
    return u01(r[0]), idx.astype(np.int32), np.stack([e0, e1], -1)


def _feistel_round(r, key):
    with np.errstate(over='ignore'):
        h = (r * np.uint32(0x9E3779B1) + key).astype(np.uint32)
        h ^= h >> np.uint32(15)
        h = (h * np.uint32(0x85EBCA77)).astype(np.uint32)
        h ^= h >> np.uint32(13)
        h = (h * np.uint32(0xC2B2AE3D)).astype(np.uint32)
        h ^= h >> np.uint32(16)
    return h


def feistel_permute(i, n, seed):
    """Keyed bijection of [0, n) (generate.cu: feistel_permute)."""
    bits = 1
    while bits < 64 and (1 << bits) < n:
        bits += 1
    half = (bits + 1) // 2
    mask = np.uint32(0xFFFFFFFF if half >= 32 else (1 << half) - 1)
    x = np.asarray(i, dtype=np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        xs = x[todo]
        l = ((xs >> np.uint64(half)).astype(np.uint32)) & mask
        r = xs.astype(np.uint32) & mask
        for k in range(4):
            with np.errstate(over='ignore'):
                key = np.uint32((((seed >> (16 * (k & 1))) & 0xFFFFFFFF) + 0x7F4A7C15 * (k + 1) + ((seed >> 32) & 0xFFFFFFFF))
                                & 0xFFFFFFFF)
            l, r = r, l ^ (_feistel_round(r, key) & mask)
        xs = (l.astype(np.uint64) << np.uint64(half)) | r.astype(np.uint64)
        x[todo] = xs
        todo[todo] = xs >= np.uint64(n)
    return x.astype(np.int64)
