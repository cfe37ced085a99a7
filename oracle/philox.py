"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) restated in NumPy.

TEST INFRASTRUCTURE ONLY (see oracle/mip360_oracle.py).  The reference draws its random numbers with torch.rand /
Tensor.uniform_ (intern/ray.py:106 and :33); the CUDA path generates them inside the consuming kernels
(mipnerf360_b200/csrc/common.cuh: philox_first_word, rng_uniform).  This file reproduces those uniforms bit for bit so
that the randomized kernels can be checked against the oracle fed the same numbers.

Convention: number m of ray r in a draw: counter = (r, (m & 7) | ((m >> 5) << 3), stream_id, epoch),
key = (seed & 0xffffffff, seed >> 32), u = (output word (m >> 3) & 3 >> 8) * 2**-24 — one Philox call yields the numbers
m, m + 8, m + 16, m + 24 of a ray.  Pinned by the known-answer vectors of the Random123 distribution
(tests/test_oracle_properties_cpu.py)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: 4 arrays of uint32-valued integers (broadcastable), key: 2 -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in counter]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def uniform(seed, stream_id, epoch, shape):
    """The float32 uniforms in [0, 1) the kernels draw for a [rays, numbers-per-ray] tensor."""
    B, K = shape
    r, m = np.meshgrid(np.arange(B, dtype=np.uint64), np.arange(K, dtype=np.uint64), indexing="ij")
    block = (m & np.uint64(7)) | ((m >> np.uint64(5)) << np.uint64(3))
    words = philox4x32_10((r, block, np.uint64(stream_id), np.uint64(epoch)), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    sel = ((m >> np.uint64(3)) & np.uint64(3)).astype(np.int64)
    w = np.choose(sel, words)
    return (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
