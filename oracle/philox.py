"""numpy Philox4x32-10 — bit-identical to jorldy_b200/csrc/philox.cuh so that env reset draws and
sampling decisions can be compared exactly between the CUDA path and the oracle."""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(seed, stream, ctr):
    """seed: int; stream, ctr: int or uint64 arrays (broadcast).  Returns 4 uint32 arrays."""
    stream = np.asarray(stream, dtype=np.uint64)
    ctr = np.asarray(ctr, dtype=np.uint64)
    stream, ctr = np.broadcast_arrays(stream, ctr)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    c0 = ctr & MASK
    c1 = ctr >> np.uint64(32)
    c2 = stream & MASK
    c3 = stream >> np.uint64(32)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ k0
        n1 = lo1
        n2 = hi0 ^ c3 ^ k1
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = np.uint64((int(k0) + W0) & 0xFFFFFFFF)
        k1 = np.uint64((int(k1) + W1) & 0xFFFFFFFF)
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def u01_double(a, b):
    a = a.astype(np.uint64)
    b = b.astype(np.uint64)
    v = ((a << np.uint64(21)) ^ (b >> np.uint64(11))) & np.uint64((1 << 53) - 1)
    return v.astype(np.float64) * (1.0 / 9007199254740992.0)


def u01_float(a):
    return (a >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
