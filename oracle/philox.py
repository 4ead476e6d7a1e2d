"""CPU oracle, part 4: the counter-based random streams of the B200 samplers.

TEST INFRASTRUCTURE ONLY (see oracle/runtime.py for the rules).

The reference draws from numpy's global Mersenne-Twister state
(metropolis_hastings.py:2,137-142), which cannot be split over thousands of
lock-step chains.  The B200 samplers either consume streams injected by the
caller (the parity mode) or generate Philox4x32-10 streams on the device
(Salmon et al., SC'11 -- published algorithm, known-answer vectors from the
Random123 distribution are checked in tests/test_oracle_golden.py).  This file
is the numpy statement of that generator and of the stream layout in
``cosmoslik_b200/csrc/philox.cuh``:

    key      = (seed_lo, seed_hi)
    counter  = (index, sub, step, stream_id)

  stream_id 1 (MH normals)  index = chain,  sub = pair p  -> z[2p], z[2p+1]
  stream_id 2 (MH uniform)  index = chain,  sub = 0       -> u
  stream_id 3 (stretch)     index = global walker, sub = half
                              words (0,1) -> u_z, (2,3) -> u_acc
  stream_id 4 (partner)     index = global walker, sub = half
                              j = (word0 * Nc) >> 32
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_MH_Z = 1
STREAM_MH_U = 2
STREAM_STRETCH = 3
STREAM_PARTNER = 4


def philox4x32(counter, key, rounds=10):
    """counter: 4 broadcastable uint32 arrays; key: 2 uint32 scalars/arrays.
    Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(*counter)]
    k0 = int(key[0]) & 0xFFFFFFFF
    k1 = int(key[1]) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(hi, lo):
    """Two 32-bit words -> double in (0, 1]:
    ((hi >> 5) * 2^26 + (lo >> 6) + 0.5) * 2^-53.  Never 0 (the samplers take its log); for
    m >= 2^52 the +0.5 is absorbed by rounding and m = 2^53 - 1 gives exactly 1.0."""
    hi = np.asarray(hi, dtype=np.uint64)
    lo = np.asarray(lo, dtype=np.uint64)
    m = (hi >> np.uint64(5)) * np.uint64(1 << 26) + (lo >> np.uint64(6))
    return (m.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def box_muller(u1, u2):
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2.0 * np.pi * u2), r * np.sin(2.0 * np.pi * u2)


def mh_streams(seed, step, chains, ndim):
    """z[chains, ndim], u[chains] for one MH step."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    chains = np.asarray(chains, dtype=np.uint32)
    npair = (ndim + 1) // 2
    z = np.empty((len(chains), 2 * npair))
    for p in range(npair):
        w = philox4x32((chains, np.uint32(p), np.uint32(step), np.uint32(STREAM_MH_Z)), key)
        z0, z1 = box_muller(u01(w[0], w[1]), u01(w[2], w[3]))
        z[:, 2 * p], z[:, 2 * p + 1] = z0, z1
    w = philox4x32((chains, np.uint32(0), np.uint32(step), np.uint32(STREAM_MH_U)), key)
    return z[:, :ndim], u01(w[0], w[1])


def stretch_streams(seed, step, half, walkers, ncomp):
    """u_z, j, u_acc for the walkers (GLOBAL indices) of one half-step."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    walkers = np.asarray(walkers, dtype=np.uint32)
    w = philox4x32((walkers, np.uint32(half), np.uint32(step), np.uint32(STREAM_STRETCH)), key)
    v = philox4x32((walkers, np.uint32(half), np.uint32(step), np.uint32(STREAM_PARTNER)), key)
    j = (v[0].astype(np.uint64) * np.uint64(ncomp)) >> np.uint64(32)
    return u01(w[0], w[1]), j.astype(np.int64), u01(w[2], w[3])
