"""numpy restatement of the device sampler (TEST INFRASTRUCTURE ONLY).

Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as
1, 2, 3", SC'11; constants as in Random123 v1.14 philox.h) — the counter-based
stream that replaces the reference's global CGAL::Random in
TTessel::split_sample (reference lib/ttessel.cpp:1783-1795, :1968-1991).
Known-answer vectors are Random123's kat_vectors for philox4x32-10.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32).copy() for v in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u53(hi, lo):
    return ((hi >> np.uint32(5)).astype(np.float64) * 67108864.0
            + (lo >> np.uint32(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def sample_splits(cum, cum_he, n_batch, seed, ctr0, j0, n):
    """Candidates j0 .. j0+n-1 of a batch of n_batch: returns (he, x, U).
    cum / cum_he: cumulative lengths of the inside halfedges in SLOT order (faces in
    index order, each in ccb order from outer_ccb) and their ids, built as
    lite_tess_upload does (tests/util.py:inside_cum).  One Philox call per
    candidate: words 0-1 -> x, words 2-3 -> U (53 bits each); the 22 bits those
    leave unused place the candidate inside its stratum of width total/n_batch."""
    j = np.arange(j0, j0 + n, dtype=np.uint64)
    ctr = np.uint64(ctr0) + j
    lo = (ctr & MASK).astype(np.uint32)
    hi = (ctr >> np.uint64(32)).astype(np.uint32)
    z = np.zeros(n, dtype=np.uint32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    r = philox4x32_10(lo, hi, z, z, k0, k1)
    x = u53(r[0], r[1])
    U = u53(r[2], r[3])
    spare = (((r[0] & np.uint32(31)) << np.uint32(17)) | ((r[1] & np.uint32(63)) << np.uint32(11))
             | ((r[2] & np.uint32(31)) << np.uint32(6)) | (r[3] & np.uint32(63)))
    jit = spare.astype(np.float64) * (1.0 / 4194304.0)
    sel = (j.astype(np.float64) + jit) * (cum[-1] / float(n_batch))
    idx = np.searchsorted(cum, sel, side="right")  # first m with cum[m] > sel
    idx = np.minimum(idx, len(cum) - 1)
    return cum_he[idx], x, U
