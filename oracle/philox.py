"""Philox4x32-10 counter-based generator + Box-Muller, restated in numpy (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference draws its Gaussian noise with torch (MultivariateNormal.rsample -> torch.randn), whose stream cannot be
reproduced on a different device or GPU count.  The product therefore draws noise in-kernel with Philox4x32-10
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the generator behind cuRAND's and torch's CUDA
streams) keyed on the GLOBAL element index, and offers mpb_philox_normal to dump exactly what a kernel consumed so the
oracle can replay it.  This file pins the integer generator bit-exactly (Random123 known-answer vectors) and gives
the float64 value of every normal for a tolerance check of the device's fast-intrinsic Box-Muller.

    counter = (group_lo, group_hi, offset_lo, offset_hi),  key = (seed_lo, seed_hi)
    group   = global element index // 4;  the four outputs of one call serve elements 4*group .. 4*group+3
    (u0,u1) -> n0 = r cos(t), n1 = r sin(t),  r = sqrt(-2 ln a), a = u0 * 2^-32 + 2^-33, t = 2 pi (u1 * 2^-32 + 2^-33) - pi
    (u2,u3) -> n2, n3 likewise.
The device forms a and t in fp32 (uint -> float conversion, then ONE fused multiply-add with fp32 constants); that
rounding is part of the specification and is reproduced here exactly, so the only device-vs-oracle difference left is
the accuracy of the MUFU lg2 / sin / cos approximations: |dn| <~ 4e-7 r + 2e-7 / r (the second term matters only
for |n| < 1e-2, where a is within 5e-5 of 1 and lg2's absolute error of 2^-22 is amplified by 1 / r).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Arrays (or scalars) of uint32 -> four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32).copy() for v in (c0, c1, c2, c3))
    k0, k1 = np.asarray(k0, dtype=np.uint32).copy(), np.asarray(k1, dtype=np.uint32).copy()
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = (k0 + W0).astype(np.uint32), (k1 + W1).astype(np.uint32)
    return c0, c1, c2, c3


def philox_u32(seed, offset, first, count):
    """uint32 outputs for global elements first .. first+count-1 (element e uses output e % 4 of group e // 4)."""
    e = np.arange(first, first + count, dtype=np.uint64)
    grp = e >> np.uint64(2)
    lanes = (e & np.uint64(3)).astype(np.int64)
    seed, offset = np.uint64(seed), np.uint64(offset)
    out = philox4x32_10((grp & MASK).astype(np.uint32), (grp >> np.uint64(32)).astype(np.uint32),
                        np.uint32(offset & MASK), np.uint32(offset >> np.uint64(32)),
                        np.uint32(seed & MASK), np.uint32(seed >> np.uint64(32)))
    return np.stack(out, axis=-1)[np.arange(count), lanes], np.stack(out, axis=-1)


def philox_normal(seed, offset, first, count):
    """float64 normals for global elements first .. first+count-1."""
    _, quad = philox_u32(seed, offset, first, count)
    e = np.arange(first, first + count, dtype=np.uint64)
    lane = (e & np.uint64(3)).astype(np.int64)
    uf = quad.astype(np.float32).astype(np.float64)                     # __uint2float_rn
    pair = lane >> 1
    ua, ut = uf[np.arange(count), 2 * pair], uf[np.arange(count), 2 * pair + 1]
    # fmaf(uf, c1, c0) with fp32 constants: exact in float64, then ONE rounding to fp32
    a = (ua * float(np.float32(2.3283064365386963e-10)) + float(np.float32(1.1641532182693481e-10))).astype(np.float32).astype(np.float64)
    t = (ut * float(np.float32(1.4629180792671596e-09)) + float(np.float32(-3.1415925803542134))).astype(np.float32).astype(np.float64)
    r = np.sqrt(-2.0 * np.log(a))
    return np.where((lane & 1) == 0, r * np.cos(t), r * np.sin(t))
