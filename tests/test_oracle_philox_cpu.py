"""CPU: the numpy restatement of the noise generator (oracle/philox.py) against the Random123 known-answer vectors of
Philox4x32-10 and basic distribution checks; the GPU test (tests/test_gpu_noise.py) compares the device against it."""
import numpy as np

from oracle.philox import philox4x32_10, philox_normal, philox_u32

KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = philox4x32_10(*ctr, *key)
        assert tuple(int(v) for v in got) == want


def test_element_indexing_and_moments():
    u, quad = philox_u32(seed=7, offset=3, first=5, count=11)
    assert u.shape == (11,) and quad.shape == (11, 4)
    # element e is output e % 4 of group e // 4: elements 5,6,7 share group 1 with lanes 1,2,3
    assert (quad[0] == quad[1]).all() and (quad[1] == quad[2]).all() and not (quad[2] == quad[3]).all()
    assert u[0] == quad[0, 1] and u[2] == quad[2, 3] and u[3] == quad[3, 0]
    a = philox_normal(11, 0, 0, 1 << 18)
    b = philox_normal(11, 0, 1000, 64)
    assert np.array_equal(a[1000:1064], b), 'a slice of the stream equals the stream started there'
    assert abs(a.mean()) < 1e-2 and abs(a.std() - 1) < 1e-2
    assert abs(np.mean(a ** 3)) < 3e-2 and abs(np.mean(a ** 4) - 3) < 6e-2
    assert not np.array_equal(a[:64], philox_normal(11, 1, 0, 64)), 'the draw counter must change the stream'
    assert not np.array_equal(a[:64], philox_normal(12, 0, 0, 64)), 'the seed must change the stream'
