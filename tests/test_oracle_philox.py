"""Oracle pin: Philox4x32-10 against the Random123 known-answer vectors
(kat_vectors of the Random123 distribution, philox4x32 10 rounds)."""
import numpy as np

from oracle import philox as px


def _kat(ctr, key):
    return [int(v) for v in px.philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1])]


def test_random123_known_answers():
    assert _kat([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_u52_open_interval_and_exact():
    assert px.u52(0, 0) == 2.0 ** -53
    assert px.u52(0xffffffff, 0xffffffff) == 1.0 - 2.0 ** -53
    a = np.arange(0, 2 ** 32, 2 ** 27, dtype=np.uint64)
    u = px.u52(a, a[::-1])
    assert np.all((u > 0.0) & (u < 1.0))


def test_vectorised_matches_scalar():
    c0 = np.arange(17, dtype=np.uint64)
    ua, ub = px.uniforms(12345678901234567, c0, 3, 5, px.GL_TAG_MOVE)
    for i in range(17):
        a, b = px.uniforms(12345678901234567, int(c0[i]), 3, 5, px.GL_TAG_MOVE)
        assert a == ua[i] and b == ub[i]
