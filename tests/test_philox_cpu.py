"""Philox4x32-10 restatement against Random123's known-answer vectors."""
import numpy as np

import philox_ref


def _one(c, k):
    r = philox_ref.philox4x32_10([c[0]], [c[1]], [c[2]], [c[3]], k[0], k[1])
    return [int(v[0]) for v in r]


def test_kat_vectors():
    # Random123 kat_vectors: philox4x32 10
    assert _one([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _one([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _one([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniforms_in_unit_interval():
    r = philox_ref.philox4x32_10(np.arange(1000), np.zeros(1000), np.zeros(1000), np.zeros(1000), 1, 2)
    u = philox_ref.u53(r[0], r[1])
    assert u.min() >= 0.0 and u.max() < 1.0
