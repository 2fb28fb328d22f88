"""CPU: Philox4x32-10 restatement against Random123's published known-answer vectors."""
import numpy as np

from oracle import philox


def test_random123_kat():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(np.array([ctr], dtype=np.uint64), key)[0]
        assert tuple(int(v) for v in got) == want


def test_instance_statistics_and_structure():
    A, b, c, x0 = philox.generate_instance(1234, 7, 200, 100)
    assert A.shape == (200, 100) and b.shape == (200,) and c.shape == (100,)
    assert abs(A.mean()) < 0.03 and abs(A.std() - 1.0) < 0.03
    assert (c >= 0).all() and ((b - A.dot(x0)) >= 0).all()       # x0 is strictly feasible by construction
    A2, _, _, _ = philox.generate_instance(1234, 8, 200, 100)
    assert not np.allclose(A, A2)
    A3, _, _, _ = philox.generate_instance(1234, 7, 200, 100)
    assert (A == A3).all()                                       # pure function of (key, instance)
    As, _, _, _ = philox.generate_instance(1234, 7, 200, 100, density=0.1)
    assert 0.07 < (As != 0).mean() < 0.13
    assert (As[As != 0] == A[As != 0]).all()                     # the mask only zeroes entries
