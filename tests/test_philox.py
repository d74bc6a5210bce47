"""Pin oracle/philox.py with the Random123 known-answer vectors (philox4x32-10)."""
import numpy as np

from oracle import philox


def _hex(words):
    return [int(w) for w in words]


def test_random123_known_answers():
    assert _hex(philox.philox4x32_10(0, 0, 0, 0, 0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert _hex(philox.philox4x32_10(f, f, f, f, f, f)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _hex(philox.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344,
                                     0xA4093822, 0x299F31D0)) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_vectorised_matches_scalar():
    ids = np.arange(1000, 1064, dtype=np.uint64)
    w = philox.philox4x32_10(ids, 0, 5, 2, 77, 0)
    for j in (0, 13, 63):
        s = philox.philox4x32_10(int(ids[j]), 0, 5, 2, 77, 0)
        assert [int(x[j]) for x in w] == [int(x) for x in s]


def test_uniforms_are_f32_exact_and_in_range():
    u = philox.reset_uniforms(3, np.arange(4096, dtype=np.uint64), np.full(4096, 9, dtype=np.uint64))
    assert u.shape == (5, 4096)
    assert (u >= 0).all() and (u < 1).all()
    assert np.array_equal(u.astype(np.float32).astype(np.float64), u)
    assert abs(u.mean() - 0.5) < 0.01
    # env ids above 2**32 use the high counter word
    a = philox.reset_uniforms(3, np.array([5], dtype=np.uint64), np.array([1], dtype=np.uint64))
    b = philox.reset_uniforms(3, np.array([5 + (1 << 32)], dtype=np.uint64), np.array([1], dtype=np.uint64))
    assert not np.array_equal(a, b)
