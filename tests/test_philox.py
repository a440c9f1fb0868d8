"""Noise source of the CUDA path: Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11), restated in the
oracle and checked against the Random123 known-answer vectors (kat_vectors: philox4x32 10), plus the
u32 -> U[-PI,PI) map that mirrors boost::uniform_real over a 32-bit engine. The
reference's own RNG stream is irreproducible by construction (wall-clock seed, jamming.cpp:36-37),
so RNG parity is distributional; the GPU-vs-oracle comparison of the stream is in test_gpu_parity."""
import numpy as np

from oracle.pyoracle import PI, OracleSim

KAT = [
    ((0x00000000,) * 4, (0x00000000,) * 2, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        assert tuple(OracleSim.philox4x32_10(ctr, key)) == want


def test_uniform_map_is_the_boost_lattice():
    f = OracleSim.lib().orc_u32_to_randuni
    assert f(0) == -PI
    assert f(0xFFFFFFFF) < PI
    assert f(1 << 31) == 0.0
    u = np.array([0, 1, 12345, 1 << 31, 0xFFFFFFFF], dtype=np.uint64)
    want = u.astype(np.float64) / 4294967296.0 * (PI - (-PI)) + (-PI)
    assert [f(int(v)) for v in u] == list(want)


def test_noise_stream_keys_on_id_step_replica():
    a = OracleSim.philox_noise(7, 3, 64)
    assert np.array_equal(a, OracleSim.philox_noise(7, 3, 64))
    assert not np.array_equal(a, OracleSim.philox_noise(7, 4, 64))
    assert not np.array_equal(a, OracleSim.philox_noise(8, 3, 64))
    assert not np.array_equal(a, OracleSim.philox_noise(7, 3, 64, replica=1))
    assert np.array_equal(a[:16], OracleSim.philox_noise(7, 3, 16))           # value depends on id only, not on N
    big = OracleSim.philox_noise(1, 0, 200000)
    assert abs(big.mean()) < 0.02 and abs(big.std() - 2 * PI / np.sqrt(12)) < 0.01
    assert big.min() >= -PI and big.max() < PI
