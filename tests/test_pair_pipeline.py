"""The two-lanes-per-chain round pipeline of k_sha256_chains_pair, checked as an algorithm on the CPU (the kernel itself is
checked bit for bit by the -m gpu parity tests: every batch of at most 4,736 messages runs through it)."""
import hashlib
import random

from tests.pair_pipeline_emulation import sha256_pair


def test_pair_pipeline_matches_sha256():
    rng = random.Random(5)
    for n in [0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 129, 1000, 4096, 16384]:
        m = rng.randbytes(n)
        assert sha256_pair(m) == hashlib.sha256(m).digest(), n
    assert sha256_pair(b"abc").hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
