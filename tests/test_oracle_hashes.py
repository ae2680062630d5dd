"""Pins the oracle's standard hashes (Zig std in the reference) against hashlib / the xxhash package."""
import hashlib
import random

import xxhash


def test_sha3_256_known_answers(po):
    # FIPS 202 known answers
    assert po.sha3_256(b"").hex() == "a7ffc6f8bf1ed76651c14756a061d662f580ff4de43b49fa82d80a4b80f8434a"
    assert po.sha3_256(b"abc").hex() == "3a985da74fe225b2045c172d6bd390bd855f086e3e9d525b46bfe24511431532"


def test_sha3_256_vs_hashlib_all_block_boundaries(po):
    rng = random.Random(1)
    for n in list(range(0, 300)) + [1000, 4096, 136 * 7, 136 * 7 - 1, 136 * 7 + 1]:
        data = bytes(rng.getrandbits(8) for _ in range(n))
        assert po.sha3_256(data) == hashlib.sha3_256(data).digest(), n


def test_sha256_vs_hashlib(po):
    rng = random.Random(2)
    for n in list(range(0, 130)) + [1000]:
        data = bytes(rng.getrandbits(8) for _ in range(n))
        assert po.sha256(data) == hashlib.sha256(data).digest(), n


def test_xxh3_small_inputs_vs_xxhash(po):
    rng = random.Random(3)
    for n in range(0, 17):
        for _ in range(50):
            data = bytes(rng.getrandbits(8) for _ in range(n))
            seed = rng.getrandbits(64) if _ % 2 else 0
            assert po.xxh3_64(data, seed) == xxhash.xxh3_64_intdigest(data, seed=seed), (n, seed)


def test_transcript_is_streaming_sha3(po):
    t = po.Transcript()
    t.append_bytes(b"hello")
    t.append_field(77)
    h = hashlib.sha3_256(b"hello" + (77).to_bytes(8, "little"))
    d = h.digest()
    assert t.challenge(po.BABYBEAR_P) == int.from_bytes(d[:8], "little") % po.BABYBEAR_P
    h.update(d)  # hash.zig:312: the transcript absorbs its own digest
    assert t.challenge(po.BABYBEAR_P) == int.from_bytes(h.digest()[:8], "little") % po.BABYBEAR_P
