"""Pin the CPU oracle: FIPS 180-4 / NIST known answers, the one digest constant the reference
holds (EmptyFileDigiest, pkg/client/push.go:25), hashlib (OpenSSL) differential, and the integer
goldens for calcParts / server part count.  CPU only."""
import hashlib
import json
import os
import random

import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

KAT = [
    (b"", "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"),     # EmptyFileDigiest, push.go:25
    (b"abc", "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"),  # FIPS 180-4 example
    (b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq",
     "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"),          # FIPS 180-4 two-block example
    (b"a" * 1000000, "cdc76e5c9914fb9281a1c7e284d73e67f1809a48a497200e046d39ccc7112cd0"),
    (b"a" * 55, "9f4390f8d30c2dd92ec9f095b65e2b9ae9b0a925a5258e241c9f1e910f734318"),
    (b"a" * 56, "b35439a4ac6f0948b6d6f9e3c6af0f5f590ce20f1bde7090ef7970686ec6738a"),
    (b"a" * 63, "7d3e74a05d7db15bce4ad9ec0658ea98e3f06eeecf16b4c6fff2da457ddc2f34"),
    (b"a" * 64, "ffe054fe7ae0cb6dc65c3af9b61d5209f439851db43d0ba5997337df154668eb"),
    (b"a" * 65, "635361c48bb9eab14198e76ea8ab7f1a41685d6ad62aa9146d301d4f17eb0ae0"),
    (b"a" * 119, "31eba51c313a5c08226adf18d4a359cfdfd8d2e816b13f4af952f7ea6584dcfb"),
    (b"a" * 120, "2f3d335432c70b580af0e8e1b3674a7c020d683aa5f73aaaedfdc55af904c21c"),
    (b"a" * 127, "c57e9278af78fa3cab38667bef4ce29d783787a2f731d4e12200270f0c32320a"),
    (b"a" * 128, "6836cf13bac400e9105071cd6af47084dfacad4e5e302c94bfed24e013afb73e"),
]


@pytest.mark.parametrize("engine", [0, 1])
def test_known_answers(oracle, engine):
    if oracle.set_engine(engine) != 0:
        pytest.skip("SHA-NI not available on this CPU")
    try:
        for msg, want in KAT:
            assert oracle.sha256(msg).hex() == want
            assert hashlib.sha256(msg).hexdigest() == want   # hashlib agrees with the published vectors too
    finally:
        oracle.set_engine(-1)


@pytest.mark.parametrize("engine", [0, 1])
def test_differential_vs_hashlib(oracle, engine):
    if oracle.set_engine(engine) != 0:
        pytest.skip("SHA-NI not available on this CPU")
    try:
        rng = random.Random(1234)
        lengths = list(range(0, 300)) + [2 ** k + d for k in range(9, 21) for d in (-1, 0, 1)] + [(8 << 20) + 1]
        for n in lengths:
            msg = rng.randbytes(n)
            assert oracle.sha256(msg) == hashlib.sha256(msg).digest(), n
    finally:
        oracle.set_engine(-1)


def test_incremental_matches_one_shot(oracle):
    rng = random.Random(7)
    msg = rng.randbytes(200_000)
    cuts = sorted(rng.sample(range(len(msg)), 40))
    pieces = [msg[a:b] for a, b in zip([0] + cuts, cuts + [len(msg)])]
    assert oracle.sha256_incremental(pieces) == hashlib.sha256(msg).digest()


def test_client_digest_is_whole_file_sha256(oracle, tmp_path):
    """Client.digest (push.go:149-161) = sha256 of the whole file, 32 KiB read loop."""
    rng = random.Random(3)
    for n in (0, 1, 32768, 32769, 1_000_003):
        p = tmp_path / f"blob{n}"
        data = rng.randbytes(n)
        p.write_bytes(data)
        d, size = oracle.client_digest(str(p))
        assert size == n and d == hashlib.sha256(data).digest()
        s = oracle.digest_string(d)
        assert s == "sha256:" + hashlib.sha256(data).hexdigest()
        assert oracle.pull_file_matches(str(p), s) == 1              # pull.go:120 "already exists"
        assert oracle.pull_file_matches(str(p), "sha256:" + "0" * 64) == 0
    assert oracle.pull_file_matches(str(tmp_path / "missing"), "sha256:" + "0" * 64) == -2   # -ENOENT -> download


def test_empty_file_digest_constant(oracle):
    assert oracle.digest_string(oracle.sha256(b"")) == \
        "sha256:e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"


def test_split_goldens(oracle):
    """calcParts o server part count, table in SURVEY.md section 8a (derived from
    extension_s3.go:99-112 and store_s3.go:198-203,273-279)."""
    with open(os.path.join(GOLDEN, "split_goldens.json")) as f:
        rows = json.load(f)
    for row in rows:
        n = oracle.server_part_count(row["size"])
        assert n == row["parts"], row
        parts = oracle.calc_parts(row["size"], n)
        assert parts[-1] == (row["last_offset"], row["last_length"]), row
        assert sum(l for _, l in parts) == row["size"]
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(n - 1))
    assert oracle.server_part_count(1000, force=True) == 3            # DefaultPartCount
    assert oracle.server_part_count(5 << 30) == 1 and oracle.server_part_count((5 << 30) + 1) == 2
    with pytest.raises(ZeroDivisionError):
        oracle.calc_parts(10, 0)


def test_tree_definition_by_hand(oracle):
    """The oracle's tree digest equals the definition evaluated with hashlib."""
    import struct
    rng = random.Random(11)
    cases = [(0, 64, 2, 1), (1, 64, 2, 1), (64, 64, 2, 2), (1000, 128, 4, 1), (1000, 128, 4, 3),
             (100_000, 1024, 8, 2), (262144, 1024, 16, 1), (300_000, 256, 2, 5)]
    for size, leaf, fanout, k in cases:
        chunk = leaf * fanout ** k
        data = rng.randbytes(size)
        level = [hashlib.sha256(data[i:i + leaf]).digest() for i in range(0, max(size, 1), leaf)]
        chunks, lv = None, 0
        while lv < k or len(level) > 1:
            level = [hashlib.sha256(b"".join(level[i:i + fanout])).digest() for i in range(0, len(level), fanout)]
            lv += 1
            if lv == k:
                chunks = level
        top = level[0]
        root = hashlib.sha256(b"modelx.tree.v1\0\0" + struct.pack("<QQII", size, leaf, fanout, 0) + top).digest()
        got_chunks, got_top, got_root = oracle.tree_digest(data, chunk, leaf, fanout)
        assert got_chunks == chunks and got_top == top and got_root == root, (size, leaf, fanout, k)
        assert len(chunks) == max(1, -(-size // chunk))
    with pytest.raises(ValueError):
        oracle.tree_digest(b"x" * 100, 100, 64, 2)      # chunk is not leaf * fanout**k


def test_generator_is_offset_consistent(oracle):
    whole = oracle.gen(0, 4096, seed=42)
    for off, n in [(0, 1), (1, 7), (5, 100), (8, 64), (1000, 3096), (4095, 1)]:
        assert oracle.gen(off, n, seed=42) == whole[off:off + n]
    assert oracle.gen(0, 64, seed=43) != whole[:64]


def test_tree_golden_vectors(oracle):
    """tests/golden/tree_golden.json was produced with hashlib and a pure-Python splitmix64
    (tests/golden/make_tree_golden.py): pins the generator, whole-message SHA-256 and the tree definition."""
    with open(os.path.join(GOLDEN, "tree_golden.json")) as f:
        rows = json.load(f)
    assert len(rows) >= 9
    for r in rows:
        data = oracle.gen(r["offset"], r["size"], r["seed"])
        assert oracle.sha256(data).hex() == r["sha256"]
        chunks, _, root = oracle.tree_digest(data, r["chunk"], r["leaf"], r["fanout"])
        assert [c.hex() for c in chunks] == r["chunks"] and root.hex() == r["root"], r["size"]
