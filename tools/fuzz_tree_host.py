"""Developer tool: randomized differential test of the HOST side of the tree digest (level assembly, ring cursor, file
ranges, tee, sharded finish) against the oracle, through the CPU test double (optionally the ASan build).  Random
(leaf, fanout, chunk = leaf * fanout^k), sizes around every boundary, tiny rings so that every slot wraps."""
import os, random, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import mock_build
from tests.oracle_lib import Oracle
import modelx_b200

lib_path = mock_build.build(os.environ.get("MXD_MOCK_SANITIZE", ""))
orc = Oracle()
rng = random.Random(int(os.environ.get("FUZZ_SEED", "1")))
iters = int(os.environ.get("FUZZ_ITERS", "300"))
work = tempfile.mkdtemp(prefix="mxtree")
engines = {}
def engine(ring):
    if ring not in engines: engines[ring] = modelx_b200.Engine(devices=[0], ring_bytes=ring, lib_path=lib_path)
    return engines[ring]
bad = 0
for it in range(iters):
    leaf = rng.choice([64, 128, 192, 1024, 4096, 16384, 1000, 77])
    fan = rng.choice([2, 3, 4, 8, 16])
    k = rng.randrange(0, 4)
    chunk = leaf * fan ** k
    base = rng.choice([0, 1, leaf, chunk, chunk * fan, chunk * rng.randrange(1, 9), rng.randrange(0, 3_000_000)])
    size = max(0, min(4_000_000, base + rng.choice([-1, 0, 1, rng.randrange(-leaf, leaf + 1)])))
    data = rng.randbytes(size)
    eng = engine(rng.choice([1 << 20, 4 << 20, 64 << 20]))
    try:
        want_chunks, _top, want_root = orc.tree_digest(data, chunk, leaf, fan)
    except ValueError as e:                       # parameters the format rejects: the library must reject them too
        try:
            eng.tree_digest(data, chunk, leaf, fan); print("MISMATCH: oracle rejects, library accepts", leaf, fan, chunk, size, e); bad += 1
        except modelx_b200.MxdError:
            pass
        continue
    if leaf % 64:                                 # the library is stricter than the format's CPU statement: leaves are whole blocks
        try:
            eng.tree_digest(data, chunk, leaf, fan); print("MISMATCH: leaf", leaf, "accepted"); bad += 1
        except modelx_b200.MxdError:
            pass
        continue
    got_chunks, got_root = eng.tree_digest(data, chunk, leaf, fan)
    ok = got_root == want_root and list(got_chunks) == list(want_chunks)
    p = os.path.join(work, "f.bin"); open(p, "wb").write(data)
    fc, fr = eng.tree_digest_file(p, chunk, leaf, fan)[:2]
    ok = ok and fr == want_root and list(fc) == list(want_chunks)
    seen = bytearray(size); cnt = [0]
    def sink(off, piece):
        seen[off:off + len(piece)] = piece; cnt[0] += len(piece)
    tc, tr, tsz = eng.tree_digest_file_tee(p, sink, chunk, leaf, fan)
    ok = ok and tr == want_root and tsz == size and bytes(seen) == data and cnt[0] == size
    # sharded: split the chunk list in two ranges on a chunk boundary, digest each range from the file, finish
    nch = max(1, -(-size // chunk))
    cut = rng.randrange(0, nch + 1)
    parts = b""
    for a, b in ((0, cut), (cut, nch)):
        if b > a:
            off = a * chunk; n = min(size, b * chunk) - off
            parts += eng.tree_chunks_file(p, off, n, chunk, leaf, fan)[: 32 * (b - a)]
    if size > 0:
        ok = ok and parts == b"".join(want_chunks) and eng.tree_finish(parts, size, chunk, leaf, fan) == want_root
    if not ok:
        bad += 1; print("MISMATCH", dict(leaf=leaf, fanout=fan, chunk=chunk, size=size))
for e in engines.values(): e.close()
print(f"tree host fuzz: {iters} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
