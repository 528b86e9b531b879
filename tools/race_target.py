"""Small target for compute-sanitizer racecheck/synccheck on the cooperative kernel (named barriers + shared memory)."""
import hashlib, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelx_b200
rng = random.Random(3)
eng = modelx_b200.Engine(devices=[0], ring_bytes=4 << 20)
msgs = [rng.randbytes(rng.choice([0, 1, 55, 56, 63, 64, 65, 200, 1000, 4096, 20000])) for _ in range(100)]
assert eng.sha256_batch(msgs) == [hashlib.sha256(m).digest() for m in msgs]
uni = [rng.randbytes(8192) for _ in range(64)]
assert eng.sha256_batch(uni) == [hashlib.sha256(m).digest() for m in uni]
h = eng.hasher(); ref = hashlib.sha256()
for n in (10, 5000, 70000):
    p = rng.randbytes(n); h.write(p); ref.update(p); assert h.sum() == ref.digest()
blob = rng.randbytes(300_000)
chunks, root = eng.tree_digest(blob, 1 << 16, 1 << 10, 8)
print("race target ok", root.hex()[:16])
