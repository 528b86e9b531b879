"""Developer tool: throughput of the incremental hasher (hash.Hash seam, helper.go:46-49): one stream = one chain."""
import os, sys, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelx_b200
eng = modelx_b200.Engine(devices=[0])
buf = os.urandom(64 << 20)
for piece in (32 << 10, 1 << 20, 16 << 20):
    h = eng.hasher(); ref = hashlib.sha256()
    t0 = time.perf_counter(); n = 0
    for rep in range(4):
        for off in range(0, len(buf), piece):
            h.write(buf[off:off + piece]); n += min(piece, len(buf) - off)
    d = h.sum(); dt = time.perf_counter() - t0
    for rep in range(4):
        ref.update(buf)
    assert d == ref.digest()
    print(f"hasher: writes of {piece>>10:6d} KiB  {n/dt/1e6:7.1f} MB/s  (digest ok)", flush=True)
    h.close()
eng.close()
