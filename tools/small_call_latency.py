"""Developer tool: host-call latency of small digests (per-call overheads of the ABI)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelx_b200
eng = modelx_b200.Engine(devices=[0])
def t(fn, n):
    fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
small = b"abc"; kb64 = os.urandom(65536); mb1 = os.urandom(1 << 20); mb64 = os.urandom(64 << 20)
print(f"sha256(3 B)            {t(lambda: eng.sha256(small), 200):8.3f} ms")
print(f"sha256(64 KiB)         {t(lambda: eng.sha256(kb64), 100):8.3f} ms")
print(f"batch 100 x 64 KiB     {t(lambda: eng.sha256_batch([kb64] * 100), 20):8.3f} ms")
print(f"tree_digest(1 MiB)     {t(lambda: eng.tree_digest(mb1), 100):8.3f} ms")
print(f"tree_digest(64 MiB)    {t(lambda: eng.tree_digest(mb64), 20):8.3f} ms   ({64/1024/ (t(lambda: eng.tree_digest(mb64), 10)/1e3):.1f} GB/s)")

import numpy as np
arr = np.frombuffer(mb64, dtype=np.uint8)
print(f"tree_digest(64 MiB np)  {t(lambda: eng.tree_digest(arr), 10):8.3f} ms  (no Python-side copy)")
if os.environ.get("MXD_DEBUG_TIMING"):
    print("--- phases of one tree_digest(64 MiB) call ---", flush=True)
    eng.tree_digest(arr)
