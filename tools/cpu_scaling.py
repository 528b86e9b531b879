"""Developer tool: oracle (CPU) tree-digest throughput vs thread count on the GPU box's host."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.oracle_lib import Oracle
import numpy as np
o = Oracle()
n = int(float(os.environ.get("CS_SIZE", 20e9)))
buf = np.empty(n, dtype=np.uint8)
t = time.perf_counter(); 
# parallel-ish fill: splitmix via oracle in 1 GiB pieces (single thread) would be slow; use numpy random pattern instead
step = 1 << 30
rng = np.random.default_rng(1)
blk = rng.integers(0, 256, size=step, dtype=np.uint8)
for off in range(0, n, step):
    m = min(step, n - off); buf[off:off + m] = blk[:m]
print("fill", time.perf_counter() - t, flush=True)
for th in (1, 8, 16, 32, 64, 128):
    if th == 1:
        m = min(n, 4_000_000_000)
        t = time.perf_counter(); o.sha256_ptr(buf.ctypes.data, m); dt = time.perf_counter() - t
        print(f"threads=1 single chain {m/dt/1e9:.2f} GB/s", flush=True); continue
    t = time.perf_counter(); o.tree_digest_ptr(buf.ctypes.data, n, 8 << 20, 16 << 10, 8, threads=th); dt = time.perf_counter() - t
    print(f"threads={th} tree {n/dt/1e9:.2f} GB/s", flush=True)
