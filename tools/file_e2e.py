"""Developer tool: mxd_tree_digest_file GB/s on a page-cache resident tmpfs file vs MXD_STAGE_THREADS."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelx_b200
path = "/dev/shm/modelx_b200_file_e2e.bin"
size = int(float(os.environ.get("FE_SIZE", 16e9)))
if not os.path.exists(path) or os.path.getsize(path) != size:
    blk = os.urandom(1 << 26)
    with open(path, "wb") as f:
        left = size
        while left > 0:
            n = min(left, len(blk)); f.write(blk[:n]); left -= n
eng = modelx_b200.Engine(devices=[0])
eng.tree_digest_file(path)
t0 = time.perf_counter(); reps = 3
for _ in range(reps): eng.tree_digest_file(path)
dt = (time.perf_counter() - t0) / reps
print(f"MXD_STAGE_THREADS={os.environ.get('MXD_STAGE_THREADS','default')}  {size/dt/1e9:.2f} GB/s", flush=True)
if os.environ.get("FE_CLEAN"): os.unlink(path)
