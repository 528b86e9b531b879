"""Developer tool: file -> tree digest GB/s (mxd_tree_digest_file on a page-cache resident tmpfs file) under the current
environment (MXD_STAGE_THREADS, MXD_STAGE_PIECE, MXD_RING_BYTES; the MXD_HOST_FEED=map experiment of round 2 -- pinning
the mapped file window by window instead of staging it -- was measured with this tool and removed); wrap in
`taskset -c 0-1` to see the CPU-starved case (8 ranks sharing a 16-CPU quota)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modelx_b200

size = int(float(os.environ.get("FB_GB", "24")) * 1e9)
path = os.environ.get("FB_FILE", "/dev/shm/modelx_b200_feed.bin")
eng = modelx_b200.Engine(devices=[0])
if not os.path.exists(path) or os.path.getsize(path) != size:
    buf = torch.empty(size // 8 * 8 + 8, dtype=torch.uint8, device="cuda")
    eng.dev_gen_fill(0, buf.data_ptr(), 0, buf.numel(), 21)
    with open(path, "wb") as f:
        step = 1 << 30
        for off in range(0, size, step):
            f.write(memoryview(buf[off:min(off + step, size)].cpu().numpy()))
    del buf
eng.tree_digest_file(path)
s0 = eng.stats()
t0 = time.perf_counter(); reps = 3
for _ in range(reps):
    chunks, root, sz = eng.tree_digest_file(path)
dt = (time.perf_counter() - t0) / reps
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("MXD_"))
cpus = len(os.sched_getaffinity(0))
print(f"[{tag or 'defaults'}] cpus={cpus} {size/1e9:g} GB  {dt*1e3:8.1f} ms  {size/dt/1e9:6.1f} GB/s  root {modelx_b200.digest_string(root)[:23]}", flush=True)
