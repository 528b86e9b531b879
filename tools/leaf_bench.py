"""Developer tool: full tree digest (kernel-only, HBM resident) for A/B of the leaf-kernel variants.
Env: QB_SIZE bytes; variants via MXD_TUNE_LEAF=legacy, MXD_TUNE_LEAF_SCHED=1, MXD_TUNE_LEAF_KPER=n, MXD_TUNE_FUSE=1,
MXD_LEAF_DEBUG=path (per-CTA placement/timing of the persistent launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modelx_b200

size = int(os.environ.get("QB_SIZE", 12_500_000_000))
reps = int(os.environ.get("QB_REPS", 5))
eng = modelx_b200.Engine(devices=[0])
data = torch.empty(size, dtype=torch.uint8, device="cuda")
eng.dev_gen_fill(0, data.data_ptr(), 0, size // 8 * 8, 1)
torch.cuda.synchronize()
chunk, leaf, fan = 8 << 20, 16 << 10, 8
nch = -(-size // chunk)
d_chunks = torch.empty(nch * 32, dtype=torch.uint8, device="cuda"); d_root = torch.empty(32, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    eng.dev_tree_digest(0, data.data_ptr(), size, (chunk, leaf, fan), d_chunks.data_ptr(), d_root.data_ptr(), st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.dev_tree_digest(0, data.data_ptr(), size, (chunk, leaf, fan), d_chunks.data_ptr(), d_root.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("MXD_TUNE"))
print(f"[{tag or 'default'}] size={size/1e9:g} GB  full tree {ms:8.3f} ms  {size/ms/1e6:8.1f} GB/s  root {modelx_b200.digest_string(bytes(d_root.cpu().numpy().tobytes()))[:23]}", flush=True)
