"""Developer tool: kernel-only GB/s of the leaf kernel for a few leaf sizes / occupancy variants."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modelx_b200

size = int(os.environ.get("QB_SIZE", 10_000_000_000))
eng = modelx_b200.Engine(devices=[0])
data = torch.empty(size, dtype=torch.uint8, device="cuda")
eng.dev_gen_fill(0, data.data_ptr(), 0, size // 8 * 8, 1)
torch.cuda.synchronize()
chunk = 8 << 20
for leaf in [int(x) for x in os.environ.get('QB_LEAVES', '4096,16384,65536,262144').split(',')]:
    nl = -(-size // leaf)
    out = torch.empty(nl * 32, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        eng.dev_sha256_segments(0, data.data_ptr(), size, leaf, out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.dev_sha256_segments(0, data.data_ptr(), size, leaf, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"MINB={os.environ.get('MXD_TUNE_MINB','8')} leaf={leaf:>7} lanes={nl:>8}  {ms:8.3f} ms  {size/ms/1e6:8.1f} GB/s", flush=True)
    # full tree
    fan = 8 if (chunk // leaf) in (8, 64, 512, 4096) else 2
    nch = -(-size // chunk)
    d_chunks = torch.empty(nch * 32, dtype=torch.uint8, device="cuda"); d_root = torch.empty(32, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        eng.dev_tree_digest(0, data.data_ptr(), size, (chunk, leaf, fan), d_chunks.data_ptr(), d_root.data_ptr(), st)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        eng.dev_tree_digest(0, data.data_ptr(), size, (chunk, leaf, fan), d_chunks.data_ptr(), d_root.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"     full tree (chunk 8 MiB fan {fan})          {ms:8.3f} ms  {size/ms/1e6:8.1f} GB/s", flush=True)
