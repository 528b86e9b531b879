"""Developer tool: whole-message (reference-semantics) batch throughput vs number of chains."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import modelx_b200

eng = modelx_b200.Engine(devices=[0])
total = 32_000_000_000
buf = torch.empty(total, dtype=torch.uint8, device="cuda")
eng.dev_gen_fill(0, buf.data_ptr(), 0, total, 7)
torch.cuda.synchronize()
for n, size in ((1, 64_000_000), (32, 128_000_000), (256, 64_000_000), (1000, 32_000_000), (4096, 4_000_000), (8192, 2_000_000), (16384, 1_000_000), (65536, 400_000)):
    spans = np.zeros((n, 2), dtype=np.uint64)
    spans[:, 0] = buf.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(size)
    spans[:, 1] = size
    d_spans = torch.from_numpy(spans.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    eng.dev_sha256_batch(0, d_spans.data_ptr(), n, d_out.data_ptr()); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.dev_sha256_batch(0, d_spans.data_ptr(), n, d_out.data_ptr()); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"COOP={os.environ.get('MXD_TUNE_COOP','default')} chains={n:6d} x {size/1e6:8.1f} MB  {ms:9.2f} ms  {n*size/ms/1e6:8.2f} GB/s  per-chain {size/ms/1e3:7.1f} MB/s", flush=True)
