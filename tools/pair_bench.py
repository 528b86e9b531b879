"""Developer tool: per-chain speed of the whole-message kernels for few chains (k_sha256_chains_pair vs
k_sha256_chains_coop; MXD_TUNE_PAIR=0 disables the pair kernel).  Digests of the first and last chain are checked
against hashlib."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import modelx_b200

eng = modelx_b200.Engine(devices=[0])
total = 32_000_000_000
buf = torch.empty(total, dtype=torch.uint8, device="cuda")
eng.dev_gen_fill(0, buf.data_ptr(), 0, total, 7)
torch.cuda.synchronize()
tag = f"PAIR={os.environ.get('MXD_TUNE_PAIR', 'default')}"
for n in [int(x) for x in os.environ.get('PAIR_BENCH_N', '1,16,32,256,1000,2368,4736,4737,9472').split(',')]:
    size = min(16_000_000, total // n // 64 * 64) + 37           # ragged tail: pad block inside the timed run
    spans = np.zeros((n, 2), dtype=np.uint64)
    spans[:, 0] = buf.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(size - 37)
    spans[:, 1] = size
    d_spans = torch.from_numpy(spans.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    eng.dev_sha256_batch(0, d_spans.data_ptr(), n, d_out.data_ptr()); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.dev_sha256_batch(0, d_spans.data_ptr(), n, d_out.data_ptr()); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = d_out.cpu().numpy().tobytes()
    for k in (() if os.environ.get('PAIR_BENCH_NOCHECK') else (0, n - 1)):
        o = k * (size - 37)
        want = hashlib.sha256(buf[o:o + size].cpu().numpy().tobytes()).digest()
        assert out[32 * k:32 * k + 32] == want, (n, k)
    print(f"{tag} chains={n:6d} x {size/1e6:8.2f} MB  {ms:9.2f} ms  {n*size/ms/1e6:8.2f} GB/s  per-chain {size/ms/1e3:7.1f} MB/s  "
          f"{ms*1e-3*1.965e9/(size/64):7.1f} clk/block  ({'digests not checked' if os.environ.get('PAIR_BENCH_NOCHECK') else 'digests ok'})", flush=True)
