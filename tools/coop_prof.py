import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, modelx_b200
eng = modelx_b200.Engine(devices=[0])
n, size = 1000, 4_000_000
buf = torch.empty(n * size, dtype=torch.uint8, device="cuda")
eng.dev_gen_fill(0, buf.data_ptr(), 0, n * size, 7)
spans = np.zeros((n, 2), dtype=np.uint64)
spans[:, 0] = buf.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(size); spans[:, 1] = size
d_spans = torch.from_numpy(spans.view(np.uint8).reshape(-1)).cuda()
d_out = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.dev_sha256_batch(0, d_spans.data_ptr(), n, d_out.data_ptr()); torch.cuda.synchronize()
