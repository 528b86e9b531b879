"""Developer probe: can a page-cache resident file be fed to the GPU WITHOUT the staging copy?
mmap(tmpfs file) -> cudaHostRegister(window) -> the copy engine reads the page cache directly (zero-copy H2D through
mxd_tree_chunks on a now-pinned pointer) -> cudaHostUnregister.  Prints GB/s of each step so the cost of pinning a
window can be compared with the ~3 GB/s per CPU core of pread-staging (bench e2e is bound by host CPUs: the container
has a 16-CPU cgroup quota)."""
import ctypes, mmap, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import modelx_b200

size = int(float(os.environ.get("PROBE_GB", "8")) * (1 << 30))
path = f"/dev/shm/modelx_b200_probe_{os.getpid()}.bin"
rt = torch.cuda.cudart()
eng = modelx_b200.Engine(devices=[0])
try:
    with open(path, "wb") as f:
        f.truncate(size)
    fd = os.open(path, os.O_RDWR)
    mm = mmap.mmap(fd, size, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
    arr = np.frombuffer(mm, dtype=np.uint8)
    t0 = time.perf_counter(); arr[::4096] = 1; t1 = time.perf_counter()
    print(f"first touch of {size/2**30:.0f} GiB mapping (page faults): {size/(t1-t0)/1e9:.1f} GB/s")
    base = arr.ctypes.data
    want = eng.tree_chunks_ptr(base, size)                     # pageable path (staging through the ring)
    t0 = time.perf_counter(); eng.tree_chunks_ptr(base, size); t1 = time.perf_counter()
    print(f"staged (memcpy into the pinned ring, 16 filler threads): {size/(t1-t0)/1e9:.1f} GB/s")
    for win_mb in (64, 256, 1024, size >> 20):
        win = win_mb << 20
        n = 0; t_reg = t_unreg = 0.0
        off = 0
        while off < size and n < 8:
            w = min(win, size - off)
            t0 = time.perf_counter()
            err = rt.cudaHostRegister(base + off, w, 0)
            t1 = time.perf_counter()
            assert int(err) == 0, err
            err = rt.cudaHostUnregister(base + off)
            t2 = time.perf_counter()
            assert int(err) == 0, err
            t_reg += t1 - t0; t_unreg += t2 - t1; n += 1; off += w
        tot = min(size, n * win)
        print(f"window {win_mb:6d} MiB: cudaHostRegister {tot/t_reg/1e9:7.1f} GB/s   cudaHostUnregister {tot/t_unreg/1e9:7.1f} GB/s   ({n} windows, one thread)")
    t0 = time.perf_counter(); assert int(rt.cudaHostRegister(base, size, 0)) == 0; t1 = time.perf_counter()
    got = eng.tree_chunks_ptr(base, size)
    t2 = time.perf_counter(); got = eng.tree_chunks_ptr(base, size); t3 = time.perf_counter()
    assert got == want
    print(f"whole file registered in {t1-t0:.2f} s ({size/(t1-t0)/1e9:.1f} GB/s); zero-copy digest from the registered page cache: {size/(t3-t2)/1e9:.1f} GB/s")
    rt.cudaHostUnregister(base)
    del arr
    mm.close(); os.close(fd)
finally:
    if os.path.exists(path):
        os.unlink(path)
    eng.close()
