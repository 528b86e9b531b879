"""Developer tool: slot timeline of one file -> tree digest pass (overlap evidence without nsys).
Uses mxd_trace_*: per ring slot the host fill time and CUDA-event times of its H2D copy and its leaf kernel.
Prints a summary and the first slots; the raw CSV goes to $TL_CSV (default gpurun_out/r2_slot_timeline.csv)."""
import csv, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modelx_b200

size = int(float(os.environ.get("TL_GB", "8")) * 1e9)
out = os.environ.get("TL_CSV", "gpurun_out/r2_slot_timeline.csv")
path = f"/dev/shm/modelx_b200_tl_{os.getpid()}.bin"
eng = modelx_b200.Engine(devices=[0])
try:
    buf = torch.empty(size // 8 * 8, dtype=torch.uint8, device="cuda")
    eng.dev_gen_fill(0, buf.data_ptr(), 0, buf.numel(), 11)
    with open(path, "wb") as f:
        f.write(memoryview(buf.cpu().numpy()))
    del buf
    eng.tree_digest_file(path)                       # warm-up
    eng.trace_enable(True)
    t0 = time.perf_counter()
    eng.tree_digest_file(path)
    wall = time.perf_counter() - t0
    eng.trace_enable(False)
    eng.trace_dump(out)
    rows = list(csv.DictReader(open(out)))
    f = lambda r, k: float(r[k])
    span = max(f(r, "kernel_end_ms") for r in rows) - min(f(r, "h2d_start_ms") for r in rows)
    copy = sum(f(r, "h2d_end_ms") - f(r, "h2d_start_ms") for r in rows)
    kern = sum(f(r, "kernel_end_ms") - f(r, "kernel_start_ms") for r in rows)
    fill = sum(f(r, "host_fill_ms") for r in rows)
    # kernel time that overlaps some H2D copy of a LATER slot
    ov = 0.0
    for i, r in enumerate(rows):
        k0, k1 = f(r, "kernel_start_ms"), f(r, "kernel_end_ms")
        for q in rows[i + 1:i + 4]:
            c0, c1 = f(q, "h2d_start_ms"), f(q, "h2d_end_ms")
            ov += max(0.0, min(k1, c1) - max(k0, c0))
    print(f"file {size/1e9:g} GB, {len(rows)} ring slots of {int(rows[0]['bytes'])>>20} MiB; wall {wall*1e3:.1f} ms = {size/wall/1e9:.1f} GB/s")
    print(f"device span {span:.1f} ms | sum H2D {copy:.1f} ms ({copy/span*100:.0f}% of span) | sum leaf kernels {kern:.1f} ms ({kern/span*100:.0f}%) | "
          f"kernel time overlapped by a later slot's H2D {ov:.1f} ms ({ov/max(kern,1e-9)*100:.0f}% of kernel time) | host fill (pread into pinned slot) {fill:.1f} ms")
    print("slot  fill_ms   h2d_start   h2d_end  kern_start  kern_end   (ms since the first copy started)")
    for r in rows[:12]:
        print(f"{r['slot']:>4} {f(r,'host_fill_ms'):8.3f} {f(r,'h2d_start_ms'):10.3f} {f(r,'h2d_end_ms'):9.3f} {f(r,'kernel_start_ms'):11.3f} {f(r,'kernel_end_ms'):9.3f}")
finally:
    if os.path.exists(path):
        os.unlink(path)
    eng.close()
