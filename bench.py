#!/usr/bin/env python
"""bench.py -- GB/s SHA-256-digested on a 100 GB synthetic blob (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port)

A "step" is one full digest of the blob: leaf SHA-256 over every 16 KiB of blob bytes, the tree
levels above them up to the 8 MiB chunk-digest list, the levels above that, and the root.

  value   whole-job GB/s with the blob already resident in HBM (CUDA-event timed, max over ranks)
  e2e     the same digest through the C ABI with the blob in pinned HOST memory: H2D copies of every
          blob byte and D2H of the chunk list + root are inside the timed region
  N > 1   strong scaling: the 100 GB blob is sharded by chunk index (rank r owns a contiguous chunk
          range); no collective on the hash path, one NCCL all-gather of the 32-byte chunk digests,
          then every rank finishes the (tiny) upper levels.

Only the cpu_baseline / --impl reference legs touch oracle/ (the CPU checker); the timed GPU legs
call libmodelxdigest.so through modelx_b200.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x6D6F64656C78  # "modelx"
GB = 1e9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=float, default=100e9, help="blob bytes (default: the 100 GB metric blob)")
    ap.add_argument("--chunk", type=int, default=8 << 20)
    ap.add_argument("--leaf", type=int, default=16 << 10)
    ap.add_argument("--fanout", type=int, default=8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample-gb", type=float, default=16.0)
    ap.add_argument("--ref-sample-gb", type=float, default=4.0)
    ap.add_argument("--file-gb", type=float, default=16.0, help="size of the tmpfs file for the e2e_file figure")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """DRAM bytes per algorithmic byte of the leaf kernel from the committed ncu capture (or None)."""
    p = os.path.join(ROOT, "profiles", "ncu_leaf_kernel.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [c for c, p in zip(sm, power) if p > 250] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ==================================================================================================
# reference arm: the reference's own CPU implementation of the path, restated in oracle/ (the Go
# toolchain is absent).  Client.digest (pkg/client/push.go:149-161): one goroutine,
# one serial SHA-256 chain over the whole file through a 32 KiB read loop.
# ==================================================================================================
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests.oracle_lib import Oracle
    orc = Oracle()
    nbytes = int(args.ref_sample_gb * (1 << 30))
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(shm, f"modelx_b200_ref_sample_{os.getpid()}.bin")
    piece = 256 << 20
    buf = ctypes.create_string_buffer(piece)
    with open(path, "wb") as f:                      # untimed: materialise the sample of the metric blob
        off = 0
        while off < nbytes:
            n = min(piece, nbytes - off)
            orc.gen_into(ctypes.addressof(buf), off, n, SEED)
            f.write(buf.raw[:n] if n < piece else buf.raw)
            off += n
    try:
        for _ in range(args.warmup):
            orc.client_digest(path)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            d, size = orc.client_digest(path)
        dt = time.perf_counter() - t0
    finally:
        os.unlink(path)
    val = nbytes * args.steps / dt / GB
    sample = (f"first {args.ref_sample_gb:g} GiB of the {args.size/1e9:g} GB metric blob as a tmpfs file per step; "
              "Client.digest = 32 KiB read loop + SHA-256 (SHA-NI), one thread: a blob is one serial chain, "
              "the reference cannot use more cores for it")
    line = {
        "impl": "reference", "metric": "GB/s SHA-256-digested on 100 GB synthetic blob", "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"{args.size/1e9:g} GB blob, whole-file sha256 (reference semantics, push.go:149-161)",
                   "blob_bytes": int(args.size), "sample_bytes": nbytes},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample,
                         "engine": "sha-ni" if orc.engine() == 1 else "portable", "host_cpus": os.cpu_count()},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "digest": orc.digest_string(d),
    }
    print(json.dumps(line), flush=True)


# ==================================================================================================
# this repo's arm
# ==================================================================================================
def run_b200(args):
    import torch
    import torch.distributed as dist
    import modelx_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    size = int(args.size)
    tp = (args.chunk, args.leaf, args.fanout)
    from modelx_b200 import shard
    nchunks = shard.chunk_count(size, args.chunk)
    per = shard.chunks_per_rank(nchunks, world)      # chunks per rank (only tail ranks may own fewer)
    c0, c1 = shard.chunk_range(rank, world, nchunks)
    b0, b1 = shard.byte_range(rank, world, size, args.chunk)
    my_bytes = b1 - b0
    my_chunks = c1 - c0

    eng = modelx_b200.Engine(devices=[local])
    stream = torch.cuda.current_stream().cuda_stream
    blob = torch.empty(max(my_bytes, 8), dtype=torch.uint8, device=dev)
    fill = (my_bytes + 7) // 8 * 8
    if fill > blob.numel():
        blob = torch.empty(fill, dtype=torch.uint8, device=dev)
    eng.dev_gen_fill(0, blob.data_ptr(), b0, fill, SEED, stream)   # bytes [b0, b1) of the one logical blob
    d_local = torch.zeros(per * 32, dtype=torch.uint8, device=dev)
    d_all = torch.zeros(world * per * 32, dtype=torch.uint8, device=dev)
    d_root = torch.zeros(32, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if my_bytes or world == 1:
            eng.dev_tree_chunks(0, blob.data_ptr(), my_bytes, tp, d_local.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_local)       # 32 B per chunk; the only exchange on the path
            eng.dev_tree_finish(0, d_all.data_ptr(), nchunks, size, tp, d_root.data_ptr(), stream)
        else:
            eng.dev_tree_finish(0, d_local.data_ptr(), nchunks, size, tp, d_root.data_ptr(), stream)

    # ---- kernel-only: blob resident in HBM ----------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.prof_enable(True)
    st0 = eng.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    prof = eng.prof_read()
    eng.prof_enable(False)
    st1 = eng.stats()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = size * args.steps / (ms_max * 1e-3) / GB
    root_dev = bytes(d_root.cpu().numpy().tobytes())
    chunk_list_dev = bytes((d_all if world > 1 else d_local).cpu().numpy().tobytes())[:nchunks * 32]
    launches = st1["kernel_launches"] - st0["kernel_launches"]

    # roofline of the dominant kernel: the leaf-level launch (reads every blob byte once)
    peak, peak_src = load_peaks()
    leaf_ms = prof["kernel_ms"] / max(prof["launches"], 1)
    leaf_bytes = prof["bytes"] / max(prof["launches"], 1)
    achieved = leaf_bytes / (leaf_ms * 1e-3) / GB if leaf_ms > 0 else 0.0
    tr = load_traffic()
    # the bound that actually binds: ALU-pipe issue (2 warp-instr/clk/SM measured; 1,056 ALU-pipe instructions per
    # 64-byte block of 32 lanes: 672 SHF + 352 LOP3 + 16 PRMT + 16 loop/address) at the SM clock seen under load
    props = torch.cuda.get_device_properties(dev)
    sm_clk_hz = (clocks["sm_mhz"] if (clocks and clocks.get("sm_mhz")) else 1965.0) * 1e6
    alu_ceiling = 4 * 32 * 64 / (1056 * 2.0) * props.multi_processor_count * sm_clk_hz / GB
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (tr["dram_bytes_per_algorithmic_byte"] * leaf_bytes) if tr else None,
                "kernel": "k_sha256_lanes (leaf level)", "kernel_ms_per_launch": leaf_ms,
                "algorithmic_bytes_per_launch": leaf_bytes, "kernel_share_of_step": prof["kernel_ms"] / ms if ms else None,
                "peak_source": peak_src,
                "alu_pipe": {"ceiling": alu_ceiling, "unit": "GB/s", "frac": achieved / alu_ceiling if alu_ceiling else None,
                             "how": "4 SMSP x 32 lanes x 64 B / (1056 ALU-pipe instr x 2 clk) x SMs x SM clock; pipe rates "
                                    "measured in profiles/r01_pipes_ubench.txt"},
                "note": ("SHA-256 is 1 B read per B digested but ~16.5 INT32 ALU-pipe instructions per byte; the binding "
                         "limit is the ALU pipe (2 warp-instr/clk/SM measured), ceiling ~1.13 TB/s = 17% of HBM peak; "
                         "see DESIGN.md section 5")}

    # ---- end to end: blob in pinned host memory, through the public C ABI ------------------------
    e2e = None
    host_ptr = 0
    if not args.no_e2e:
        t_pin = time.perf_counter()
        host_ptr = eng.host_alloc(max(my_bytes, 1))
        t_pin = time.perf_counter() - t_pin
        if my_bytes:
            # same bytes as the HBM-resident blob (D2H is setup, not timed)
            host_view = (ctypes.c_uint8 * my_bytes).from_address(host_ptr)
            import numpy as np
            host_np = np.frombuffer(host_view, dtype=np.uint8)
            host_t = torch.from_numpy(host_np)
            host_t.copy_(blob[:my_bytes])            # destination is pinned by mxd_host_alloc
            torch.cuda.synchronize()
        del blob
        torch.cuda.empty_cache()

        def e2e_step():
            mine = eng.tree_chunks_ptr(host_ptr, my_bytes, *tp)[:my_chunks * 32] if (my_bytes or world == 1) else b""
            if world > 1:
                d_local.zero_()
                if mine:
                    d_local[:len(mine)].copy_(torch.frombuffer(bytearray(mine), dtype=torch.uint8))
                dist.all_gather_into_tensor(d_all, d_local)
                allc = bytes(d_all.cpu().numpy().tobytes())[:nchunks * 32]
            else:
                allc = mine
            return allc, eng.tree_finish(allc, size, *tp)

        e2e_warm = min(args.warmup, 3)
        for _ in range(e2e_warm):
            e2e_step()
        barrier()
        s0 = eng.stats()
        eng.prof_enable(True)            # device time of the leaf launches inside the e2e region (overlap evidence)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            allc, root_e2e = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        s1 = eng.stats()
        prof_e2e = eng.prof_read()
        eng.prof_enable(False)
        barrier()
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        hb = torch.tensor([s1["h2d_bytes"] - s0["h2d_bytes"], s1["d2h_bytes"] - s0["d2h_bytes"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(hb, op=dist.ReduceOp.SUM)
        dt_max = float(t.item())
        if root_e2e != root_dev or allc != chunk_list_dev:
            raise SystemExit("e2e digest differs from the HBM-resident digest of the same bytes")
        e2e = {"value": size * args.steps / dt_max / GB, "unit": "GB/s",
               "h2d_bytes_per_step": int(hb[0].item() / args.steps), "d2h_bytes_per_step": int(hb[1].item() / args.steps),
               "ms_per_step": dt_max / args.steps * 1e3, "warmup": e2e_warm,
               "leaf_kernel_ms_per_step": prof_e2e["kernel_ms"] / args.steps, "leaf_launches_per_step": prof_e2e["launches"] / args.steps,
               "overlap": "copy engine and SMs run concurrently: leaf-kernel device time per step is hidden behind the H2D "
                          "copies (compare leaf_kernel_ms_per_step with ms_per_step; ring of 4 x 64 MiB slots, 2 streams)",
               "api": "mxd_tree_chunks(host ptr) [+ NCCL all-gather] + mxd_tree_finish",
               "host_memory": f"pinned (mxd_host_alloc, {t_pin:.1f} s to pin, untimed setup)"}

    # ---- file path (what the Go client would call): mxd_tree_digest_file on a tmpfs file, N=1 only ----
    e2e_file = None
    if rank == 0 and world == 1 and host_ptr and my_bytes and not args.no_e2e and os.path.isdir("/dev/shm"):
        fbytes = int(min(args.file_gb * 1e9, my_bytes))
        fpath = f"/dev/shm/modelx_b200_bench_{os.getpid()}.bin"
        try:
            with open(fpath, "wb") as f:
                view = (ctypes.c_uint8 * fbytes).from_address(host_ptr)
                f.write(memoryview(view))
            eng.tree_digest_file(fpath, *tp)                                   # warm-up
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                fchunks, froot, fsz = eng.tree_digest_file(fpath, *tp)
            dtf = (time.perf_counter() - t0) / reps
            e2e_file = {"value": fbytes / dtf / GB, "unit": "GB/s", "bytes": fbytes,
                        "api": "mxd_tree_digest_file (parallel pread into the pinned ring, page-cache resident tmpfs file)"}
        finally:
            if os.path.exists(fpath):
                os.unlink(fpath)

    # ---- CPU baseline on rank 0's host cores (N=1 only): oracle port, bounded sample ------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from tests.oracle_lib import Oracle
        orc = Oracle()
        if host_ptr and my_bytes:
            sample = int(min(args.cpu_sample_gb * 1e9, my_bytes))
            t0 = time.perf_counter()
            d_one = orc.sha256_ptr(host_ptr, sample)                 # reference semantics: one serial chain, 1 thread
            dt1 = time.perf_counter() - t0
            t0 = time.perf_counter()                                   # the reference's PullPushConcurrency = 3 (push.go:27)
            orc.tree_digest_ptr(host_ptr, sample, *tp, threads=3)
            dt3 = time.perf_counter() - t0
            threads = min(os.cpu_count() or 1, 32)   # plateaus at 16-32 threads on the bench box (profiles/r01_cpu_scaling.txt)
            t0 = time.perf_counter()
            want_chunks, _, want_root = orc.tree_digest_ptr(host_ptr, my_bytes, *tp, threads=threads)
            dtn = time.perf_counter() - t0
            parity = (want_root == root_dev and b"".join(want_chunks) == chunk_list_dev)
            if not parity:
                raise SystemExit("GPU tree digest differs from the CPU oracle on the full blob")
            cpu = {"value": sample / dt1 / GB, "unit": "GB/s", "cores": 1, "kind": "port",
                   "sample": f"first {sample/1e9:g} GB of the blob, one SHA-256 chain on one thread (what the reference does "
                             "for one blob, push.go:149-161), SHA-NI, data already in memory (no read syscalls)",
                   "engine": "sha-ni" if orc.engine() == 1 else "portable", "host_cpus": os.cpu_count(),
                   "three_threads_tree": {"value": sample / dt3 / GB, "unit": "GB/s", "cores": 3,
                                          "sample": f"first {sample/1e9:g} GB, tree digest on 3 threads (the reference's PullPushConcurrency)"},
                   "all_cores_tree": {"value": my_bytes / dtn / GB, "unit": "GB/s", "cores": threads,
                                      "sample": "the whole blob, same tree digest chunk-parallel on host threads (32: the "
                                                "measured plateau, more threads are slower on this box)",
                                      "parity_with_gpu": parity}}
    if host_ptr:
        eng.host_free(host_ptr)

    if rank == 0:
        line = {
            "metric": "GB/s SHA-256-digested on 100 GB synthetic blob", "value": value, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"{size/1e9:g} GB blob (splitmix64 counter stream) resident in HBM, modelx.tree.v1 digest",
                       "blob_bytes": size, "chunk": args.chunk, "leaf": args.leaf, "fanout": args.fanout,
                       "chunks": nchunks, "parallelism": f"chunk-range sharding x{world}" if world > 1 else "single GPU",
                       "l2": "input per GPU >> 126 MB L2, read once per step (no flush needed)"},
            "clocks": clocks, "e2e": e2e, "e2e_file": e2e_file, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "root": modelx_b200.digest_string(root_dev),
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    import __graft_entry__ as g
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        g.build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
