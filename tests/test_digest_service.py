"""The whole-message digest service (modelx_b200/csrc/mxd_lockstep.cu): the reference's one-SHA-256-per-file semantics
(push.go:149-161, pull.go:115-123) for concurrent callers -- coalescing, per-call cancellation (push.go:150-159,
mbar.go:108-115), bounded open files, per-file status, several ranges of one file in one pass.

Runs against the CPU test double here and against the CUDA build on the B200 (fixture ``any_engine``)."""
import hashlib
import os
import random
import threading
import time

import pytest

import modelx_b200


def _files(tmp_path, sizes, seed=1):
    rng = random.Random(seed)
    paths, want = [], []
    for i, n in enumerate(sizes):
        p = tmp_path / f"f{i}"
        data = rng.randbytes(n)
        p.write_bytes(data)
        paths.append(str(p))
        want.append(hashlib.sha256(data).digest())
    return paths, want


def test_three_concurrent_callers_coalesce_into_shared_rounds(any_engine, tmp_path):
    """Client.Push runs 3 goroutines, each calling Client.digest on one file (push.go:27,34-52).  Through the service
    they are lanes of the same rounds: far fewer kernel launches than three calls one after another, same digests."""
    paths, want = _files(tmp_path, [48_000_000, 48_000_000, 48_000_000])
    # default ring: 64 MiB slots, a stream advances 8 MiB per round -> 6 rounds per file alone, 6-8 for all three together
    with modelx_b200.Engine(devices=[0], lib_path=any_engine._lib._name) as eng:
        l0 = eng.stats()["kernel_launches"]
        for p, w in zip(paths, want):
            assert eng.sha256_file(p)[0] == w
        serial = eng.stats()["kernel_launches"] - l0
        got = [None] * 3
        start = threading.Barrier(3)

        def work(i):
            start.wait()
            got[i] = eng.sha256_file(paths[i])[0]
        l1 = eng.stats()["kernel_launches"]
        ts = [threading.Thread(target=work, args=(i,)) for i in range(3)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        together = eng.stats()["kernel_launches"] - l1
    assert got == want
    assert together < 0.6 * serial, (together, serial)       # 3 lanes per launch instead of 1 (a late joiner costs a few rounds)


def test_cancel_one_of_three_concurrent_digests(any_engine, tmp_path):
    """push.go:150-159: ctx.Done() of ONE digest must fail that call only.  The other two finish with correct digests,
    the canceled call returns CANCELED (never a digest), and nothing sticks to the engine afterwards."""
    paths, want = _files(tmp_path, [60_000_000, 60_000_000, 60_000_000], seed=2)
    with modelx_b200.Engine(devices=[0], ring_bytes=4 << 20, lib_path=any_engine._lib._name) as eng:
        ops = [eng.op() for _ in range(3)]
        res = [None] * 3
        start = threading.Barrier(4)

        def work(i):
            start.wait()
            try:
                res[i] = ("ok", ops[i].sha256_file(paths[i])[0])
            except modelx_b200.MxdError as e:
                res[i] = ("err", e.status)
        ts = [threading.Thread(target=work, args=(i,)) for i in range(3)]
        [t.start() for t in ts]
        start.wait()
        time.sleep(0.003)
        ops[1].cancel()
        [t.join(timeout=120) for t in ts]
        assert not any(t.is_alive() for t in ts)
        assert res[0] == ("ok", want[0]) and res[2] == ("ok", want[2])
        assert res[1] == ("err", -6) or res[1] == ("ok", want[1])      # canceled, or it had already finished -- never a wrong digest
        # the canceled operation stays canceled (a Go context does); siblings and new operations are unaffected
        with pytest.raises(modelx_b200.MxdError) as ei:
            ops[1].sha256(b"abc")
        assert ei.value.status == -6
        assert ops[0].sha256(b"abc") == hashlib.sha256(b"abc").digest()
        assert eng.sha256_file(paths[1])[0] == want[1]
        [o.close() for o in ops]
        assert eng.stats()["open_files"] == 0


def test_shared_operation_cancels_siblings_like_mbar(any_engine, tmp_path):
    """progress/mbar.go:108-115: the first failing blob cancels its siblings through the context they share."""
    paths, want = _files(tmp_path, [40_000_000] * 3, seed=3)
    with modelx_b200.Engine(devices=[0], ring_bytes=4 << 20, lib_path=any_engine._lib._name) as eng, eng.op() as push_ctx:
        res = []

        def work(i):
            try:
                res.append(("ok", push_ctx.sha256_file(paths[i])[0] == want[i]))
            except modelx_b200.MxdError as e:
                res.append(("err", e.status))
        ts = [threading.Thread(target=work, args=(i,)) for i in range(3)]
        [t.start() for t in ts]
        push_ctx.cancel()
        [t.join(timeout=120) for t in ts]
        assert all(r in (("err", -6), ("ok", True)) for r in res) and len(res) == 3
        assert eng.sha256_file(paths[0])[0] == want[0]          # the engine itself is fine


def test_more_files_than_the_fd_limit(backend, tmp_path, monkeypatch):
    """a batch larger than the open-file budget must work; the reference holds 3 files open."""
    sizes = [random.Random(i).randrange(0, 40_000) for i in range(300)] + [2_000_000]
    paths, want = _files(tmp_path, sizes, seed=4)
    monkeypatch.setenv("MXD_MAX_OPEN_FILES", "16")
    peak = [0]
    with modelx_b200.Engine(devices=[0], ring_bytes=4 << 20, lib_path=backend) as eng:
        stop = threading.Event()

        def watch():
            while not stop.is_set():
                peak[0] = max(peak[0], eng.stats()["open_files"])
        t = threading.Thread(target=watch)
        t.start()
        got, got_sizes = eng.sha256_files(paths)
        stop.set(); t.join()
        assert got == want and got_sizes == sizes
        assert peak[0] <= 16 and eng.stats()["open_files"] == 0


def test_per_file_status(any_engine, tmp_path):
    paths, want = _files(tmp_path, [1000, 70_000, 0], seed=5)
    os.mkdir(tmp_path / "adir")
    jobs = [{"path": paths[0]}, {"path": str(tmp_path / "missing")}, {"path": paths[1]}, {"path": str(tmp_path / "adir")},
            {"path": paths[2]}, {"path": paths[1], "ranges": [(0, 70_001)]}]
    res = any_engine.sha256_file_jobs(jobs)
    assert [r["status"] for r in res] == [0, -4, 0, -4, 0, -4]        # one unreadable file does not fail its siblings
    assert res[0]["digests"][0] == want[0] and res[2]["digests"][0] == want[1] and res[4]["digests"][0] == want[2]
    assert res[2]["size"] == 70_000
    with pytest.raises(modelx_b200.MxdError) as ei:                   # the convenience forms report the first failure
        any_engine.sha256_files([paths[0], str(tmp_path / "missing")])
    assert ei.value.status == -4 and "missing" in ei.value.detail


def test_ranges_of_one_file_in_one_pass_with_tee(any_engine, tmp_path):
    """The whole file and every calcParts range (extension_s3.go:99-112) hashed in ONE pass, bytes teed once."""
    size = 23_456_789
    data = random.Random(6).randbytes(size)
    p = tmp_path / "blob"
    p.write_bytes(data)
    parts = modelx_b200.calc_parts(size, 27)
    ranges = [(0, size)] + parts + [(5, 0), (size, 0), (1, 1), (63, 130), (size - 1, 1), (12_345_678, 64)]
    seen = bytearray(size)
    count = [0]
    lock = threading.Lock()

    def sink(off, piece):
        with lock:
            seen[off:off + len(piece)] = piece
            count[0] += len(piece)
    with modelx_b200.Engine(devices=[0], ring_bytes=4 << 20, lib_path=any_engine._lib._name) as eng:
        r0 = eng.stats()["src_bytes_read"]
        digs, got_size = eng.sha256_file_ranges(str(p), ranges, sink)
        read = eng.stats()["src_bytes_read"] - r0
    assert got_size == size and digs == [hashlib.sha256(data[o:o + n]).digest() for o, n in ranges]
    assert bytes(seen) == data and count[0] == size                   # every byte handed to the tee exactly once
    assert size <= read < size * 1.002                                # one pass (a round re-reads < 64 bytes per unaligned chain)
    with pytest.raises(modelx_b200.MxdError):
        any_engine.sha256_file_ranges(str(p), [(size - 10, 11)])      # a range past EOF is an error, not a short hash


def test_routing_advice():
    """mxd_batch_pays_off: a single whole-file digest never pays off on the GPU (one serial chain), wide batches do."""
    assert not modelx_b200.batch_pays_off(1, 140 * 10 ** 9, 140 * 10 ** 9)      # BASELINE config 4 in compat mode: stay on the CPU
    assert not modelx_b200.batch_pays_off(3, 3 * 10 ** 9, 10 ** 9)
    assert not modelx_b200.batch_pays_off(32, 16 * 10 ** 9, 5 * 10 ** 8)        # config 3 (32 x 0.5 GB): below break-even
    assert modelx_b200.batch_pays_off(1000, 128 * 10 ** 9, 128 * 10 ** 6)       # config 5 (1000 x 128 MB)
    assert modelx_b200.batch_pays_off(256, 256 * 64 * 10 ** 6, 64 * 10 ** 6)
    assert not modelx_b200.batch_pays_off(0, 0, 0)


def test_c_program_cancels_one_of_three(backend, tmp_path):
    """the same contract from plain C through dlopen -- what the cgo shim binds."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "cancel_one_of_three"
    subprocess.run(["gcc", "-O1", "-o", str(exe), os.path.join(root, "tests", "c", "cancel_one_of_three.c"), "-ldl", "-lpthread"], check=True)
    paths, want = _files(tmp_path, [30_000_000, 30_000_000, 30_000_000], seed=9)
    lib = backend or modelx_b200.LIB_PATH
    args = [str(exe), lib]
    for p, w in zip(paths, want):
        args += [p, w.hex()]
    out = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("PASS"), out.stdout + out.stderr


@pytest.mark.parametrize("mapped", ["0", "1"])
def test_file_truncated_while_hashing_is_an_error_not_a_crash(backend, tmp_path, monkeypatch, mapped):
    """A file that shrinks under the hash must come back as an I/O error ("file shrank") on both staging paths: pread
    (default) and the opt-in mapped streaming-store copy (MXD_STAGE_MMAP=1), where it would otherwise be a SIGBUS that
    kills the host process.  (The knob is read once per process: the mapped case runs in a subprocess.)"""
    if mapped == "1":
        import subprocess
        import sys
        env = dict(os.environ, MXD_STAGE_MMAP="1", MXD_TRUNC_CHILD="1")
        out = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", __file__, "-k",
                              "truncated and " + ("mock" if backend else "cuda") + " and 0", "-m", "gpu or not gpu"],
                             env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
        return
    any_engine = modelx_b200.Engine(devices=[0], lib_path=backend)
    p = tmp_path / "shrinking.bin"
    p.write_bytes(os.urandom(40_000_000))
    calls = []

    def sink(off, piece):
        if not calls:
            os.truncate(p, 4096)            # the bytes of the following rounds are gone
        calls.append(off)
    res = any_engine.sha256_file_jobs([{"path": str(p), "sink": sink}])
    assert res[0]["status"] == -4
    p.write_bytes(os.urandom(40_000_000))
    first = []

    def sink2(off, piece):
        if not first:
            os.truncate(p, 4096)
        first.append(off)
    with modelx_b200.Engine(devices=[0], ring_bytes=4 << 20, lib_path=any_engine._lib._name) as small:   # 1 MiB slots: many fills
        with pytest.raises(modelx_b200.MxdError) as ei:
            small.tree_digest_file_tee(str(p), sink2, 1 << 20, 16 << 10, 8)
    assert ei.value.status == -4 and "shrank" in ei.value.detail
    good = tmp_path / "good.bin"            # the engine (and the process) are fine afterwards
    data = os.urandom(3_000_000)
    good.write_bytes(data)
    assert any_engine.sha256_file(str(good))[0] == hashlib.sha256(data).digest()
    any_engine.close()


def test_random_ranges_and_files_against_hashlib(backend, tmp_path):
    """Randomised differential test of the round/window/chain logic: random files, each with random (overlapping, empty,
    unaligned, gapped) ranges, hashed together through a ring so small that every stream needs many rounds; every digest
    must equal hashlib's on those bytes and every tee must see every byte of its file exactly once."""
    rng = random.Random(20260921)
    with modelx_b200.Engine(devices=[0], ring_bytes=4 << 20, lib_path=backend) as eng:      # 1 MiB slots
        for case in range(40):
            jobs, datas, seen = [], [], []
            for k in range(rng.randrange(1, 9)):
                n = rng.choice([0, 1, 63, 64, 65, rng.randrange(0, 5000), rng.randrange(0, 3_000_000)])
                data = rng.randbytes(n)
                p = tmp_path / f"c{case}_f{k}"
                p.write_bytes(data)
                ranges = []
                for _ in range(rng.randrange(0, 7)):
                    off = rng.randrange(0, n + 1)
                    ln = rng.choice([0, min(1, n - off), rng.randrange(0, n - off + 1), n - off])
                    ranges.append((off, ln))
                buf = bytearray(n)
                count = [0]
                lock = threading.Lock()

                def sink(off, piece, buf=buf, count=count, lock=lock):
                    with lock:
                        buf[off:off + len(piece)] = piece
                        count[0] += len(piece)
                use_sink = rng.random() < 0.5
                jobs.append({"path": str(p), "ranges": ranges or None, "sink": sink if use_sink else None})
                datas.append((data, ranges, use_sink)); seen.append((buf, count))
            res = eng.sha256_file_jobs(jobs)
            for r, (data, ranges, use_sink), (buf, count) in zip(res, datas, seen):
                assert r["status"] == 0 and r["size"] == len(data)
                want = [hashlib.sha256(data[o:o + n]).digest() for o, n in ranges] or [hashlib.sha256(data).digest()]
                assert r["digests"][:len(want)] == want, (case, ranges, len(data))
                if use_sink:
                    assert bytes(buf) == data and count[0] == len(data)


def test_mixed_concurrent_load_stays_consistent(backend, tmp_path):
    """Threads mixing tree digests, whole-file jobs, host batches, hashers and cancels on one engine (what a busy
    registry-side verifier would do): every result is either correct or CANCELED for an operation that was canceled."""
    rng = random.Random(7)
    blobs = [rng.randbytes(rng.randrange(1, 4_000_000)) for _ in range(6)]
    paths = []
    for i, b in enumerate(blobs):
        p = tmp_path / f"m{i}"
        p.write_bytes(b)
        paths.append(str(p))
    want_sha = [hashlib.sha256(b).digest() for b in blobs]
    errors = []
    with modelx_b200.Engine(devices=[0], ring_bytes=8 << 20, lib_path=backend) as eng:
        want_tree = [eng.tree_digest(b, 1 << 20, 16 << 10, 8)[1] for b in blobs]

        def worker(seed):
            r = random.Random(seed)
            try:
                for _ in range(12):
                    i = r.randrange(len(blobs))
                    kind = r.randrange(5)
                    if kind == 0:
                        assert eng.tree_digest_file(paths[i], 1 << 20, 16 << 10, 8)[1] == want_tree[i]
                    elif kind == 1:
                        got, _ = eng.sha256_files([paths[i], paths[(i + 1) % len(paths)]])
                        assert got == [want_sha[i], want_sha[(i + 1) % len(paths)]]
                    elif kind == 2:
                        assert eng.sha256_batch([blobs[i][:1000], b"", blobs[i]]) == [hashlib.sha256(blobs[i][:1000]).digest(), hashlib.sha256(b"").digest(), want_sha[i]]
                    elif kind == 3:
                        h = eng.hasher()
                        h.write(blobs[i][:777]); h.write(blobs[i][777:])
                        assert h.sum() == want_sha[i]
                        h.close()
                    else:
                        with eng.op() as op:
                            t = threading.Timer(r.random() * 0.01, op.cancel)
                            t.start()
                            try:
                                assert op.sha256_file(paths[i])[0] == want_sha[i]
                            except modelx_b200.MxdError as e:
                                assert e.status == -6
                            t.join()
            except Exception as e:  # pragma: no cover
                errors.append(repr(e))
        ts = [threading.Thread(target=worker, args=(s,)) for s in range(6)]
        [t.start() for t in ts]
        [t.join(timeout=300) for t in ts]
        assert not any(t.is_alive() for t in ts)
        assert not errors, errors
        assert eng.stats()["open_files"] == 0
