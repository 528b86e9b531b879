"""GPU parity: every digest the CUDA path produces, called through the C ABI, must equal the CPU
oracle bit for bit on the same bytes.  Run on the B200 box: pytest -m gpu."""
import hashlib
import os
import random
import threading

import numpy as np
import pytest

import modelx_b200
from tests.test_oracle_pinning import KAT

pytestmark = pytest.mark.gpu

SEED = 0x6D6F64656C78  # "modelx"


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


# ------------------------------------------------------------------------------------------------
# whole-message digests (the reference's semantics: digest.FromBytes / FromReader)
# ------------------------------------------------------------------------------------------------
def test_known_answers_single_and_batch(engine):
    for msg, want in KAT:
        assert engine.sha256(msg).hex() == want
    got = engine.sha256_batch([m for m, _ in KAT])
    assert [g.hex() for g in got] == [w for _, w in KAT]
    assert modelx_b200.digest_string(engine.sha256(b"")) == \
        "sha256:e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"   # EmptyFileDigiest, push.go:25


def test_ragged_batch_vs_oracle(engine, oracle):
    rng = random.Random(2024)
    lengths = list(range(0, 300)) + [2 ** k + d for k in range(9, 21) for d in (-1, 0, 1)] + \
        [(8 << 20) - 1, 8 << 20, (8 << 20) + 1]
    rng.shuffle(lengths)
    msgs = [rng.randbytes(n) for n in lengths]
    got = engine.sha256_batch(msgs)
    for m, g in zip(msgs, got):
        assert g == oracle.sha256(m), len(m)


def test_empty_batch_and_empty_message(engine):
    assert engine.sha256_batch([]) == []
    assert engine.sha256_batch([b"", b""]) == [hashlib.sha256(b"").digest()] * 2


def test_device_spans_all_alignments(engine, oracle):
    """Device-resident messages at every byte alignment (exercises the unaligned load path)."""
    torch = _torch()
    rng = np.random.default_rng(5)
    host = rng.integers(0, 256, size=1 << 20, dtype=np.uint8)
    dev = torch.from_numpy(host).cuda()
    base = dev.data_ptr()
    spans, want = [], []
    for off in list(range(0, 20)) + [31, 33, 4097]:
        for n in (0, 1, 3, 55, 56, 64, 65, 127, 128, 1000, 65537):
            spans.append((base + off, n))
            want.append(oracle.sha256(host[off:off + n].tobytes()))
    got = engine.sha256_batch_ptrs(spans)
    assert got == want
    # aligned-only batch (fast path): every span 16-byte aligned
    spans = [(base + 16 * i, 4096 + 64 * i + (i % 5)) for i in range(100)]
    want = [oracle.sha256(host[16 * i:16 * i + 4096 + 64 * i + (i % 5)].tobytes()) for i in range(100)]
    assert engine.sha256_batch_ptrs(spans) == want


def test_verify_batch(engine, oracle):
    msgs = [os.urandom(n) for n in (0, 10, 1000, 70000)]
    want = [hashlib.sha256(m).digest() for m in msgs]
    want[2] = bytes(32)
    assert engine.verify_batch(msgs, want) == [True, True, False, True]


def test_long_message_bit_length_over_2_32(engine, oracle):
    """One 600 MB message: its bit length (4.8e9) needs the high word of the 64-bit length field."""
    torch = _torch()
    n = 600_000_000 + 13
    dev = torch.empty(n + 3, dtype=torch.uint8, device="cuda")
    engine.dev_gen_fill(0, dev.data_ptr(), 0, (n + 3) // 8 * 8, SEED)
    torch.cuda.synchronize()
    host = dev.cpu().numpy()
    got = engine.sha256_ptr(dev.data_ptr(), n)
    assert got == oracle.sha256_ptr(host.ctypes.data, n)


# ------------------------------------------------------------------------------------------------
# incremental hasher (hash.Hash shape, helper.go:46-49)
# ------------------------------------------------------------------------------------------------
def test_hasher_incremental(engine):
    rng = random.Random(99)
    h = engine.hasher()
    ref = hashlib.sha256()
    assert h.sum() == ref.digest()
    total = 0
    for n in [0, 1, 63, 64, 65, 1000, 4 << 20, (4 << 20) + 1, 3, 9_000_000, 17]:
        piece = rng.randbytes(n)
        h.write(piece)
        ref.update(piece)
        total += n
        assert h.sum() == ref.digest()       # Sum must not disturb the running state
        assert h.sum() == ref.digest()
        assert h.written() == total and h.size() == 32 and h.block_size() == 64
    h.reset()
    assert h.sum() == hashlib.sha256(b"").digest() and h.written() == 0
    h.write(b"abc")
    assert h.sum().hex() == KAT[1][1]
    h.close()


# ------------------------------------------------------------------------------------------------
# files: Client.digest (push.go:149-161), pullFile check (pull.go:115-123)
# ------------------------------------------------------------------------------------------------
def test_file_digests_match_reference_path(engine, oracle, tmp_path):
    rng = random.Random(31)
    paths, datas = [], []
    for i, n in enumerate([0, 1, 64, 32768, 32769, 1_000_003, 5_000_000, 70_000_001]):
        p = tmp_path / f"blob{i}.bin"
        d = rng.randbytes(n)
        p.write_bytes(d)
        paths.append(str(p))
        datas.append(d)
    for p, d in zip(paths, datas):
        got, size = engine.sha256_file(p)
        want, wsize = oracle.client_digest(p)
        assert (got, size) == (want, wsize) and size == len(d)
    got, sizes = engine.sha256_files(paths)
    assert got == [hashlib.sha256(d).digest() for d in datas] and sizes == [len(d) for d in datas]
    want = [hashlib.sha256(d).digest() for d in datas]
    want[3] = bytes(32)
    ok = engine.verify_files(paths, want)
    assert ok == [i != 3 for i in range(len(paths))]
    with pytest.raises(modelx_b200.MxdError) as ei:
        engine.sha256_file(str(tmp_path / "missing.bin"))
    assert ei.value.status == -4   # MXD_ERR_IO, errno preserved in the message


# ------------------------------------------------------------------------------------------------
# chunked tree digest
# ------------------------------------------------------------------------------------------------
TREE_CASES = [  # (size, chunk, leaf, fanout)
    (0, 128, 64, 2), (1, 128, 64, 2), (63, 128, 64, 2), (64, 128, 64, 2), (65, 128, 64, 2), (128, 128, 64, 2),
    (129, 128, 64, 2), (1000, 256, 64, 4), (1000, 256, 64, 2), (100_000, 4096, 1024, 4), (262_144, 16384, 1024, 16),
    (262_145, 16384, 1024, 4), ((1 << 20) - 1, 1 << 18, 1 << 12, 8), (1 << 20, 1 << 18, 1 << 12, 64),
    ((1 << 20) + 1, 1 << 18, 1 << 12, 2), (20_000_003, 1 << 20, 16 << 10, 8), (50_000_000, 8 << 20, 16 << 10, 8),
    (50_000_000, 8 << 20, 16 << 10, 512), (50_000_000, 8 << 20, 64 << 10, 2), (50_000_000, 8 << 20, 4 << 10, 2048),
]


@pytest.mark.parametrize("size,chunk,leaf,fanout", TREE_CASES)
def test_tree_digest_host_vs_oracle(engine, oracle, size, chunk, leaf, fanout):
    blob = oracle.gen(7, size, SEED)
    chunks, root = engine.tree_digest(blob, chunk, leaf, fanout)
    want_chunks, _, want_root = oracle.tree_digest(blob, chunk, leaf, fanout)
    assert chunks == want_chunks
    assert root == want_root


def test_tree_default_params(engine, oracle):
    blob = oracle.gen(0, 20_000_000, SEED + 9)
    chunks, root = engine.tree_digest(blob)          # 8 MiB chunks, 16 KiB leaves, fan-out 8
    want_chunks, _, want_root = oracle.tree_digest(blob, 8 << 20, 16 << 10, 8)
    assert (chunks, root) == (want_chunks, want_root) and len(chunks) == 3


def test_tree_digest_streams_through_small_ring(oracle):
    """A ring far smaller than the blob: many slots, slot boundaries inside chunks."""
    size = 40_000_000 + 77
    blob = oracle.gen(0, size, SEED + 1)
    with modelx_b200.Engine(devices=[0], ring_bytes=4 << 20) as eng:
        chunks, root = eng.tree_digest(blob, 8 << 20, 16 << 10, 8)
        st = eng.stats()
    want_chunks, _, want_root = oracle.tree_digest(blob, 8 << 20, 16 << 10, 8)
    assert chunks == want_chunks and root == want_root
    assert st["h2d_bytes"] >= size and st["kernel_launches"] >= 40


def test_tree_digest_file_and_pinned(engine, oracle, tmp_path):
    size = 30_000_000 + 5
    blob = oracle.gen(0, size, SEED + 2)
    p = tmp_path / "model.safetensors"
    p.write_bytes(blob)
    want_chunks, _, want_root = oracle.tree_digest(blob, 8 << 20, 16 << 10, 8)
    chunks, root, sz = engine.tree_digest_file(str(p))
    assert (chunks, root, sz) == (want_chunks, want_root, size)
    # same bytes from pinned host memory (zero-copy H2D path)
    import ctypes
    ptr = engine.host_alloc(size)
    ctypes.memmove(ptr, blob, size)
    chunks2, root2 = engine.tree_digest_ptr(ptr, size)
    engine.host_free(ptr)
    assert (chunks2, root2) == (want_chunks, want_root)


def test_tree_sharded_equals_whole(engine, oracle):
    """Chunk ranges hashed independently (what each rank does) + finish == one-shot digest."""
    size = 9 * (1 << 20) + 4321
    chunk, leaf, fanout = 1 << 20, 16 << 10, 8
    blob = oracle.gen(0, size, SEED + 3)
    whole_chunks, whole_root = engine.tree_digest(blob, chunk, leaf, fanout)
    nch = len(whole_chunks)
    for world in (2, 3, 4):
        gathered = b""
        for r in range(world):
            c0, c1 = nch * r // world, nch * (r + 1) // world
            piece = blob[c0 * chunk:min(c1 * chunk, size)]
            gathered += engine.tree_chunks(piece, chunk, leaf, fanout)[:32 * (c1 - c0)]
        assert gathered == b"".join(whole_chunks)
        assert engine.tree_finish(gathered, size, chunk, leaf, fanout) == whole_root


def test_device_generator_matches_oracle(engine, oracle):
    torch = _torch()
    n = 1 << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    engine.dev_gen_fill(0, dev.data_ptr(), 4096, n, SEED)
    torch.cuda.synchronize()
    assert dev.cpu().numpy().tobytes() == oracle.gen(4096, n, SEED)


def test_device_resident_tree_async(engine, oracle):
    """mxd_dev_* forms on the caller's stream: data generated in HBM, digests left in HBM."""
    torch = _torch()
    size = 100_000_000
    chunk, leaf, fanout = 8 << 20, 16 << 10, 8
    nch = -(-size // chunk)
    data = torch.empty(size, dtype=torch.uint8, device="cuda")
    d_chunks = torch.empty(nch * 32, dtype=torch.uint8, device="cuda")
    d_root = torch.empty(32, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    engine.dev_gen_fill(0, data.data_ptr(), 0, size, SEED, st)
    engine.dev_tree_digest(0, data.data_ptr(), size, (chunk, leaf, fanout), d_chunks.data_ptr(), d_root.data_ptr(), st)
    torch.cuda.synchronize()
    host = data.cpu().numpy()
    want_chunks, _, want_root = oracle.tree_digest_ptr(host.ctypes.data, size, chunk, leaf, fanout, threads=32)
    assert d_chunks.cpu().numpy().tobytes() == b"".join(want_chunks)
    assert d_root.cpu().numpy().tobytes() == want_root
    # compare kernel: all equal, then flip one expected digest
    d_want = d_chunks.clone()
    d_ok = torch.empty(nch, dtype=torch.uint8, device="cuda")
    engine.dev_compare(0, d_chunks.data_ptr(), d_want.data_ptr(), nch, d_ok.data_ptr(), st)
    torch.cuda.synchronize()
    assert d_ok.cpu().tolist() == [1] * nch
    d_want[32 * 5 + 7] ^= 1
    engine.dev_compare(0, d_chunks.data_ptr(), d_want.data_ptr(), nch, d_ok.data_ptr(), st)
    torch.cuda.synchronize()
    assert d_ok.cpu().tolist() == [int(i != 5) for i in range(nch)]


def test_full_size_config2_10GB(engine, oracle):
    """BASELINE config 2: one 10 GB blob, HBM-resident.  Chunk list and root vs the threaded oracle."""
    torch = _torch()
    size = 10_000_000_000
    chunk, leaf, fanout = 8 << 20, 16 << 10, 8
    nch = -(-size // chunk)
    data = torch.empty(size, dtype=torch.uint8, device="cuda")
    d_chunks = torch.empty(nch * 32, dtype=torch.uint8, device="cuda")
    d_root = torch.empty(32, dtype=torch.uint8, device="cuda")
    engine.dev_gen_fill(0, data.data_ptr(), 0, size, SEED)
    engine.dev_tree_digest(0, data.data_ptr(), size, (chunk, leaf, fanout), d_chunks.data_ptr(), d_root.data_ptr())
    torch.cuda.synchronize()
    host = data.cpu().numpy()
    del data
    want_chunks, _, want_root = oracle.tree_digest_ptr(host.ctypes.data, size, chunk, leaf, fanout, threads=64)
    assert nch == 1193 and len(want_chunks) == nch
    assert d_chunks.cpu().numpy().tobytes() == b"".join(want_chunks)
    assert d_root.cpu().numpy().tobytes() == want_root
    # spot-check the generator at a far offset too
    assert host[size - 4096:].tobytes() == oracle.gen(size - 4096, 4096, SEED)


# ------------------------------------------------------------------------------------------------
# re-entrancy and cancellation (the reference calls the path from 3 goroutines; ctx cancel)
# ------------------------------------------------------------------------------------------------
def test_concurrent_callers(engine, oracle):
    blobs = [oracle.gen(0, 3_000_000 + 1000 * i, SEED + 10 + i) for i in range(6)]
    want = [oracle.tree_digest(b, 1 << 20, 16 << 10, 8)[2] for b in blobs]
    got = [None] * len(blobs)
    errs = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = engine.tree_digest(blobs[i], 1 << 20, 16 << 10, 8)[1]
                assert engine.sha256(blobs[i][:1000]) == hashlib.sha256(blobs[i][:1000]).digest()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(blobs))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs and got == want


def test_cancel(oracle):
    """An operation handle is one Go context (push.go:150-159): canceling it fails the calls made through it -- also
    calls that start afterwards -- and nobody else's; a root-level cancel only aborts what is in flight."""
    blob = oracle.gen(0, 20_000_000, SEED)
    want = oracle.tree_digest(blob, 1 << 20, 16 << 10, 8)[2]
    with modelx_b200.Engine(devices=[0], ring_bytes=4 << 20) as eng:
        with eng.op() as op, eng.op() as other:
            op.cancel()
            assert op.is_canceled() and not other.is_canceled()
            with pytest.raises(modelx_b200.MxdError) as ei:
                op.tree_digest(blob, 1 << 20, 16 << 10, 8)
            assert ei.value.status == -6
            with pytest.raises(modelx_b200.MxdError) as ei:
                op.sha256(blob[:1000])
            assert ei.value.status == -6
            assert other.tree_digest(blob, 1 << 20, 16 << 10, 8)[1] == want       # a sibling operation is untouched
            assert eng.sha256(blob[:1000]) == hashlib.sha256(blob[:1000]).digest()  # and so is the root handle
            op.reset_cancel()
            assert op.tree_digest(blob, 1 << 20, 16 << 10, 8)[1] == want
        eng.cancel()                                                               # nothing in flight: a no-op, not sticky
        assert eng.tree_digest(blob, 1 << 20, 16 << 10, 8)[1] == want


# ------------------------------------------------------------------------------------------------
# one process driving several GPUs (mxd_open with a device list): chunk ranges per device
# ------------------------------------------------------------------------------------------------
def test_in_process_multi_gpu_tree_digest(oracle, tmp_path):
    torch = _torch()
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs in one box")
    size = 200_000_000 + 12345
    blob = oracle.gen(0, size, SEED + 20)
    want_chunks, _, want_root = oracle.tree_digest(blob, 8 << 20, 16 << 10, 8)
    p = tmp_path / "blob.bin"
    p.write_bytes(blob)
    with modelx_b200.Engine(devices=list(range(ndev))) as eng:
        assert eng.device_count() == ndev
        chunks, root = eng.tree_digest(blob)                       # pageable host memory, split across devices
        assert (chunks, root) == (want_chunks, want_root)
        chunks, root, sz = eng.tree_digest_file(str(p))            # file, split across devices
        assert (chunks, root, sz) == (want_chunks, want_root, size)
        # whole-file digests are spread round-robin over the devices
        paths = []
        for i in range(2 * ndev):
            q = tmp_path / f"f{i}"
            q.write_bytes(blob[i * 1000:i * 1000 + 3_000_000 + i])
            paths.append(str(q))
        got, _ = eng.sha256_files(paths)
        assert got == [hashlib.sha256(blob[i * 1000:i * 1000 + 3_000_000 + i]).digest() for i in range(2 * ndev)]
        for pth in paths[:ndev + 1]:
            assert eng.sha256_file(pth)[0] == hashlib.sha256(open(pth, "rb").read()).digest()


def test_cancel_from_another_thread_mid_stream(oracle):
    """ctx cancel while a large pageable blob is streaming: the call returns MXD_ERR_CANCELED promptly
    (push.go:156-159 semantics) and the engine is usable again after reset."""
    import numpy as np
    size = 8_000_000_000          # >= 145 ms of streaming even at full PCIe rate; the cancel lands 20 ms in
    blob = np.zeros(size, dtype=np.uint8)
    blob[::4096] = 7
    with modelx_b200.Engine(devices=[0], ring_bytes=64 << 20) as eng:
        result = {}

        def work():
            try:
                eng.tree_digest_ptr(blob.ctypes.data, size)
                result["rc"] = 0
            except modelx_b200.MxdError as e:
                result["rc"] = e.status

        t = threading.Thread(target=work)
        t.start()
        import time
        time.sleep(0.02)
        eng.cancel()
        t.join(timeout=60)
        assert not t.is_alive() and result["rc"] == -6
        eng.reset_cancel()
        small = oracle.gen(0, 5_000_000, SEED)
        assert eng.tree_digest(small)[1] == oracle.tree_digest(small, 8 << 20, 16 << 10, 8)[2]


def test_files_with_skewed_sizes(engine, tmp_path):
    """Many small files, a few empty ones and one large one in the same lock-step batch: finished
    messages must release their share of the ring and the long one must still come out right."""
    rng = random.Random(77)
    sizes = [0, 0, 1, 63, 64, 65, 4096, 100_000, 1_000_000] * 6 + [150_000_000] + [rng.randrange(0, 300_000) for _ in range(60)]
    rng.shuffle(sizes)
    paths, want = [], []
    for i, n in enumerate(sizes):
        p = tmp_path / f"s{i}"
        data = rng.randbytes(n)
        p.write_bytes(data)
        paths.append(str(p))
        want.append(hashlib.sha256(data).digest())
    got, got_sizes = engine.sha256_files(paths)
    assert got == want and got_sizes == sizes
    with modelx_b200.Engine(devices=[0], ring_bytes=4 << 20) as small:       # 1 MiB slots: many rounds
        got2, _ = small.sha256_files(paths)
    assert got2 == want


def test_tree_digest_of_unaligned_device_buffer(engine, oracle):
    """Device-resident blob that starts at an odd address (e.g. a slice of a larger tensor)."""
    torch = _torch()
    rng = np.random.default_rng(9)
    host = rng.integers(0, 256, size=5_000_011, dtype=np.uint8)
    dev = torch.from_numpy(host).cuda()
    for off in (1, 2, 3, 5, 8, 13):
        n = host.size - off - (off % 7)
        chunks, root = engine.tree_digest_ptr(dev.data_ptr() + off, n, 1 << 20, 16 << 10, 8)
        want_chunks, _, want_root = oracle.tree_digest(host[off:off + n].tobytes(), 1 << 20, 16 << 10, 8)
        assert (chunks, root) == (want_chunks, want_root), off


def test_tree_golden_vectors_on_gpu(engine, oracle):
    """The committed hashlib-made vectors (tests/golden/tree_golden.json) through the CUDA path."""
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "tree_golden.json")) as f:
        rows = json.load(f)
    torch = _torch()
    for r in rows:
        # data generated ON the GPU by k_gen_fill at the golden offset, digested in place
        n8 = (r["offset"] + r["size"] + 7) // 8 * 8
        dev = torch.empty(max(n8, 8), dtype=torch.uint8, device="cuda")
        engine.dev_gen_fill(0, dev.data_ptr(), 0, n8, r["seed"])
        torch.cuda.synchronize()
        ptr = dev.data_ptr() + r["offset"]
        assert engine.sha256_ptr(ptr, r["size"]).hex() == r["sha256"]
        chunks, root = engine.tree_digest_ptr(ptr, r["size"], r["chunk"], r["leaf"], r["fanout"])
        assert [c.hex() for c in chunks] == r["chunks"] and root.hex() == r["root"], r["size"]


def test_file_part_digests_follow_calc_parts(engine, tmp_path):
    """calcParts (extension_s3.go:99-112) feeding per-part digests: the multipart split and the hash as one call."""
    rng = random.Random(5)
    data = rng.randbytes(50_000_017)
    p = tmp_path / "blob.bin"
    p.write_bytes(data)
    for nparts in (1, 3, 7):
        parts = modelx_b200.calc_parts(len(data), nparts)
        got = engine.sha256_file_parts(str(p), parts)
        assert got == [hashlib.sha256(data[o:o + l]).digest() for o, l in parts]
    # 8 MiB "multipart chunks" as in BASELINE config 4, plus overlapping / empty ranges
    parts = [(o, min(8 << 20, len(data) - o)) for o in range(0, len(data), 8 << 20)] + [(5, 0), (1, 100), (0, len(data))]
    got = engine.sha256_file_parts(str(p), parts)
    assert got == [hashlib.sha256(data[o:o + l]).digest() for o, l in parts]
    with pytest.raises(modelx_b200.MxdError) as ei:
        engine.sha256_file_parts(str(p), [(len(data) - 10, 11)])
    assert ei.value.status == -4


@pytest.mark.gpu
@pytest.mark.parametrize("pair_max", ["0", "100000"])
def test_small_launch_kernels_agree(tmp_path, pair_max):
    """Launches of at most 4,736 messages run in k_sha256_chains_pair, up to 32,768 in k_sha256_chains_coop.  A fresh
    process with MXD_TUNE_PAIR=0 sends everything small through the cooperative kernel, =100000 everything up to 32,768
    messages through the pair kernel (4 CTAs per SM and more): same digests either way -- ragged batch, files, hasher."""
    import subprocess
    import sys
    script = r'''
import hashlib, os, random, sys
import modelx_b200
rng = random.Random(11)
eng = modelx_b200.Engine(devices=[0])
msgs = [rng.randbytes(n) for n in list(range(0, 200)) + [1000, 4095, 4096, 4097, 65536, 1_000_003]]
assert eng.sha256_batch(msgs) == [hashlib.sha256(m).digest() for m in msgs]
many = [rng.randbytes(rng.randrange(0, 700)) for _ in range(9000)]            # > 4,736 messages in one launch
assert eng.sha256_batch(many) == [hashlib.sha256(m).digest() for m in many]
paths = []
for i, n in enumerate([0, 1, 63, 64, 5_000_000, 70_000_001, 33]):
    p = os.path.join(sys.argv[1], f"f{i}"); d = rng.randbytes(n); open(p, "wb").write(d); paths.append((p, hashlib.sha256(d).digest(), n))
digs, sizes = eng.sha256_files([p for p, _, _ in paths])
assert digs == [d for _, d, _ in paths] and sizes == [n for _, _, n in paths]
h = eng.hasher(); ref = hashlib.sha256()
for n in (0, 1, 63, 64, 65, 100_000, 4 << 20, 3):
    b = rng.randbytes(n); h.write(b); ref.update(b)
assert h.sum() == ref.digest()
print("pair-ok", eng.stats()["kernel_launches"])
'''
    env = dict(os.environ, MXD_TUNE_PAIR=pair_max)
    out = subprocess.run([sys.executable, "-c", script, str(tmp_path)], capture_output=True, text=True, timeout=600, env=env,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0 and "pair-ok" in out.stdout, out.stdout + out.stderr
