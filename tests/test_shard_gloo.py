"""world_size-2/3 gloo runs on CPU of the N>1 path's host logic: chunk-range sharding, the fixed
size all-gather of chunk digests, and finishing the tree from the gathered list.  The per-rank
hashing is done by the CPU oracle here (no GPU in this test); the GPU version of the same flow is
tests/test_gpu_parity.py::test_tree_sharded_equals_whole and bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from modelx_b200 import shard

SEED = 0x6D6F64656C78


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, size, chunk, leaf, fanout, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.oracle_lib import Oracle
        orc = Oracle()
        nchunks = shard.chunk_count(size, chunk)
        per = shard.chunks_per_rank(nchunks, world)
        b0, b1 = shard.byte_range(rank, world, size, chunk)
        c0, c1 = shard.chunk_range(rank, world, nchunks)
        local = torch.zeros(per * 32, dtype=torch.uint8)
        if c1 > c0:
            piece = orc.gen(b0, b1 - b0, SEED)             # rank-local bytes of the one logical blob
            mine, _, _ = orc.tree_digest(piece, chunk, leaf, fanout, threads=2)
            assert len(mine) == c1 - c0
            local[:32 * (c1 - c0)] = torch.frombuffer(bytearray(b"".join(mine)), dtype=torch.uint8)
        allc = shard.gather_chunk_digests(local, world, nchunks)
        q.put((rank, bytes(allc.numpy().tobytes())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,size", [(2, 9 * (1 << 16) + 123), (3, 4 * (1 << 16)), (2, 1), (3, 7 * (1 << 16) + 1)])
def test_sharded_chunk_list_equals_whole(world, size):
    chunk, leaf, fanout = 1 << 16, 1 << 10, 8
    from tests.oracle_lib import Oracle
    orc = Oracle()
    blob = orc.gen(0, size, SEED)
    want_chunks, _, want_root = orc.tree_digest(blob, chunk, leaf, fanout)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, size, chunk, leaf, fanout, q)) for r in range(world)]
    [p.start() for p in procs]
    results = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, allc in results:
        assert allc == b"".join(want_chunks), f"rank {rank} assembled a different chunk list"
    # finishing from the gathered list (what every rank does after the all-gather)
    import hashlib
    import struct
    level = list(want_chunks)
    while len(level) > 1:
        level = [hashlib.sha256(b"".join(level[i:i + fanout])).digest() for i in range(0, len(level), fanout)]
    root = hashlib.sha256(b"modelx.tree.v1\0\0" + struct.pack("<QQII", size, leaf, fanout, 0) + level[0]).digest()
    assert root == want_root


def test_ranges_partition_the_chunk_space():
    for nchunks in (1, 2, 7, 8, 9, 1193, 11921, 16690):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard.chunk_range(r, world, nchunks) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nchunks
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            per = shard.chunks_per_rank(nchunks, world)
            # only the tail may be short: the all-gather prefix is exactly the chunk list
            assert all(c1 - c0 == per for c0, c1 in ranges if c1 < nchunks)
    assert shard.byte_range(7, 8, 100_000_000_000, 8 << 20) == (10437 * (8 << 20), 100_000_000_000)
