"""Generates tests/golden/tree_golden.json: modelx.tree.v1 vectors evaluated with hashlib only (no oracle,
no GPU), over blobs from the splitmix64 counter stream (re-implemented here in pure Python).
Run: python tests/golden/make_tree_golden.py"""
import hashlib
import json
import os
import struct

M = (1 << 64) - 1


def splitmix_bytes(offset, n, seed):
    out = bytearray()
    j0, j1 = offset // 8, (offset + n + 7) // 8
    for j in range(j0, j1):
        z = (seed + (j + 1) * 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        out += struct.pack("<Q", z ^ (z >> 31))
    s = offset - 8 * j0
    return bytes(out[s:s + n])


def tree(data, chunk, leaf, fanout):
    size = len(data)
    k, span = 0, leaf
    while span < chunk:
        span *= fanout
        k += 1
    assert span == chunk and k >= 1
    level = [hashlib.sha256(data[i:i + leaf]).digest() for i in range(0, max(size, 1), leaf)]
    chunks, lv = None, 0
    while lv < k or len(level) > 1:
        level = [hashlib.sha256(b"".join(level[i:i + fanout])).digest() for i in range(0, len(level), fanout)]
        lv += 1
        if lv == k:
            chunks = level
    root = hashlib.sha256(b"modelx.tree.v1\0\0" + struct.pack("<QQII", size, leaf, fanout, 0) + level[0]).digest()
    return chunks, root


CASES = [  # (seed, offset, size, chunk, leaf, fanout)
    (0x6D6F64656C78, 0, 0, 8 << 20, 16 << 10, 8), (0x6D6F64656C78, 0, 1, 8 << 20, 16 << 10, 8),
    (0x6D6F64656C78, 3, 16384, 8 << 20, 16 << 10, 8), (0x6D6F64656C78, 0, 16385, 8 << 20, 16 << 10, 8),
    (1, 0, 200_000, 1 << 16, 1 << 10, 8), (2, 5, 1_000_003, 1 << 18, 1 << 12, 8), (3, 0, 300_000, 4096, 64, 2),
    (4, 0, 2_500_000, 1 << 20, 16 << 10, 8), (5, 7, 1 << 20, 1 << 20, 16 << 10, 64),
]

if __name__ == "__main__":
    rows = []
    for seed, off, size, chunk, leaf, fanout in CASES:
        data = splitmix_bytes(off, size, seed)
        chunks, root = tree(data, chunk, leaf, fanout)
        rows.append({"seed": seed, "offset": off, "size": size, "chunk": chunk, "leaf": leaf, "fanout": fanout,
                     "sha256": hashlib.sha256(data).hexdigest(), "chunks": [c.hex() for c in chunks], "root": root.hex()})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tree_golden.json"), "w") as f:
        json.dump(rows, f, indent=1)
    print("wrote", len(rows), "vectors")
