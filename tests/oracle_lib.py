"""ctypes loader for oracle/liboracle.so -- the CPU checker.  TEST INFRASTRUCTURE ONLY: imported
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
modelx_b200/."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Tuple

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(_ROOT, "oracle", "liboracle.so")


class _Part(C.Structure):
    _fields_ = [("offset", C.c_int64), ("length", C.c_int64)]


class _Span(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("len", C.c_uint64)]


class _Ctx(C.Structure):
    _fields_ = [("h", C.c_uint32 * 8), ("nbytes", C.c_uint64), ("nbuf", C.c_uint32), ("buf", C.c_uint8 * 64)]


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_LIB):
            import subprocess
            subprocess.run(["make", "-C", os.path.dirname(ORACLE_LIB)], check=True)
        L = C.CDLL(ORACLE_LIB)
        u8p = C.POINTER(C.c_uint8)
        L.orc_sha256.argtypes = [C.c_void_p, C.c_size_t, u8p]
        L.orc_sha256_init.argtypes = [C.POINTER(_Ctx)]
        L.orc_sha256_update.argtypes = [C.POINTER(_Ctx), C.c_void_p, C.c_size_t]
        L.orc_sha256_final.argtypes = [C.POINTER(_Ctx), u8p]
        L.orc_sha256_set_engine.argtypes = [C.c_int]
        L.orc_client_digest.argtypes = [C.c_char_p, u8p, C.POINTER(C.c_uint64)]
        L.orc_digest_string.argtypes = [u8p, C.c_char_p]
        L.orc_pull_file_matches.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_calc_parts.argtypes = [C.c_int64, C.c_int64, C.POINTER(_Part)]
        L.orc_server_part_count.argtypes = [C.c_int64, C.c_int]
        L.orc_server_part_count.restype = C.c_int64
        L.orc_tree_shape.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
        L.orc_tree_digest.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, u8p,
                                      C.POINTER(C.c_uint64), u8p, u8p]
        L.orc_tree_root.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, u8p, u8p]
        L.orc_hash_segments.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, u8p]
        L.orc_sha256_batch.argtypes = [C.POINTER(_Span), C.c_uint64, C.c_int, u8p]
        L.orc_gen_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        self.L = L

    # --- SHA-256 -----------------------------------------------------------------------------
    def set_engine(self, engine: int) -> int:
        return self.L.orc_sha256_set_engine(engine)

    def engine(self) -> int:
        return self.L.orc_sha256_engine()

    def sha256(self, data: bytes) -> bytes:
        out = (C.c_uint8 * 32)()
        buf = C.create_string_buffer(bytes(data), max(len(data), 1))
        self.L.orc_sha256(buf, len(data), out)
        return bytes(out)

    def sha256_ptr(self, ptr: int, n: int) -> bytes:
        out = (C.c_uint8 * 32)()
        self.L.orc_sha256(ptr, n, out)
        return bytes(out)

    def sha256_incremental(self, pieces) -> bytes:
        ctx = _Ctx()
        self.L.orc_sha256_init(C.byref(ctx))
        for p in pieces:
            buf = C.create_string_buffer(bytes(p), max(len(p), 1))
            self.L.orc_sha256_update(C.byref(ctx), buf, len(p))
        out = (C.c_uint8 * 32)()
        self.L.orc_sha256_final(C.byref(ctx), out)
        return bytes(out)

    # --- reference call sites ----------------------------------------------------------------
    def client_digest(self, path: str) -> Tuple[bytes, int]:
        out = (C.c_uint8 * 32)()
        size = C.c_uint64()
        rc = self.L.orc_client_digest(path.encode(), out, C.byref(size))
        if rc < 0:
            raise OSError(-rc, os.strerror(-rc), path)
        return bytes(out), size.value

    def digest_string(self, d: bytes) -> str:
        out = C.create_string_buffer(72)
        self.L.orc_digest_string((C.c_uint8 * 32).from_buffer_copy(d), out)
        return out.value.decode()

    def pull_file_matches(self, path: str, want: str) -> int:
        return self.L.orc_pull_file_matches(path.encode(), want.encode())

    def calc_parts(self, total: int, n: int) -> List[Tuple[int, int]]:
        parts = (_Part * max(n, 1))()
        if self.L.orc_calc_parts(total, n, parts) != 0:
            raise ZeroDivisionError("calcParts")
        return [(p.offset, p.length) for p in parts[:n]]

    def server_part_count(self, size: int, force: bool = False) -> int:
        return int(self.L.orc_server_part_count(size, 1 if force else 0))

    # --- tree (modelx-b200's own definition, restated on the CPU) ---------------------------------
    def tree_shape(self, size: int, chunk: int, leaf: int, fanout: int) -> List[int]:
        counts = (C.c_uint64 * 80)()
        lv = self.L.orc_tree_shape(size, leaf, fanout, chunk, counts, 80)
        if lv < 0:
            raise ValueError("bad tree parameters")
        return [int(counts[i]) for i in range(lv)]

    def tree_digest(self, data, chunk: int, leaf: int, fanout: int, threads: int = 8):
        """-> (chunk_digests list, top, root)"""
        if isinstance(data, (bytes, bytearray)):
            n = len(data)
            buf = C.create_string_buffer(bytes(data), max(n, 1))
            ptr = C.addressof(buf)
        else:  # numpy
            n = data.nbytes
            ptr = data.ctypes.data
        return self.tree_digest_ptr(ptr, n, chunk, leaf, fanout, threads)

    def tree_digest_ptr(self, ptr: int, n: int, chunk: int, leaf: int, fanout: int, threads: int = 8):
        nch = max(1, -(-n // chunk))
        chunks = (C.c_uint8 * (32 * nch))()
        got = C.c_uint64()
        top = (C.c_uint8 * 32)()
        root = (C.c_uint8 * 32)()
        rc = self.L.orc_tree_digest(ptr, n, leaf, fanout, chunk, threads, chunks, C.byref(got), top, root)
        if rc != 0:
            raise ValueError("bad tree parameters")
        raw = bytes(chunks)
        return [raw[32 * i:32 * i + 32] for i in range(got.value)], bytes(top), bytes(root)

    def hash_segments_ptr(self, ptr: int, n: int, seg: int, threads: int = 8) -> bytes:
        nseg = max(1, -(-n // seg))
        out = (C.c_uint8 * (32 * nseg))()
        self.L.orc_hash_segments(ptr, n, seg, threads, out)
        return bytes(out)

    def sha256_batch_ptrs(self, spans, threads: int = 8) -> List[bytes]:
        n = len(spans)
        arr = (_Span * max(n, 1))()
        for i, (p, ln) in enumerate(spans):
            arr[i].ptr = p
            arr[i].len = ln
        out = (C.c_uint8 * (32 * max(n, 1)))()
        self.L.orc_sha256_batch(arr, n, threads, out)
        raw = bytes(out)
        return [raw[32 * i:32 * i + 32] for i in range(n)]

    # --- synthetic data ----------------------------------------------------------------------------
    def gen(self, offset: int, n: int, seed: int) -> bytes:
        buf = C.create_string_buffer(max(n, 1))
        self.L.orc_gen_fill(buf, offset, n, seed)
        return buf.raw[:n]

    def gen_into(self, ptr: int, offset: int, n: int, seed: int) -> None:
        self.L.orc_gen_fill(ptr, offset, n, seed)
