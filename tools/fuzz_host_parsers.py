"""Developer tool: mutation fuzzing of the host-side parsers that see untrusted text -- the manifest JSON (mxc_pull_check,
mxc_fs_put_manifest), modelx.yaml (mxc_parse_manifest), digest strings and tar.gz archives (mxc_untgz) -- against the
CPU test double built with AddressSanitizer + UBSan.  Run through tools/fuzz_host_parsers.sh (preloads the runtimes).
A crash / sanitizer report is the finding; return codes are not checked (most inputs are rejected, as they should be)."""
import ctypes as C, gzip, io, json, os, random, sys, tarfile, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import mock_build
import modelx_b200

lib_path = mock_build.build(os.environ.get("MXD_MOCK_SANITIZE", ""))
eng = modelx_b200.Engine(devices=[0], lib_path=lib_path)
lib = eng._lib
rng = random.Random(int(os.environ.get("FUZZ_SEED", "1")))
iters = int(os.environ.get("FUZZ_ITERS", "3000"))
work = tempfile.mkdtemp(prefix="mxfuzz")
base = os.path.join(work, "model"); os.makedirs(base)
open(os.path.join(base, "a.bin"), "wb").write(b"x" * 1000)
open(os.path.join(base, "modelx.yaml"), "w").write("description: d\nframework: f\ntags:\n- a\nmaintainers:\n- m\nannotations:\n  k: v\n")
good = json.dumps({"schemaVersion": 0, "mediaType": "application/vnd.modelx.model.manifest.v1+json",
                   "config": {"name": "modelx.yaml", "digest": "sha256:" + "0" * 64, "size": 3, "mode": 420, "modified": "2024-01-01T00:00:00Z"},
                   "blobs": [{"name": "a.bin", "digest": "sha256:" + "1" * 64, "size": 1000, "annotations": {"k": "v"}, "urls": ["u"]},
                             {"name": "d", "mediaType": "application/vnd.modelx.model.directory.v1.tar+gz", "digest": "sha256:" + "2" * 64}],
                   "annotations": {"é": "😀"}})
tokens = ['"', "\\", "{", "}", "[", "]", ",", ":", "null", "true", "1e999", "-", "\\u", "\\ud800", "\x00", "../", "/", "sha256:", "9" * 40, " ", "\n"]

def mutate(s: str) -> bytes:
    b = bytearray(s.encode("utf-8", "surrogatepass"))
    for _ in range(rng.randrange(1, 6)):
        k = rng.randrange(5)
        pos = rng.randrange(len(b) + 1)
        if k == 0 and b: del b[pos % len(b): pos % len(b) + rng.randrange(1, 20)]
        elif k == 1: b[pos:pos] = rng.choice(tokens).encode()
        elif k == 2 and b: b[pos % len(b)] = rng.randrange(1, 256)
        elif k == 3 and b: a = pos % len(b); b[a:a] = b[a: a + rng.randrange(1, 60)] * rng.randrange(1, 4)
        else: b[pos:pos] = bytes(rng.randrange(1, 256) for _ in range(rng.randrange(1, 8)))
    return bytes(b).replace(b"\x00", b"\x01")

out = C.c_void_p()
def take():
    if out.value: lib.mxc_free(out); out.value = None

n_ok = 0
for i in range(iters):
    m = mutate(good)
    rc = lib.mxc_pull_check(eng.handle, base.encode(), m, C.byref(out)); take(); n_ok += rc == 0
    lib.mxc_fs_put_manifest(os.path.join(work, "reg").encode(), b"library/m", b"v" + str(i % 7).encode(), b"application/json", m)
    y = mutate(open(os.path.join(base, "modelx.yaml")).read())
    open(os.path.join(base, "fz.yaml"), "wb").write(y)
    lib.mxc_parse_manifest(base.encode(), b"fz.yaml", C.byref(out)); take()
    d = mutate("sha256:" + "ab" * 32)
    buf = (C.c_uint8 * 32)(); lib.mxd_digest_parse(d, buf)
    lib.mxc_blob_digest_path(mutate("library/m"), d, C.byref(out)); take()
    if i % 5 == 0:                                    # the digest cache file and a stored manifest / index read back by pull
        os.makedirs(os.path.join(base, ".modelx"), exist_ok=True)
        cache = json.dumps({"a.bin": {"size": 1000, "mtime_ns": os.stat(os.path.join(base, "a.bin")).st_mtime_ns, "digest": "sha256:" + "3" * 64}})
        open(os.path.join(base, ".modelx", "digests.json"), "wb").write(mutate(cache))
        lib.mxc_push_digest(eng.handle, base.encode(), b"modelx.yaml", 2, C.byref(out)); take()
        reg = os.path.join(work, "reg2"); mdir = os.path.join(reg, "library", "m", "manifests"); os.makedirs(mdir, exist_ok=True)
        open(os.path.join(mdir, "v1"), "wb").write(m)
        open(os.path.join(reg, "library", "m", "index.json"), "wb").write(mutate(good))
        lib.mxc_fs_get_manifest(reg.encode(), b"library/m", b"v1", C.byref(out)); take()
        lib.mxc_fs_get_index(reg.encode(), b"library/m", C.byref(out)); take()
        lib.mxc_fs_get_index(reg.encode(), b"", C.byref(out)); take()
        lib.mxc_pull_local(eng.handle, reg.encode(), b"library/m", b"v1", os.path.join(work, "into").encode(), C.byref(out)); take()
    if i % 10 == 0:                                   # archives: a valid tar.gz with a few entries, then byte damage inside the tar stream
        raw = io.BytesIO()
        with tarfile.open(fileobj=raw, mode="w") as tf:
            for name in ("f.txt", "sub/g.bin", "l" * 120 + ".x"):
                ti = tarfile.TarInfo(name); ti.size = 300; tf.addfile(ti, io.BytesIO(os.urandom(300)))
            ln = tarfile.TarInfo("lnk"); ln.type = tarfile.SYMTYPE; ln.linkname = "f.txt"; tf.addfile(ln)
        t = bytearray(raw.getvalue())
        for _ in range(rng.randrange(1, 8)): t[rng.randrange(len(t))] = rng.randrange(256)
        ap = os.path.join(work, "fz.tar.gz")
        with gzip.open(ap, "wb") as f: f.write(bytes(t[: rng.randrange(len(t) // 2, len(t) + 1)]))
        lib.mxc_untgz(ap.encode(), os.path.join(work, "x%d" % (i % 5)).encode())
eng.close()
print(f"fuzz done: {iters} iterations, pull_check accepted {n_ok}")
