"""The host mirror of pkg/client / pkg/registry end to end: the digest phase of Push, the check phase of Pull,
BASELINE config 1 (one 64 MB random blob: digest + PutBlob into the in-process FS registry), the read-once push with
per-part digests, verified pull, index.json and directory blobs.

Every test runs twice (fixture ``any_engine``): against the CPU test double of the library in the CPU-only container
(host logic: `-m "not gpu"`) and against the CUDA build on the B200 (`-m gpu`, the parity run proper)."""
import hashlib
import json
import os
import stat
import threading
import time

import pytest

import modelx_b200
from modelx_b200 import _native as N
from modelx_b200 import client

EMPTY = "sha256:e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"


def _go_time(ns: int) -> str:
    """time.Time.MarshalJSON of a file's ModTime in the local zone (RFC3339Nano)."""
    sec, frac = divmod(ns, 10 ** 9)
    lt = time.localtime(sec)
    s = time.strftime("%Y-%m-%dT%H:%M:%S", lt)
    if frac:
        s += "." + ("%09d" % frac).rstrip("0")
    off = lt.tm_gmtoff
    if off == 0:
        return s + "Z"
    a = abs(off)
    return s + ("+" if off > 0 else "-") + "%02d:%02d" % (a // 3600, (a % 3600) // 60)


def _model(tmp_path, big=64_000_000):
    d = tmp_path / "model"
    d.mkdir()
    files = {
        "modelx.yaml": b"framework: pytorch\nmodelFiles: []\n",
        "model-00001.safetensors": os.urandom(big),
        "tokenizer.json": b'{"a": 1}' * 1000,
        "empty.txt": b"",
        "README.md": b"# model\n",
    }
    for n, b in files.items():
        (d / n).write_bytes(b)
    os.chmod(d / "tokenizer.json", 0o600)
    (d / ".cache").write_bytes(b"skip me")
    return d, files


def test_push_digest_phase_matches_reference_semantics(any_engine, tmp_path):
    d, files = _model(tmp_path, big=5_000_000)
    cl = client.Client(any_engine)
    m = json.loads(cl.push_digest_json(str(d)))
    assert [b["name"] for b in m["blobs"]] == sorted(n for n in files if n != "modelx.yaml")
    for desc in m["blobs"] + [m["config"]]:
        data = files[desc["name"]]
        st = os.stat(d / desc["name"])
        assert desc["digest"] == "sha256:" + hashlib.sha256(data).hexdigest()          # push.go:160
        assert desc.get("size", 0) == len(data)                                         # omitempty: 0 is dropped
        assert desc["mode"] == stat.S_IMODE(st.st_mode)                                 # regular file: permission bits only
        assert desc["modified"] == _go_time(st.st_mtime_ns)
    assert m["config"]["mediaType"] == "application/vnd.modelx.model.config.v1.yaml"
    assert [b for b in m["blobs"] if b["name"] == "empty.txt"][0]["digest"] == EMPTY


def test_push_digest_with_tree_annotation(any_engine, oracle, tmp_path):
    d, files = _model(tmp_path, big=20_000_000)
    m = json.loads(client.Client(any_engine).push_digest_json(str(d), with_tree=True))
    blob = [b for b in m["blobs"] if b["name"].endswith(".safetensors")][0]
    _, _, root = oracle.tree_digest(files[blob["name"]], 8 << 20, 16 << 10, 8)
    assert blob["annotations"]["modelx.tree.v1"] == \
        f"{modelx_b200.digest_string(root)};leaf=16384;fanout=8;chunk=8388608;chunks=3"
    assert blob["digest"] == "sha256:" + hashlib.sha256(files[blob["name"]]).hexdigest()   # wire-compatible identity kept


def test_directory_blobs_are_packed_pushed_and_pulled(any_engine, tmp_path):
    """pushDirectory / pullDirectory (push.go:102-118, pull.go:146-208): a sub-directory becomes a tar.gz blob under
    <dir>/.modelx/<name>.tar.gz whose digest is taken while the archive is written (helper.go:46-50)."""
    d, files = _model(tmp_path, big=1000)
    (d / "tokenizer").mkdir()
    (d / "tokenizer" / "vocab.txt").write_bytes(b"a\nb\n" * 5000)
    (d / "tokenizer" / "merges").mkdir()
    (d / "tokenizer" / "merges" / ("m" * 130 + ".bin")).write_bytes(os.urandom(70_000))     # name > 100 bytes: long-name record
    os.chmod(d / "tokenizer" / "vocab.txt", 0o640)
    cl = client.Client(any_engine)
    m = json.loads(cl.push_digest_json(str(d)))
    desc = [b for b in m["blobs"] if b["name"] == "tokenizer"][0]
    archive = d / ".modelx" / "tokenizer.tar.gz"
    assert desc["mediaType"] == "application/vnd.modelx.model.directory.v1.tar+gz"
    assert desc["digest"] == "sha256:" + hashlib.sha256(archive.read_bytes()).hexdigest()   # digest of the archive bytes
    assert desc["size"] == archive.stat().st_size and desc["mode"] & (1 << 31)             # os.ModeDir, push.go:107
    import tarfile
    with tarfile.open(archive) as tf:                                                        # any tar reader extracts it
        names = sorted(tf.getnames())
        assert names == ["merges", "merges/" + "m" * 130 + ".bin", "vocab.txt"]
        assert tf.getmember("vocab.txt").mode == 0o640 and tf.getmember("vocab.txt").mtime == 0   # ClearAttributes
    assert cl.tgz(str(d / "tokenizer"))[0] == desc["digest"]                                # deterministic: TGZ(dir, "") again
    reg = client.LocalRegistry(str(tmp_path / "reg"), any_engine)
    rep = cl.push(reg, "library/m", "v1", str(d))
    assert {b["name"]: b["status"] for b in rep["blobs"]}["tokenizer"] == "done"
    into = tmp_path / "pulled"
    res = {r["name"]: r["status"] for r in cl.pull(reg, "library/m", "v1", str(into))}
    assert res["tokenizer"] == "done"
    assert (into / "tokenizer" / "vocab.txt").read_bytes() == b"a\nb\n" * 5000
    assert stat.S_IMODE(os.stat(into / "tokenizer" / "vocab.txt").st_mode) == 0o640
    assert (into / "tokenizer" / "merges" / ("m" * 130 + ".bin")).stat().st_size == 70_000
    res2 = {r["name"]: r["status"] for r in cl.pull(reg, "library/m", "v1", str(into))}
    assert res2["tokenizer"] == "already exists"                                            # pull.go:149-156: re-archive and compare
    # an archive that tries to escape the target directory is refused
    evil = tmp_path / "evil.tar.gz"
    import io
    with tarfile.open(evil, "w:gz") as tf:
        ti = tarfile.TarInfo("../escape.txt"); ti.size = 1
        tf.addfile(ti, io.BytesIO(b"x"))
    with pytest.raises(modelx_b200.MxdError):
        client.untgz(str(evil), str(tmp_path / "x"), lib=any_engine._lib)
    assert not (tmp_path / "escape.txt").exists()
    evil2 = tmp_path / "evil2.tar.gz"                  # a symlink out of the tree, then a file written "through" it
    with tarfile.open(evil2, "w:gz") as tf:
        ln = tarfile.TarInfo("out"); ln.type = tarfile.SYMTYPE; ln.linkname = "../../outside"
        tf.addfile(ln)
        ti = tarfile.TarInfo("out/pwned.txt"); ti.size = 1
        tf.addfile(ti, io.BytesIO(b"x"))
    (tmp_path / "outside").mkdir()
    with pytest.raises(modelx_b200.MxdError):
        client.untgz(str(evil2), str(tmp_path / "y" / "z"), lib=any_engine._lib)
    assert not (tmp_path / "outside" / "pwned.txt").exists()
    # a header that announces a multi-gigabyte extended-header body (or a huge device/fifo entry) must not be allocated
    import gzip
    evil3 = tmp_path / "evil3.tar.gz"
    hdr = bytearray(512)
    hdr[0:9] = b"PaxHeader"; hdr[100:107] = b"0000644"; hdr[124:135] = b"77777777777"; hdr[156:157] = b"x"
    hdr[257:263] = b"ustar\0"; hdr[148:156] = b" " * 8
    hdr[148:155] = b"%06o\0" % sum(hdr)
    with gzip.open(evil3, "wb") as f:
        f.write(bytes(hdr) + b"\0" * 1024)
    with pytest.raises(modelx_b200.MxdError):
        client.untgz(str(evil3), str(tmp_path / "w"), lib=any_engine._lib)


def test_config1_push_then_pull_through_local_registry(any_engine, tmp_path):
    """BASELINE config 1: 64 MB random blob, pkg/client digest + PutBlob against the in-process registry."""
    d, files = _model(tmp_path)
    reg = client.LocalRegistry(str(tmp_path / "data" / "registry"), any_engine)
    cl = client.Client(any_engine)
    r0 = any_engine.stats()["src_bytes_read"]
    rep = cl.push(reg, "library/llama", "v1", str(d))
    total = sum(len(v) for v in files.values())
    assert any_engine.stats()["src_bytes_read"] - r0 == total       # READ ONCE: bytes read from disk == blob bytes
    status = {b["name"]: b["status"] for b in rep["blobs"]}
    assert status == {"README.md": "done", "empty.txt": "empty", "model-00001.safetensors": "done",
                      "tokenizer.json": "done", "modelx.yaml": "done"}
    base = tmp_path / "data" / "registry" / "library" / "llama"
    for b in rep["blobs"]:
        hexd = b["digest"].split(":")[1]
        assert b["digest"] == "sha256:" + hashlib.sha256(files[b["name"]]).hexdigest()
        if b["status"] == "empty":
            assert not (base / "blobs" / "sha256" / hexd).exists()          # push.go:165-168: never uploaded
            continue
        stored = (base / "blobs" / "sha256" / hexd).read_bytes()
        assert stored == files[b["name"]]                                   # written by the tee, not by a second read
        meta = json.loads((base / "blobs" / "sha256" / (hexd + ".meta")).read_text())
        assert meta == {"contentType": "application/octet-stream", "contentLength": len(stored)}
    assert not [p for p in os.listdir(base / "blobs" / "sha256") if p.startswith(".incoming")]
    assert (base / "manifests" / "v1").read_text() == rep["manifest_json"] == reg.get_manifest_json("library/llama", "v1")
    # second push: content addressed dedupe (HeadBlob, push.go:169-177)
    rep2 = cl.push(reg, "library/llama", "v2", str(d))
    assert {b["status"] for b in rep2["blobs"]} == {"exists", "empty"}

    # index.json after PutManifest (store_fs.go:87-104 -> RefreshIndex :185-238 -> RefreshGlobalIndex :287-330)
    idx = reg.get_index("library/llama")
    assert [v["name"] for v in idx["manifests"]] == ["v1", "v2"] and idx["schemaVersion"] == 0
    assert idx["manifests"][0]["size"] == total
    assert idx["manifests"][0]["modified"] == _go_time(os.stat(base / "manifests" / "v1").st_mtime_ns)
    assert json.loads((base / "index.json.meta").read_text())["contentType"] == "application/vnd.modelx.model.index.v1.json"
    glob = reg.get_index()
    assert [(r["name"], r["mediaType"]) for r in glob["manifests"]] == [("library/llama", "application/vnd.modelx.model.index.v1.json")]

    # pull into a fresh directory, then again (pull.go:115-123 "already exists")
    into = tmp_path / "pulled"
    res = {r["name"]: r["status"] for r in cl.pull(reg, "library/llama", "v1", str(into))}
    assert res == {"README.md": "done", "empty.txt": "empty", "model-00001.safetensors": "done",
                   "tokenizer.json": "done", "modelx.yaml": "done"}
    for n, data in files.items():
        assert (into / n).read_bytes() == data
    assert stat.S_IMODE(os.stat(into / "tokenizer.json").st_mode) == 0o600    # desc.Mode.Perm(), pull.go:129
    res2 = {r["name"]: r["status"] for r in cl.pull(reg, "library/llama", "v1", str(into))}
    assert set(res2.values()) == {"already exists"}
    # corrupt one local file: the check must notice and re-download
    (into / "README.md").write_bytes(b"tampered")
    chk = {r["name"]: r["state"] for r in cl.pull_check(str(into), rep["manifest_json"])}
    assert chk["README.md"] == "differs" and chk["tokenizer.json"] == "already exists"
    os.unlink(into / "tokenizer.json")
    assert {r["name"]: r["state"] for r in cl.pull_check(str(into), rep["manifest_json"])}["tokenizer.json"] == "missing"
    res3 = {r["name"]: r["status"] for r in cl.pull(reg, "library/llama", "v1", str(into))}
    assert res3["README.md"] == "done" and res3["tokenizer.json"] == "done" and (into / "README.md").read_bytes() == files["README.md"]


def test_pull_rejects_a_corrupted_store_blob(any_engine, tmp_path):
    """New: what pull copies is hashed while it is written; the reference writes it unverified
    (pull.go:137-142).  A flipped bit in the store yields DIGEST_INVALID, no file and no leftovers; good blobs land."""
    d, files = _model(tmp_path, big=9_000_000)
    reg = client.LocalRegistry(str(tmp_path / "reg"), any_engine)
    cl = client.Client(any_engine)
    rep = cl.push(reg, "library/m", "v1", str(d))
    dg = {b["name"]: b["digest"] for b in rep["blobs"]}["model-00001.safetensors"]
    stored = tmp_path / "reg" / "library" / "m" / "blobs" / "sha256" / dg.split(":")[1]
    raw = bytearray(stored.read_bytes()); raw[4_000_000] ^= 0x10; stored.write_bytes(raw)
    into = tmp_path / "pulled"
    with pytest.raises(modelx_b200.MxdError) as ei:
        cl.pull(reg, "library/m", "v1", str(into))
    assert ei.value.status == N.MXC_ERR_DIGEST_INVALID and "model-00001.safetensors" in ei.value.detail
    assert not (into / "model-00001.safetensors").exists()
    assert (into / "tokenizer.json").read_bytes() == files["tokenizer.json"]
    assert not [p for p in os.listdir(into) if p.endswith(".modelx-partial")]


def test_manifest_strings_are_not_trusted(any_engine, tmp_path):
    """a manifest digest like sha256:../../etc/shadow or a name like ../x must never reach a path join."""
    reg = client.LocalRegistry(str(tmp_path / "reg"), any_engine)
    cl = client.Client(any_engine)
    for evil in ("sha256:../../../../etc/passwd", "sha256:" + "g" * 64, "md5:" + "0" * 64, "sha256:" + "A" * 64):
        with pytest.raises(modelx_b200.MxdError) as ei:
            reg.exists_blob("library/m", evil)
        assert ei.value.status == N.MXC_ERR_DIGEST_INVALID
    good = {"name": "a.bin", "digest": EMPTY, "modified": "0001-01-01T00:00:00Z"}
    for bad_desc in ({**good, "name": "../x"}, {**good, "name": "a/b"}, {**good, "name": ".."}, {**good, "digest": "sha256:../../x"}):
        man = json.dumps({"schemaVersion": 0, "config": {"name": "modelx.yaml", "modified": "0001-01-01T00:00:00Z"}, "blobs": [bad_desc]})
        with pytest.raises(modelx_b200.MxdError) as ei:
            reg.put_manifest("library/m", "v1", man)
        assert ei.value.status == N.MXC_ERR_MANIFEST
        with pytest.raises(modelx_b200.MxdError) as ei:
            cl.pull_check(str(tmp_path), man)
        assert ei.value.status == N.MXC_ERR_MANIFEST
    for bad_repo in ("../m", "/abs/m", "a//b", ""):
        with pytest.raises(modelx_b200.MxdError):
            reg.put_manifest(bad_repo, "v1", json.dumps({"schemaVersion": 0, "config": good, "blobs": []}))
    with pytest.raises(modelx_b200.MxdError):
        reg.put_manifest("library/m", "../v1", json.dumps({"schemaVersion": 0, "config": good, "blobs": []}))


def test_put_blob_verify_is_a_tee_and_never_destroys_a_good_blob(any_engine, tmp_path):
    """SURVEY 8f.2 + ADVICE r1: the body is hashed while it is written to a temp file (one read of the source); a
    mismatch is DIGEST_INVALID, nothing becomes visible and an existing blob under that key survives."""
    reg = client.LocalRegistry(str(tmp_path / "reg"), any_engine)
    src = tmp_path / "b.bin"
    src.write_bytes(os.urandom(3_000_000))
    good = "sha256:" + hashlib.sha256(src.read_bytes()).hexdigest()
    bad = "sha256:" + "00" * 32
    r0 = any_engine.stats()["src_bytes_read"]
    with pytest.raises(modelx_b200.MxdError) as ei:
        reg.put_blob("library/m", bad, str(src), verify=True)
    assert ei.value.status == N.MXC_ERR_DIGEST_INVALID and "digest invalid" in ei.value.detail
    assert any_engine.stats()["src_bytes_read"] - r0 == 3_000_000          # hashed and written in one read
    assert not reg.exists_blob("library/m", bad)
    blobdir = tmp_path / "reg" / "library" / "m" / "blobs" / "sha256"
    assert not [p for p in os.listdir(blobdir)]                            # no temp file, no .meta left behind
    reg.put_blob("library/m", good, str(src), verify=True)
    assert reg.exists_blob("library/m", good)
    other = tmp_path / "other.bin"
    other.write_bytes(os.urandom(1000))
    with pytest.raises(modelx_b200.MxdError):
        reg.put_blob("library/m", good, str(other), verify=True)           # wrong body for an existing key ...
    assert (blobdir / good.split(":")[1]).read_bytes() == src.read_bytes()  # ... the good blob is still there, intact
    reg.put_blob("library/m", bad, str(src), verify=False)                 # reference behaviour: stored unverified
    assert reg.exists_blob("library/m", bad)
    # tree-keyed blobs verify against the tree root
    _, root = any_engine.tree_digest(src.read_bytes())
    tree_key = modelx_b200.digest_string(root)
    reg.put_blob("library/t", tree_key, str(src), verify="tree")
    assert reg.exists_blob("library/t", tree_key)
    with pytest.raises(modelx_b200.MxdError) as ei:
        reg.put_blob("library/t", good, str(src), verify="tree")           # whole-file digest is not the tree root
    assert ei.value.status == N.MXC_ERR_DIGEST_INVALID and not reg.exists_blob("library/t", good)


def test_push_digest_cache_skips_unchanged_files(any_engine, tmp_path):
    """Opt-in digest cache: unchanged (size, mtime) -> no re-hash; a touched or edited file is re-hashed."""
    d, files = _model(tmp_path, big=8_000_000)
    cl = client.Client(any_engine)
    first = cl.push_digest_json(str(d), use_cache=True)
    cache = json.loads((d / ".modelx" / "digests.json").read_text())
    assert set(cache) == set(files) and cache["README.md"]["digest"] == "sha256:" + hashlib.sha256(files["README.md"]).hexdigest()
    b0 = any_engine.stats()["bytes_hashed"]
    assert cl.push_digest_json(str(d), use_cache=True) == first
    assert any_engine.stats()["bytes_hashed"] == b0                     # nothing was hashed the second time
    assert cl.push_digest_json(str(d)) == first                     # and the uncached path agrees
    # edit one file (same size, new mtime) and corrupt the cache entry of another: both get re-hashed correctly
    time.sleep(0.01)
    (d / "README.md").write_bytes(b"# MODEL\n")
    m = json.loads(cl.push_digest_json(str(d), use_cache=True))
    readme = [b for b in m["blobs"] if b["name"] == "README.md"][0]
    assert readme["digest"] == "sha256:" + hashlib.sha256(b"# MODEL\n").hexdigest()
    assert any_engine.stats()["bytes_hashed"] - b0 < 5_000_000 + 8_000_100          # only README (+ the uncached full pass above)
    (d / ".modelx" / "digests.json").write_text("{broken")
    assert json.loads(cl.push_digest_json(str(d), use_cache=True)) == m          # unreadable cache = no cache


def test_read_once_push_with_part_digests_into_an_uploader(any_engine, tmp_path):
    """S3Extension.Upload's consumer on the tee.  Each blob is read once; the same GPU rounds give
    the blob digest and the SHA-256 of every calcParts range; a refused part is re-read and re-sent (retry x3,
    extension_s3.go:133-148); at most max_concurrent writes are in flight (UploadPartConcurrency, :18)."""
    d, files = _model(tmp_path, big=12_000_000)

    class Up:
        def __init__(self):
            self.data, self.parts, self.done, self.restarts = {}, {}, {}, []
            self.fail_once = {(None, 1)}        # refuse part 1 of the big blob once
            self.inflight, self.max_inflight, self.lock = 0, 0, threading.Lock()

        def begin(self, blob, name, size, parts):
            self.data[blob] = bytearray(size); self.parts[blob] = (name, parts)

        def part_write(self, blob, part, offset, data):
            with self.lock:
                self.inflight += 1; self.max_inflight = max(self.max_inflight, self.inflight)
            try:
                name, parts = self.parts[blob]
                assert parts[part][0] <= offset and offset + len(data) <= parts[part][0] + parts[part][1]   # never straddles a part
                if name.endswith(".safetensors") and (None, part) in self.fail_once:
                    self.fail_once.discard((None, part))
                    return False
                self.data[blob][offset:offset + len(data)] = data
                return True
            finally:
                with self.lock:
                    self.inflight -= 1

        def part_restart(self, blob, part):
            self.restarts.append((self.parts[blob][0], part))

        def complete(self, blob, digest, part_digests):
            self.done[self.parts[blob][0]] = (digest, part_digests)
            return "done"

    up = Up()
    cl = client.Client(any_engine)
    r0 = any_engine.stats()["src_bytes_read"]
    rep = cl.push_stream(str(d), up, force_multipart=True, max_concurrent=3)
    total = sum(len(v) for v in files.values())
    assert any_engine.stats()["src_bytes_read"] - r0 == total                       # one read of every file ...
    big = "model-00001.safetensors"
    assert rep["reread_bytes"] == 4_000_000 and up.restarts == [(big, 1)]           # ... plus the one refused part, re-sent
    assert up.max_inflight <= 3
    by_name = {b["name"]: b for b in rep["blobs"]}
    for name, data in files.items():
        i = [k for k, v in up.parts.items() if v[0] == name][0]
        assert bytes(up.data[i]) == data
        want_parts = modelx_b200.calc_parts(len(data), modelx_b200.server_part_count(len(data), True)) if data else [(0, 0)]
        assert [(p["offset"], p["length"]) for p in by_name[name]["parts"]] == want_parts      # extension_s3.go:99-112
        assert by_name[name]["digest"] == "sha256:" + hashlib.sha256(data).hexdigest() == up.done[name][0]
        for p, got in zip(by_name[name]["parts"], up.done[name][1]):
            want = hashlib.sha256(data[p["offset"]:p["offset"] + p["length"]]).digest()
            assert got == want and p["sha256"] == want.hex()
    m = rep["manifest"]
    assert [b["digest"] for b in m["blobs"]] == [by_name[b["name"]]["digest"] for b in m["blobs"]]


def test_read_once_tree_keyed_push_and_pull(any_engine, oracle, tmp_path):
    """every blob is read once -- the ring feeds the GPU (tree digest) and the store in the same pass --
    and is stored under its tree root; pull verifies with the tree digest."""
    d, files = _model(tmp_path, big=40_000_000)
    reg = client.LocalRegistry(str(tmp_path / "reg"), any_engine)
    cl = client.Client(any_engine)
    b0 = any_engine.stats()
    rep = cl.push_tree(reg, "library/llama", "v1", str(d))
    b1 = any_engine.stats()
    total = sum(len(v) for v in files.values())
    assert b1["src_bytes_read"] - b0["src_bytes_read"] == total            # bytes read from disk == blob bytes
    assert total <= b1["bytes_hashed"] - b0["bytes_hashed"] < total * 1.01 + 4096   # each byte went through the leaf kernel once (+ tree levels)
    base = tmp_path / "reg" / "library" / "llama" / "blobs" / "sha256"
    m = json.loads(rep["manifest_json"])
    for desc in m["blobs"] + [m["config"]]:
        data = files[desc["name"]]
        _, _, root = oracle.tree_digest(data, 8 << 20, 16 << 10, 8)
        assert desc["digest"] == modelx_b200.digest_string(root)
        nch = max(1, -(-len(data) // (8 << 20)))
        assert desc["annotations"]["modelx.digest"] == f"tree.v1;leaf=16384;fanout=8;chunk=8388608;chunks={nch}"
        hexd = desc["digest"].split(":")[1]
        if len(data) == 0:
            assert not (base / hexd).exists()
        else:
            assert (base / hexd).read_bytes() == data                       # written by the tee, not by a second read
            assert json.loads((base / (hexd + ".meta")).read_text())["contentLength"] == len(data)
    assert not [p for p in os.listdir(base) if p.startswith(".incoming")]  # temporaries renamed or removed
    assert {b["name"]: b["status"] for b in rep["blobs"]}["empty.txt"] == "empty"
    assert {b["status"] for b in cl.push_tree(reg, "library/llama", "v2", str(d))["blobs"]} == {"exists", "empty"}
    into = tmp_path / "pulled"
    res = {r["name"]: r["status"] for r in cl.pull(reg, "library/llama", "v1", str(into))}
    assert set(res.values()) == {"done", "empty"}
    for n, data in files.items():
        assert (into / n).read_bytes() == data
    assert set(r["status"] for r in cl.pull(reg, "library/llama", "v1", str(into))) == {"already exists"}
    (into / "tokenizer.json").write_bytes(b"x")
    assert {r["name"]: r["state"] for r in cl.pull_check(str(into), rep["manifest_json"])}["tokenizer.json"] == "differs"
    # a corrupted tree-keyed store blob is rejected on pull as well
    dg = {b["name"]: b["digest"] for b in rep["blobs"]}["model-00001.safetensors"]
    raw = bytearray((base / dg.split(":")[1]).read_bytes()); raw[123] ^= 1; (base / dg.split(":")[1]).write_bytes(raw)
    with pytest.raises(modelx_b200.MxdError) as ei:
        cl.pull(reg, "library/llama", "v1", str(tmp_path / "pulled2"))
    assert ei.value.status == N.MXC_ERR_DIGEST_INVALID and not (tmp_path / "pulled2" / "model-00001.safetensors").exists()


def test_tee_sink_sees_every_byte_once(any_engine, tmp_path):
    size = 70_000_000 + 3
    data = os.urandom(size)
    p = tmp_path / "b.bin"
    p.write_bytes(data)
    got = bytearray(size)
    seen = []
    lock = threading.Lock()

    def sink(offset, piece):
        got[offset:offset + len(piece)] = piece
        with lock:
            seen.append((offset, len(piece)))

    chunks, root, sz = any_engine.tree_digest_file_tee(str(p), sink)
    assert sz == size and bytes(got) == data
    seen.sort()
    assert seen[0][0] == 0 and all(seen[i][0] + seen[i][1] == seen[i + 1][0] for i in range(len(seen) - 1))
    assert seen[-1][0] + seen[-1][1] == size and max(ln for _, ln in seen) <= 4 << 20
    assert (chunks, root) == any_engine.tree_digest_file(str(p))[:2]
