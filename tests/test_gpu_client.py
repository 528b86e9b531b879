"""GPU tests of the host mirror: the digest phase of Push, the check phase of Pull, and BASELINE
config 1 (one 64 MB random blob: digest + PutBlob into the in-process FS registry) end to end."""
import datetime
import hashlib
import json
import os
import stat
import time

import pytest

import modelx_b200
from modelx_b200 import _native as N
from modelx_b200 import client

pytestmark = pytest.mark.gpu
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


def test_push_digest_phase_matches_reference_semantics(engine, tmp_path):
    d, files = _model(tmp_path, big=5_000_000)
    cl = client.Client(engine)
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


def test_push_digest_with_tree_annotation(engine, oracle, tmp_path):
    d, files = _model(tmp_path, big=20_000_000)
    m = json.loads(client.Client(engine).push_digest_json(str(d), with_tree=True))
    blob = [b for b in m["blobs"] if b["name"].endswith(".safetensors")][0]
    _, _, root = oracle.tree_digest(files[blob["name"]], 8 << 20, 16 << 10, 8)
    assert blob["annotations"]["modelx.tree.v1"] == \
        f"{modelx_b200.digest_string(root)};leaf=16384;fanout=8;chunk=8388608;chunks=3"
    assert blob["digest"] == "sha256:" + hashlib.sha256(files[blob["name"]]).hexdigest()   # wire-compatible identity kept


def test_directory_blobs_are_unsupported(engine, tmp_path):
    d, _ = _model(tmp_path, big=1000)
    (d / "subdir").mkdir()
    with pytest.raises(modelx_b200.MxdError) as ei:
        client.Client(engine).push_digest_json(str(d))
    assert ei.value.status == N.MXC_ERR_UNSUPPORTED


def test_config1_push_then_pull_through_local_registry(engine, tmp_path):
    """BASELINE config 1: 64 MB random blob, pkg/client digest + PutBlob against the in-process registry."""
    d, files = _model(tmp_path)
    reg = client.LocalRegistry(str(tmp_path / "data" / "registry"), engine)
    cl = client.Client(engine)
    rep = cl.push(reg, "library/llama", "v1", str(d), verify=True)
    status = {b["name"]: b["status"] for b in rep["blobs"]}
    assert status == {"README.md": "done", "empty.txt": "empty", "model-00001.safetensors": "done",
                      "tokenizer.json": "done", "modelx.yaml": "done"}
    base = tmp_path / "data" / "registry" / "library" / "llama"
    for b in rep["blobs"]:
        hexd = b["digest"].split(":")[1]
        if b["status"] == "empty":
            assert not (base / "blobs" / "sha256" / hexd).exists()          # push.go:165-168: never uploaded
            continue
        stored = (base / "blobs" / "sha256" / hexd).read_bytes()
        assert stored == files[b["name"]] and hashlib.sha256(stored).hexdigest() == hexd
        meta = json.loads((base / "blobs" / "sha256" / (hexd + ".meta")).read_text())
        assert meta == {"contentType": "application/octet-stream", "contentLength": len(stored)}
    assert (base / "manifests" / "v1").read_text() == rep["manifest_json"] == reg.get_manifest_json("library/llama", "v1")
    # second push: content addressed dedupe (HeadBlob, push.go:169-177)
    rep2 = cl.push(reg, "library/llama", "v2", str(d))
    assert {b["status"] for b in rep2["blobs"]} == {"exists", "empty"}

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


def test_put_blob_verify_rejects_wrong_digest(engine, tmp_path):
    """New behaviour: the stored body is re-hashed; a mismatch is DIGEST_INVALID and nothing is kept."""
    reg = client.LocalRegistry(str(tmp_path / "reg"), engine)
    src = tmp_path / "b.bin"
    src.write_bytes(os.urandom(3_000_000))
    good = "sha256:" + hashlib.sha256(src.read_bytes()).hexdigest()
    bad = "sha256:" + "00" * 32
    with pytest.raises(modelx_b200.MxdError) as ei:
        reg.put_blob("library/m", bad, str(src), verify=True)
    assert ei.value.status == N.MXC_ERR_DIGEST_INVALID and "digest invalid" in ei.value.detail
    assert not reg.exists_blob("library/m", bad)
    reg.put_blob("library/m", good, str(src), verify=True)
    assert reg.exists_blob("library/m", good)
    reg.put_blob("library/m", bad, str(src), verify=False)       # reference behaviour: stored unverified
    assert reg.exists_blob("library/m", bad)
    # tree-keyed blobs verify against the tree root
    _, root = engine.tree_digest(src.read_bytes())
    tree_key = modelx_b200.digest_string(root)
    reg.put_blob("library/t", tree_key, str(src), verify="tree")
    assert reg.exists_blob("library/t", tree_key)
    with pytest.raises(modelx_b200.MxdError) as ei:
        reg.put_blob("library/t", good, str(src), verify="tree")   # whole-file digest is not the tree root
    assert ei.value.status == N.MXC_ERR_DIGEST_INVALID and not reg.exists_blob("library/t", good)


def test_push_digest_cache_skips_unchanged_files(engine, tmp_path):
    """Opt-in digest cache: unchanged (size, mtime) -> no re-hash; a touched or edited file is re-hashed."""
    d, files = _model(tmp_path, big=8_000_000)
    cl = client.Client(engine)
    first = cl.push_digest_json(str(d), use_cache=True)
    cache = json.loads((d / ".modelx" / "digests.json").read_text())
    assert set(cache) == set(files) and cache["README.md"]["digest"] == "sha256:" + hashlib.sha256(files["README.md"]).hexdigest()
    b0 = engine.stats()["bytes_hashed"]
    assert cl.push_digest_json(str(d), use_cache=True) == first
    assert engine.stats()["bytes_hashed"] == b0                     # nothing was hashed the second time
    assert cl.push_digest_json(str(d)) == first                     # and the uncached path agrees
    # edit one file (same size, new mtime) and corrupt the cache entry of another: both get re-hashed correctly
    time.sleep(0.01)
    (d / "README.md").write_bytes(b"# MODEL\n")
    m = json.loads(cl.push_digest_json(str(d), use_cache=True))
    readme = [b for b in m["blobs"] if b["name"] == "README.md"][0]
    assert readme["digest"] == "sha256:" + hashlib.sha256(b"# MODEL\n").hexdigest()
    assert engine.stats()["bytes_hashed"] - b0 < 5_000_000 + 8_000_100          # only README (+ the uncached full pass above)
    (d / ".modelx" / "digests.json").write_text("{broken")
    assert json.loads(cl.push_digest_json(str(d), use_cache=True)) == m          # unreadable cache = no cache


def test_read_once_tree_keyed_push_and_pull(engine, oracle, tmp_path):
    """every blob is read once -- the ring feeds the GPU (tree digest) and the store in the same pass --
    and is stored under its tree root; pull verifies with the tree digest."""
    d, files = _model(tmp_path, big=40_000_000)
    reg = client.LocalRegistry(str(tmp_path / "reg"), engine)
    cl = client.Client(engine)
    b0 = engine.stats()["bytes_hashed"]
    rep = cl.push_tree(reg, "library/llama", "v1", str(d))
    hashed = engine.stats()["bytes_hashed"] - b0
    total = sum(len(v) for v in files.values())
    assert total <= hashed < total * 1.01 + 4096            # each byte went through the leaf kernel exactly once (+ tree levels)
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


def test_tee_sink_sees_every_byte_once(engine, tmp_path):
    size = 70_000_000 + 3
    data = os.urandom(size)
    p = tmp_path / "b.bin"
    p.write_bytes(data)
    got = bytearray(size)
    seen = []
    import threading
    lock = threading.Lock()

    def sink(offset, piece):
        got[offset:offset + len(piece)] = piece
        with lock:
            seen.append((offset, len(piece)))

    chunks, root, sz = engine.tree_digest_file_tee(str(p), sink)
    assert sz == size and bytes(got) == data
    seen.sort()
    assert seen[0][0] == 0 and all(seen[i][0] + seen[i][1] == seen[i + 1][0] for i in range(len(seen) - 1))
    assert seen[-1][0] + seen[-1][1] == size and max(l for _, l in seen) <= 4 << 20
    assert (chunks, root) == engine.tree_digest_file(str(p))[:2]
