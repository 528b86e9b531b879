"""CPU-side checks of the C++ host mirror (include/modelx_client.h): ParseManifest, the JSON encoding
of pkg/types, and the local FS store layout.  Nothing here hashes (no GPU needed)."""
import json
import os
import re

import pytest

import modelx_b200
from modelx_b200 import _native as N
from modelx_b200 import client

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZERO = "0001-01-01T00:00:00Z"
FILE = "application/vnd.modelx.model.file.v1"
DIRT = "application/vnd.modelx.model.directory.v1.tar+gz"
CONF = "application/vnd.modelx.model.config.v1.yaml"
MANI = "application/vnd.modelx.model.manifest.v1.json"


def test_client_header_and_library_agree():
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "modelx_client.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(mxc_[a-z0-9_]+)\s*\(", text)))
    lib = modelx_b200.load()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n) and n in N.PROTOTYPES, n


def _model_dir(tmp_path):
    d = tmp_path / "model"
    d.mkdir()
    (d / "modelx.yaml").write_text("framework: pytorch\n")
    (d / "weights.bin").write_bytes(b"\x01" * 1000)
    (d / "README.md").write_text("hi")
    (d / "a<&>.txt").write_text("")
    (d / ".git").mkdir()
    (d / ".hidden").write_text("x")
    (d / "tokenizer").mkdir()
    return d


def test_parse_manifest_matches_go_encoding(tmp_path):
    """ParseManifest, push.go:67-100, rendered as encoding/json would: struct field order, omitempty,
    HTML-safe escapes, zero time.Time, blobs sorted by name (bytewise), dot-files skipped."""
    d = _model_dir(tmp_path)
    got = client.parse_manifest_json(str(d))
    want = ('{"schemaVersion":0,"mediaType":"%s","config":{"name":"modelx.yaml","mediaType":"%s","modified":"%s"},"blobs":['
            '{"name":"README.md","mediaType":"%s","modified":"%s"},'
            '{"name":"a\\u003c\\u0026\\u003e.txt","mediaType":"%s","modified":"%s"},'
            '{"name":"tokenizer","mediaType":"%s","modified":"%s"},'
            '{"name":"weights.bin","mediaType":"%s","modified":"%s"}]}') % (MANI, CONF, ZERO, FILE, ZERO, FILE, ZERO, DIRT, ZERO, FILE, ZERO)
    assert got == want
    with open(os.path.join(os.path.dirname(__file__), "golden", "parse_manifest_expected.json")) as f:
        assert got == f.read().strip()
    m = json.loads(got)
    assert [b["name"] for b in m["blobs"]] == ["README.md", "a<&>.txt", "tokenizer", "weights.bin"]


def test_parse_manifest_empty_dir_has_null_blobs(tmp_path):
    d = tmp_path / "empty"
    d.mkdir()
    # nil slice -> null; no config entry -> zero Descriptor
    assert client.parse_manifest_json(str(d)) == \
        '{"schemaVersion":0,"mediaType":"%s","config":{"name":"","modified":"%s"},"blobs":null}' % (MANI, ZERO)
    with pytest.raises(modelx_b200.MxdError) as ei:
        client.parse_manifest_json(str(tmp_path / "nope"))
    assert ei.value.status == N.MXD_ERR_IO


def test_blob_digest_path():
    dg = "sha256:" + "ab" * 32
    assert client.blob_digest_path("library/llama", dg) == "library/llama/blobs/sha256/" + "ab" * 32   # store.go:56-61
    with pytest.raises(modelx_b200.MxdError):
        client.blob_digest_path("r", "nodigest")


def test_fs_store_layout_without_gpu(tmp_path):
    """FSRegistryStore.PutBlob -> LocalFSProvider.Put: <hex>.meta (indented JSON) then <hex>."""
    reg = client.LocalRegistry(str(tmp_path / "data" / "registry"))
    src = tmp_path / "blob.bin"
    src.write_bytes(os.urandom(12345))
    dg = "sha256:" + "cd" * 32          # the reference does not verify the digest on PutBlob (registry.go:144-164)
    assert not reg.exists_blob("library/m", dg)
    reg.put_blob("library/m", dg, str(src))
    base = tmp_path / "data" / "registry" / "library" / "m" / "blobs" / "sha256"
    assert (base / ("cd" * 32)).read_bytes() == src.read_bytes()
    assert (base / ("cd" * 32 + ".meta")).read_text() == \
        '{\n  "contentType": "application/octet-stream",\n  "contentLength": 12345\n}'   # json.MarshalIndent(meta, "", "  ")
    assert reg.exists_blob("library/m", dg)
    for bad in ("sha256:xyz", "sha256:" + "AB" * 32, "md5:" + "0" * 32):
        with pytest.raises(modelx_b200.MxdError) as ei:
            reg.put_blob("library/m", bad, str(src))
        assert ei.value.status == N.MXC_ERR_DIGEST_INVALID
    with pytest.raises(modelx_b200.MxdError):
        reg.put_blob("library/m", dg, str(src), content_type="")     # registry.go:147-151
    with pytest.raises(modelx_b200.MxdError) as ei:
        reg.put_blob("library/m", dg, str(src), verify=True)         # verify needs the GPU engine
    assert ei.value.status == N.MXD_ERR_INVALID


def test_manifest_roundtrip_is_byte_stable(tmp_path):
    reg = client.LocalRegistry(str(tmp_path / "reg"))
    text = ('{"schemaVersion":1,"mediaType":"%s","config":{"name":"modelx.yaml","mediaType":"%s","digest":"sha256:%s",'
            '"size":10,"mode":420,"modified":"2024-05-06T07:08:09.123456789+08:00"},"blobs":[{"name":"w.bin","mediaType":"%s",'
            '"digest":"sha256:%s","size":5,"mode":493,"urls":["http://x/y"],"modified":"2023-01-02T03:04:05Z",'
            '"annotations":{"a":"1","b":"\\u003cx\\u003e"}}],"annotations":{"k":"v"}}') % (MANI, CONF, "11" * 32, FILE, "22" * 32)
    reg.put_manifest("library/m", "v1", text)
    assert reg.get_manifest_json("library/m", "v1") == text            # decode + json.Marshal is the identity here
    meta = (tmp_path / "reg" / "library" / "m" / "manifests" / "v1.meta").read_text()
    assert json.loads(meta) == {"contentType": MANI, "contentLength": len(text)}
    # key order / whitespace of the input do not matter, output is canonical Go order
    shuffled = json.dumps(json.loads(text), indent=2, sort_keys=True)
    reg.put_manifest("library/m", "v2", shuffled)
    assert reg.get_manifest_json("library/m", "v2") == text
    with pytest.raises(modelx_b200.MxdError) as ei:
        reg.put_manifest("library/m", "v3", "{not json")
    assert ei.value.status == N.MXC_ERR_MANIFEST
    with pytest.raises(modelx_b200.MxdError) as ei:
        reg.get_manifest_json("library/m", "missing")
    assert ei.value.status == N.MXC_ERR_NOT_FOUND


def test_json_string_escaping_follows_encoding_json(tmp_path):
    """File names with quotes, control characters, non-ASCII, U+2028 and invalid UTF-8 are rendered as Go's
    encoding/json would (HTML-safe escapes, invalid bytes -> \\ufffd)."""
    d = tmp_path / "m"
    d.mkdir()
    (d / "modelx.yaml").write_text("x")
    names = [b'q"uote', b"tab\there", b"uni-\xe6\xa8\xa1\xe5\x9e\x8b.bin", b"ls\xe2\x80\xa8sep", b"bad\xff\xfebytes", b"trunc\xe6\xa8", b"back\\slash"]
    for n in names:
        with open(os.path.join(os.fsencode(str(d)), n), "wb") as f:
            f.write(b"1")
    got = client.parse_manifest_json(str(d))
    blobs = got[got.index('"blobs":['):]
    for frag in ('"q\\"uote"', '"tab\\there"', '"uni-模型.bin"', '"ls\\u2028sep"', '"bad\\ufffd\\ufffdbytes"',
                 '"trunc\\ufffd\\ufffd"', '"back\\\\slash"'):
        assert frag in blobs, frag
    m = json.loads(got)                       # and it is valid JSON
    assert len(m["blobs"]) == len(names)
    # byte-wise name order, as strings.Compare sorts them (push.go:98)
    def go_decode(b):   # Go replaces every undecodable byte by one U+FFFD (Python's "replace" merges truncated sequences)
        return "".join("\ufffd" if 0xDC80 <= ord(ch) <= 0xDCFF else ch for ch in b.decode("utf-8", "surrogateescape"))
    assert [b["name"] for b in m["blobs"]] == [go_decode(n) for n in sorted(names)]


def test_manifest_parser_handles_escapes_and_rejects_garbage(tmp_path):
    reg = client.LocalRegistry(str(tmp_path / "reg"))
    # \\uXXXX escapes (incl. a surrogate pair), nested unknown fields, numbers with exponents are tolerated on input
    text = ('{"schemaVersion": 1, "unknown": {"a": [1, 2.5e3, true, null]}, "config": {"name": "c\\u00e9\\ud83d\\ude00.yaml", '
            '"modified": "2024-01-01T00:00:00Z"}, "blobs": [{"name": "w\\tx", "size": 7, "modified": "2024-01-01T00:00:00Z"}]}')
    reg.put_manifest("r", "v", text)
    out = reg.get_manifest_json("r", "v")
    assert out == ('{"schemaVersion":1,"config":{"name":"cé😀.yaml","modified":"2024-01-01T00:00:00Z"},'
                   '"blobs":[{"name":"w\\tx","size":7,"modified":"2024-01-01T00:00:00Z"}]}')
    assert json.loads(out)["config"]["name"] == "cé😀.yaml"
    for bad in ('', '[]', '{"config": 5}', '{"blobs": [1]}', '{"a": "unterminated}', '{"a": 1,}', '{' * 100):
        with pytest.raises(modelx_b200.MxdError) as ei:
            reg.put_manifest("r", "bad", bad)
        assert ei.value.status == N.MXC_ERR_MANIFEST, bad
