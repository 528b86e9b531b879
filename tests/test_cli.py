"""The re-hosted CLI (modelx_b200/cli.py): reference syntax, init template, and init -> push -> pull -> list against a
local-store registry.  The flows run on the test double here and on the CUDA build on the B200."""
import hashlib
import json
import os

import pytest
import yaml

from modelx_b200 import cli


@pytest.fixture
def repos(tmp_path, monkeypatch):
    monkeypatch.setenv("MODELX_REPOS_FILE", str(tmp_path / "home" / ".modelx" / "repos.json"))
    monkeypatch.delenv("MODELX_AUTH", raising=False)
    rm = cli.RepoManager()
    rm.set("myrepo", "https://registry.example.com", token="tok")
    rm.set("local", "file://" + str(tmp_path / "registry"))
    return rm


def test_parse_reference_follows_reference_go(repos):
    """cmd/modelx/model/reference.go:33-86, incl. the examples of `modelx push/pull --help` (push.go:19-32, pull.go:16-29).
    The reference's own reference_test.go is stale against this code (it predates the "library/" default):
    its second case still holds and is checked verbatim, the other two are checked against what the code does."""
    r = cli.parse_reference("https://registry.example.com:8443/repository/name@v1")
    assert (r.registry, r.repository, r.version) == ("https://registry.example.com:8443", "repository/name", "v1")
    r = cli.parse_reference("https://registry.example.com/repository@sha256:abcdef")
    assert (r.registry, r.repository, r.version) == ("https://registry.example.com", "library/repository", "sha256:abcdef")
    r = cli.parse_reference("https://registry.example.com/repo/name")
    assert (r.registry, r.repository, r.version) == ("https://registry.example.com", "repo/name", "")
    assert str(r) == "https://registry.example.com/repo/name"
    # the CLI examples: <repo alias>/<project>/<name>[@version]
    r = cli.parse_reference("myrepo/project/demo", repos)
    assert (r.registry, r.repository, r.version, r.authorization) == ("https://registry.example.com", "project/demo", "", "Bearer tok")
    r = cli.parse_reference("myrepo/project/demo@v1", repos)
    assert (r.repository, r.version) == ("project/demo", "v1") and str(r) == "https://registry.example.com/project/demo@v1"
    r = cli.parse_reference("myrepo/demo@v2", repos)
    assert r.repository == "library/demo"                                  # reference.go:75-77
    r = cli.parse_reference("https://myrepo/project/demo?token=abc")
    assert r.authorization == "Bearer abc" and r.repository == "project/demo"
    with pytest.raises(KeyError):
        cli.parse_reference("nosuchrepo/project/demo", repos)
    r = cli.parse_reference("local/project/demo@v1", repos)
    assert r.registry.startswith("file://") and (r.repository, r.version) == ("project/demo", "v1")


def test_parse_reference_unknown_alias_without_scheme_is_an_error(repos):
    with pytest.raises(KeyError):
        cli.parse_reference("registry.example.com/project/demo@v3", repos)


def test_init_writes_the_reference_template(tmp_path, capsys):
    """InitModelx (init.go:39-104): modelx.yaml = yaml.Marshal of the template ModelConfig, README.md, mode 0755."""
    path = tmp_path / "demo"
    cli.init_model(str(path))
    cfg = yaml.safe_load((path / "modelx.yaml").read_text())
    assert cfg["description"] == "This is a modelx model" and cfg["framework"] == "<some framework>"
    assert cfg["tags"] == ["modelx", "<other>"] and cfg["mantainers"] == ["maintainer"] and cfg["modelfiles"] == []
    assert cfg["resources"] == {"cpu": "4", "memory": "16Gi", "gpu": {"nvdia": {"nvdia/gpu": "1"},
                                "gpu-manager": {"tencent.com/vcuda-core": "50", "tencent.com/vcuda-memory": "25"}}}
    assert cfg["config"] == {"inputs": {}, "outputs": {}}
    assert (path / "README.md").read_text() == "# demo\n\nAwesome model descrition.\n"
    assert os.stat(path / "modelx.yaml").st_mode & 0o777 == 0o755
    assert "Modelx model initialized in" in capsys.readouterr().out
    with pytest.raises(FileExistsError):
        cli.init_model(str(path))
    (path / "README.md").write_text("mine")
    cli.init_model(str(path), force=True)
    assert (path / "README.md").read_text() == "mine"          # init.go:93-99: README only when absent


def test_init_push_pull_list_roundtrip(backend, repos, tmp_path, capsys, monkeypatch):
    """init -> push -> pull against the local store; the version shows up in index.json."""
    model = tmp_path / "llama"
    assert cli.main(["init", str(model)]) == 0
    weights = os.urandom(5_000_000)
    (model / "model.safetensors").write_bytes(weights)
    (model / "tokenizer").mkdir()
    (model / "tokenizer" / "vocab.json").write_text('{"a": 0}')
    lib = ["--lib", backend] if backend else []
    assert cli.main(lib + ["push", "local/project/llama@v1", str(model)]) == 0
    out = capsys.readouterr().out
    assert "Pushing to file://" in out and "model.safetensors" in out and "done" in out
    store = tmp_path / "registry"
    idx = json.loads((store / "project" / "llama" / "index.json").read_text())
    assert [m["name"] for m in idx["manifests"]] == ["v1"]
    glob = json.loads((store / "index.json").read_text())
    assert [m["name"] for m in glob["manifests"]] == ["project/llama"]
    man = json.loads((store / "project" / "llama" / "manifests" / "v1").read_text())
    blob = [b for b in man["blobs"] if b["name"] == "model.safetensors"][0]
    assert blob["digest"] == "sha256:" + hashlib.sha256(weights).hexdigest()
    assert (store / "project" / "llama" / "blobs" / "sha256" / blob["digest"][7:]).read_bytes() == weights
    # default version is "latest" (pkg/client/registry.go:34-36); default target dir is the repository's base name
    assert cli.main(lib + ["push", "local/project/llama", str(model)]) == 0
    assert (store / "project" / "llama" / "manifests" / "latest").exists()
    monkeypatch.chdir(tmp_path)
    assert cli.main(lib + ["pull", "local/project/llama@v1"]) == 0
    assert "Pulling file://" in capsys.readouterr().out
    assert (tmp_path / "llama" / "model.safetensors").read_bytes() == weights   # pulled over the source dir: "already exists"
    assert cli.main(lib + ["pull", "local/project/llama@v1", str(tmp_path / "copy")]) == 0
    assert (tmp_path / "copy" / "model.safetensors").read_bytes() == weights
    assert json.loads((tmp_path / "copy" / "tokenizer" / "vocab.json").read_text()) == {"a": 0}
    assert yaml.safe_load((tmp_path / "copy" / "modelx.yaml").read_text())["description"] == "This is a modelx model"
    capsys.readouterr()
    assert cli.main(lib + ["list", "local/project/llama"]) == 0
    listed = capsys.readouterr().out
    assert "v1" in listed and "latest" in listed
    assert cli.main(lib + ["list", "local/project/llama@v1"]) == 0
    assert "model.safetensors" in capsys.readouterr().out
    # tree-keyed push through the CLI
    assert cli.main(lib + ["push", "--tree", "local/project/llama@t1", str(model)]) == 0
    tman = json.loads((store / "project" / "llama" / "manifests" / "t1").read_text())
    assert all(b["annotations"]["modelx.digest"].startswith("tree.v1;") for b in tman["blobs"])
    assert cli.main(lib + ["pull", "local/project/llama@t1", str(tmp_path / "copy2")]) == 0
    assert (tmp_path / "copy2" / "model.safetensors").read_bytes() == weights


def test_http_registries_are_reported_as_unsupported(repos, tmp_path):
    (tmp_path / "m").mkdir()
    (tmp_path / "m" / "modelx.yaml").write_text("description: x\n")
    with pytest.raises(SystemExit) as ei:
        cli.push_model("myrepo/project/demo@v1", str(tmp_path / "m"), repos=repos, lib_path="unused")
    assert "UNSUPPORTED" in str(ei.value)
    with pytest.raises(SystemExit) as ei:
        cli.push_model("local/project/demo@v1", str(tmp_path / "nomodel"), repos=repos)
    assert "read model config" in str(ei.value)
