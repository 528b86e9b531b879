"""``modelx-b200`` -- the modelx CLI (cmd/modelx: init / push / pull / list / repo) re-hosted over the C ABI.

Same commands, arguments and reference syntax as the Go CLI:

    modelx-b200 init <path> [-f]                       cmd/modelx/model/init.go:14-104
    modelx-b200 push <repo>/<project>/<name>[@version] [dir]    cmd/modelx/model/push.go:15-80
    modelx-b200 pull <repo>/<project>/<name>[@version] [into]   cmd/modelx/model/pull.go:12-69
    modelx-b200 list <repo>[/<project>/<name>[@version]]        cmd/modelx/model/list.go
    modelx-b200 repo add|list|remove                            cmd/modelx/repo/*.go

What differs: the transport.  modelx talks HTTP to modelxd (out of scope here); this CLI talks to a
registry laid out exactly like modelxd's local store (pkg/registry: <base>/<repository>/{blobs,manifests,index.json}),
addressed as ``file:///path`` -- add one with ``repo add``.  A directory it writes can be served by a stock modelxd.
All hashing happens on the GPU through libmodelxdigest.so (push reads every blob once; pull verifies what it copies).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from dataclasses import dataclass
from typing import List, Optional
from urllib.parse import parse_qs, urlsplit

MODEL_CONFIG_FILE = "modelx.yaml"      # ModelConfigFileName
README_FILE = "README.md"
SPLITOR_REPO = "/"                     # cmd/modelx/repo/list.go:12-13
SPLITOR_VERSION = "@"
AUTH_ENV = "MODELX_AUTH"


# ---- cmd/modelx/repo/repo.go: ~/.modelx/repos.json --------------------------------------------------------------
class RepoManager:
    def __init__(self, path: Optional[str] = None):
        self.path = path or os.environ.get("MODELX_REPOS_FILE") or os.path.join(os.path.expanduser("~"), ".modelx", "repos.json")

    def _load(self) -> dict:
        try:
            with open(self.path) as f:
                return json.load(f) or {}
        except FileNotFoundError:
            os.makedirs(os.path.dirname(self.path), exist_ok=True)
            return {}

    def _save(self, data: dict) -> None:
        os.makedirs(os.path.dirname(self.path), exist_ok=True)
        with open(self.path, "w") as f:
            json.dump(data, f, indent="\t")

    def set(self, name: str, url: str, token: str = "") -> None:
        parts = urlsplit(url)
        if not parts.scheme or not (parts.netloc or parts.path):
            raise ValueError(f"invalid url: {url}")
        data = self._load()
        repos = data.get("repos") or []
        item = {k: v for k, v in (("name", name), ("url", url), ("token", token)) if v}
        for i, r in enumerate(repos):
            if r.get("name") == name:
                repos[i] = item
                break
        else:
            repos.append(item)
        data["repos"] = repos
        self._save(data)

    def get(self, name: str) -> dict:
        for r in self._load().get("repos") or []:
            if r.get("name") == name or r.get("url") == name:
                return r
        raise KeyError(f"repo {name} not found")

    def remove(self, name: str) -> None:
        data = self._load()
        repos = data.get("repos") or []
        for i, r in enumerate(repos):
            if r.get("name") == name:
                del repos[i]
                data["repos"] = repos
                self._save(data)
                return
        raise KeyError(f"repo {name} not found")

    def list(self) -> List[dict]:
        return self._load().get("repos") or []


# ---- cmd/modelx/model/reference.go:14-86 ---------------------------------------------------------------------------
@dataclass
class Reference:
    registry: str
    repository: str
    version: str
    authorization: str = ""

    def __str__(self) -> str:                       # Reference.String(), reference.go:22-27
        if not self.version:
            return f"{self.registry}/{self.repository}"
        return f"{self.registry}/{self.repository}@{self.version}"


def parse_reference(raw: str, repos: Optional[RepoManager] = None) -> Reference:
    """ParseReference (reference.go:33-86), statement by statement; ``file://`` registries are this re-host's addition."""
    auth = os.environ.get(AUTH_ENV, "")
    if "://" not in raw:                            # <repo alias>/<project>/<name>[@version]
        splits = raw.split(SPLITOR_REPO, 1)
        details = (repos or RepoManager()).get(splits[0])
        if not auth:
            auth = "Bearer " + details.get("token", "")
        raw = details["url"] + "/" + splits[1] if len(splits) == 2 else details["url"]
    is_file = raw.startswith("file://")
    if not is_file and not raw.startswith("http://") and not raw.startswith("https://"):
        raw = "https://" + raw
    u = urlsplit(raw)
    if is_file:
        # file:///base/path//project/name@v  is ambiguous, so a file registry is only reachable through an alias whose
        # url is the base: the alias url is a prefix of `raw` here
        base = None
        for r in (repos or RepoManager()).list():
            if r.get("url", "").startswith("file://") and (raw == r["url"] or raw.startswith(r["url"].rstrip("/") + "/")):
                if base is None or len(r["url"]) > len(base):
                    base = r["url"].rstrip("/")
        if base is None:
            raise ValueError("invalid reference: a file:// registry must be added with `repo add` and used through its name")
        registry, path = base, raw[len(base):]
    else:
        if not u.netloc:
            raise ValueError("invalid reference: missing host")
        token = parse_qs(u.query).get("token", [""])[0]
        if token:
            auth = "Bearer " + token
        registry, path = f"{u.scheme}://{u.netloc}", u.path
    splits = path.split(SPLITOR_VERSION, 1)
    version = splits[1] if len(splits) == 2 and splits[1] else ""
    repository = splits[0][1:] if splits[0] else ""
    if repository and "/" not in repository:
        repository = "library/" + repository        # reference.go:75-77
    return Reference(registry, repository, version, auth)


# ---- cmd/modelx/model/init.go:39-104 ---------------------------------------------------------------------------------
# yaml.Marshal(ModelConfig{...}) of gopkg.in/yaml.v3: struct fields in declaration order (config.go:8-18; yaml.v3 takes
# the lower-cased Go field name when there is no yaml tag), map keys sorted, 4-space indent, empty values spelled out.
INIT_TEMPLATE = """description: This is a modelx model
framework: <some framework>
task: ""
tags:
    - modelx
    - <other>
resources:
    cpu: "4"
    gpu:
        gpu-manager:
            tencent.com/vcuda-core: "50"
            tencent.com/vcuda-memory: "25"
        nvdia:
            nvdia/gpu: "1"
    memory: 16Gi
mantainers:
    - maintainer
annotations: {}
modelfiles: []
config:
    inputs: {}
    outputs: {}
"""


def init_model(path: str, force: bool = False) -> None:
    if os.path.exists(path) and not force:
        raise FileExistsError(f"path {path} already exists")
    os.makedirs(path, mode=0o755, exist_ok=True)
    cfg = os.path.join(path, MODEL_CONFIG_FILE)
    with open(cfg, "w") as f:
        f.write(INIT_TEMPLATE)
    os.chmod(cfg, 0o755)
    base = os.path.basename(os.path.normpath(path))
    readme = os.path.join(path, README_FILE)
    if base and not os.path.exists(readme):
        with open(readme, "w") as f:
            f.write(f"# {base}\n\nAwesome model descrition.\n")
        os.chmod(readme, 0o755)
    print(f"Modelx model initialized in {path}")


# ---- push / pull / list ------------------------------------------------------------------------------------------------
def _store_of(ref: Reference) -> str:
    if not ref.registry.startswith("file://"):
        raise SystemExit(f"UNSUPPORTED: {ref.registry}: the HTTP transport to modelxd is out of scope for modelx-b200; "
                         "use a file:// registry (a modelxd local-store directory) added with `repo add`")
    return ref.registry[len("file://"):] or "/"


def _engine(lib_path: Optional[str]):
    import modelx_b200
    return modelx_b200.Engine(lib_path=lib_path)


def _print_status(rows, key):
    for r in rows:
        print(f"{r['name']:<40} {r[key]}")


def push_model(raw_ref: str, directory: str = "", tree: bool = False, repos=None, lib_path=None) -> dict:
    """PushModel (cmd/modelx/model/push.go:61-80)."""
    from .client import Client, LocalRegistry
    ref = parse_reference(raw_ref, repos)
    directory = directory or "."
    cfg = os.path.join(directory, MODEL_CONFIG_FILE)
    try:
        with open(cfg) as f:
            import yaml
            yaml.safe_load(f)                      # parse model config (push.go:70-77)
    except OSError as e:
        raise SystemExit(f"read model config:{MODEL_CONFIG_FILE} {e}")
    except Exception as e:  # yaml error
        raise SystemExit(f"parse model config:{MODEL_CONFIG_FILE} {e}")
    if not ref.repository:
        raise SystemExit("repository is not specified")
    version = ref.version or "latest"               # pkg/client/registry.go:34-36
    print(f"Pushing to {ref} ")
    store = _store_of(ref)
    if not tree:
        # routing advice (mxd_batch_pays_off): a whole-file SHA-256 is one serial chain, ~0.07 GB/s on a GPU lane
        import modelx_b200
        sizes = [os.path.getsize(os.path.join(directory, n)) for n in os.listdir(directory)
                 if not n.startswith(".") and os.path.isfile(os.path.join(directory, n))]
        if sizes and not modelx_b200.batch_pays_off(len(sizes), sum(sizes), max(sizes)) and max(sizes) > (256 << 20):
            print(f"note: {len(sizes)} blob(s), largest {max(sizes)/1e9:.1f} GB: reference-identical (whole-file) digests are serial "
                  "chains and hash faster on CPU cores than on the GPU below ~64 blobs; `--tree` uses the chunked identity "
                  "(whole GPU, read once)", file=sys.stderr)
    with _engine(lib_path) as eng:
        reg = LocalRegistry(store, eng)
        cl = Client(eng)
        rep = cl.push_tree(reg, ref.repository, version, directory) if tree else cl.push(reg, ref.repository, version, directory)
    _print_status(rep["blobs"], "status")
    return rep


def pull_model(raw_ref: str, into: str = "", repos=None, lib_path=None) -> list:
    """PullModelx (cmd/modelx/model/pull.go:56-69)."""
    from .client import Client, LocalRegistry
    ref = parse_reference(raw_ref, repos)
    if not ref.repository:
        raise SystemExit("repository is not specified")
    into = into or os.path.basename(ref.repository)
    version = ref.version or "latest"
    print(f"Pulling {ref} into {into} ")
    store = _store_of(ref)
    with _engine(lib_path) as eng:
        res = Client(eng).pull(LocalRegistry(store, eng), ref.repository, version, into)
    _print_status(res, "status")
    return res


def list_models(raw_ref: str, repos=None, lib_path=None) -> dict:
    """`modelx list`: the registry's index, a repository's versions, or one version's files (model/list.go)."""
    from . import _native as N
    from .client import LocalRegistry
    ref = parse_reference(raw_ref, repos)
    reg = LocalRegistry(_store_of(ref))
    if lib_path:
        reg._lib = N.load(lib_path)
    if ref.repository and ref.version:
        m = json.loads(reg.get_manifest_json(ref.repository, ref.version))
        for d in [m["config"]] + (m.get("blobs") or []):
            print(f"{d['name']:<40} {d.get('size', 0):>14} {d.get('digest', '')}")
        return m
    idx = reg.get_index(ref.repository)
    for d in idx.get("manifests") or []:
        print(f"{d['name']:<40} {d.get('size', 0):>14} {d.get('modified', '')}")
    return idx


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="modelx-b200", description="modelx client re-hosted on the B200 digest engine")
    ap.add_argument("--lib", default=None, help=argparse.SUPPRESS)     # tests: path of the CPU test double
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("init", help="init an new model at path")
    p.add_argument("path")
    p.add_argument("-f", "--force", action="store_true", help="force init")
    p = sub.add_parser("push", help="push a model to a modelx repository")
    p.add_argument("ref")
    p.add_argument("dir", nargs="?", default="")
    p.add_argument("--tree", action="store_true", help="store blobs under their modelx.tree.v1 root (read-once, GPU-wide digest; new)")
    p = sub.add_parser("pull", help="pull a model from a repository")
    p.add_argument("ref")
    p.add_argument("into", nargs="?", default="")
    p = sub.add_parser("list", help="list repositories / versions / files")
    p.add_argument("ref")
    p = sub.add_parser("repo", help="Repository management")
    rsub = p.add_subparsers(dest="repo_cmd", required=True)
    q = rsub.add_parser("add")
    q.add_argument("name")
    q.add_argument("url")
    q.add_argument("-t", "--token", default="")
    rsub.add_parser("list")
    q = rsub.add_parser("remove")
    q.add_argument("name")
    a = ap.parse_args(argv)
    try:
        if a.cmd == "init":
            init_model(a.path, a.force)
        elif a.cmd == "push":
            push_model(a.ref, a.dir, a.tree, lib_path=a.lib)
        elif a.cmd == "pull":
            pull_model(a.ref, a.into, lib_path=a.lib)
        elif a.cmd == "list":
            list_models(a.ref, lib_path=a.lib)
        elif a.cmd == "repo":
            rm = RepoManager()
            if a.repo_cmd == "add":
                rm.set(a.name, a.url, a.token)
            elif a.repo_cmd == "remove":
                rm.remove(a.name)
            else:
                for r in rm.list():
                    print(f"{r.get('name', ''):<20} {r.get('url', '')}")
    except (KeyError, ValueError, FileExistsError) as e:
        print(f"Error: {e.args[0] if e.args else e}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
