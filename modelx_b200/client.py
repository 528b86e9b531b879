"""Python face of the C++ host mirror (include/modelx_client.h): the digest path of
``pkg/client.Client.{Push,Pull}``, ``pkg/types`` JSON and the local FS store of ``pkg/registry``.
All logic lives in modelx_b200/csrc/host/client_host.cpp; this module only marshals arguments."""
from __future__ import annotations

import ctypes as C
import json
from typing import Optional

from . import _native as N
from .engine import Engine


def _take(lib, p: C.c_void_p) -> str:
    try:
        return C.cast(p, C.c_char_p).value.decode()
    finally:
        lib.mxc_free(p)


def parse_manifest_json(basedir: str, configfile: str = "modelx.yaml") -> str:
    """ParseManifest (pkg/client/push.go:67-100) -> the exact JSON Go's encoding/json would emit."""
    lib = N.load()
    out = C.c_void_p()
    N.check(lib.mxc_parse_manifest(basedir.encode(), configfile.encode(), C.byref(out)), "mxc_parse_manifest")
    return _take(lib, out)


def blob_digest_path(repository: str, digest: str) -> str:
    lib = N.load()
    out = C.c_void_p()
    N.check(lib.mxc_blob_digest_path(repository.encode(), digest.encode(), C.byref(out)), "mxc_blob_digest_path")
    return _take(lib, out)


class LocalRegistry:
    """FSRegistryStore over LocalFSProvider (pkg/registry/store_fs.go, fs_local.go): files on disk only."""

    def __init__(self, basepath: str, engine: Optional[Engine] = None):
        self.basepath = basepath
        self.engine = engine
        self._lib = N.load()

    def put_blob(self, repository: str, digest: str, srcfile: str, content_type: str = "application/octet-stream",
                 verify=False) -> None:
        """verify: False/0 (reference behaviour: store unverified), True/1 (whole-file digest), "tree"/2 (tree root)."""
        ctx = self.engine.handle if self.engine else None
        mode = 2 if verify in ("tree", 2) else (1 if verify else 0)
        N.check(self._lib.mxc_fs_put_blob(ctx, self.basepath.encode(), repository.encode(), digest.encode(),
                                          content_type.encode(), srcfile.encode(), mode), "mxc_fs_put_blob")

    def exists_blob(self, repository: str, digest: str) -> bool:
        rc = self._lib.mxc_fs_exists_blob(self.basepath.encode(), repository.encode(), digest.encode())
        if rc < 0:
            raise N.MxdError(rc, "mxc_fs_exists_blob")
        return bool(rc)

    def put_manifest(self, repository: str, reference: str, manifest_json: str,
                     content_type: str = "application/vnd.modelx.model.manifest.v1.json") -> None:
        N.check(self._lib.mxc_fs_put_manifest(self.basepath.encode(), repository.encode(), reference.encode(),
                                              content_type.encode(), manifest_json.encode()), "mxc_fs_put_manifest")

    def get_manifest_json(self, repository: str, reference: str) -> str:
        out = C.c_void_p()
        N.check(self._lib.mxc_fs_get_manifest(self.basepath.encode(), repository.encode(), reference.encode(),
                                              C.byref(out)), "mxc_fs_get_manifest")
        return _take(self._lib, out)


class Client:
    """The digest path of pkg/client.Client (push.go / pull.go) on a GPU engine."""

    def __init__(self, engine: Engine):
        self.engine = engine
        self._lib = N.load()

    def push_digest_json(self, basedir: str, configfile: str = "modelx.yaml", with_tree: bool = False,
                         use_cache: bool = False) -> str:
        out = C.c_void_p()
        flags = (1 if with_tree else 0) | (2 if use_cache else 0)
        N.check(self._lib.mxc_push_digest(self.engine.handle, basedir.encode(), configfile.encode(),
                                          flags, C.byref(out)), "mxc_push_digest")
        return _take(self._lib, out)

    def pull_check(self, basedir: str, manifest_json: str) -> list:
        out = C.c_void_p()
        N.check(self._lib.mxc_pull_check(self.engine.handle, basedir.encode(), manifest_json.encode(), C.byref(out)),
                "mxc_pull_check")
        return json.loads(_take(self._lib, out))

    def push(self, registry: LocalRegistry, repository: str, version: str, basedir: str,
             configfile: str = "modelx.yaml", verify: bool = False) -> dict:
        """Client.Push (push.go:29-65) against the in-process FS store."""
        out = C.c_void_p()
        N.check(self._lib.mxc_push_local(self.engine.handle, basedir.encode(), configfile.encode(),
                                         registry.basepath.encode(), repository.encode(), version.encode(),
                                         1 if verify else 0, C.byref(out)), "mxc_push_local")
        text = _take(self._lib, out)
        rep = json.loads(text)
        rep["manifest_json"] = text[len('{"manifest":'):text.rindex(',"blobs":[')]
        return rep

    def push_tree(self, registry: LocalRegistry, repository: str, version: str, basedir: str,
                  configfile: str = "modelx.yaml") -> dict:
        """Read-once, tree-keyed push: each blob streams once through the pinned ring to the GPU
        (modelx.tree.v1) and into the store, and is stored under its tree root."""
        out = C.c_void_p()
        N.check(self._lib.mxc_push_local_tree(self.engine.handle, basedir.encode(), configfile.encode(),
                                              registry.basepath.encode(), repository.encode(), version.encode(),
                                              C.byref(out)), "mxc_push_local_tree")
        text = _take(self._lib, out)
        rep = json.loads(text)
        rep["manifest_json"] = text[len('{"manifest":'):text.rindex(',"blobs":[')]
        return rep

    def pull(self, registry: LocalRegistry, repository: str, version: str, into: str) -> list:
        """Client.Pull (pull.go:19-39) against the in-process FS store."""
        out = C.c_void_p()
        N.check(self._lib.mxc_pull_local(self.engine.handle, registry.basepath.encode(), repository.encode(),
                                         version.encode(), into.encode(), C.byref(out)), "mxc_pull_local")
        return json.loads(_take(self._lib, out))
