/*
 * modelx_client.h -- C ABI of the host-side mirror of kubegems/modelx's pkg/client (digest path of
 * Push/Pull), pkg/types (manifest/descriptor JSON) and pkg/registry's local FS blob store.
 *
 * The reference is Go and cannot be compiled in this environment, so its host logic for the hot
 * path is restated in C++ (modelx_b200/csrc/host/client_host.cpp) on top of modelx_digest.h, with
 * the same names, argument meaning and error behaviour, and exported here so tests (ctypes) and a
 * future C++ CLI can drive it.  Everything outside the path (HTTP, S3, auth, progress bars, CLI
 * flags, tar+gzip of directory blobs) is out of scope: DESIGN.md section 8.
 *
 * Strings are NUL-terminated UTF-8.  char** outputs are malloc'ed by the library; free them with
 * mxc_free.  JSON produced here is byte-identical to Go's encoding/json of the corresponding
 * pkg/types value (field order, omitempty, HTML-safe escaping, RFC 3339 times).
 */
#ifndef MODELX_CLIENT_H
#define MODELX_CLIENT_H

#include "modelx_digest.h"

#ifdef __cplusplus
extern "C" {
#endif

/* additional status codes (same numbering space as mxd_status) */
#define MXC_ERR_DIGEST_INVALID (-20) /* errors.NewDigestInvalidError, pkg/errors/errors.go:69-71 */
#define MXC_ERR_UNSUPPORTED    (-21) /* errors.NewUnsupportedError, errors.go:61-63 (e.g. directory blobs) */
#define MXC_ERR_MANIFEST       (-22) /* errors.NewManifestInvalidError */
#define MXC_ERR_NOT_FOUND      (-23) /* ErrRegistryStoreNotFound, pkg/registry/store.go:14 */

const char* mxc_last_error(void);
void mxc_free(char* p);

/* ParseManifest (pkg/client/push.go:67-100): top-level entries of basedir, dot-files skipped,
 * configfile -> Config, directories -> tar+gz blobs, files -> file blobs, blobs sorted by name.
 * No digests yet.  -> types.Manifest JSON. */
int mxc_parse_manifest(const char* basedir, const char* configfile, char** manifest_json);

/* The digest phase of Client.Push (push.go:29-52 + pushFile :120-147): ParseManifest, then for
 * every file blob and the config: stat -> Size/Mode/Modified and the whole-file SHA-256 digest,
 * all files hashed as ONE lock-step GPU batch (the reference runs 3 goroutines).
 * flags: MXC_PUSH_TREE additionally records the modelx.tree.v1 root and parameters of each blob under
 * Descriptor.Annotations["modelx.tree.v1"]; MXC_PUSH_CACHE (new, opt-in) reuses digests
 * remembered in <basedir>/.modelx/digests.json for files whose size and mtime (ns) are unchanged, and
 * updates that file -- the reference re-hashes every blob on every push (push.go:125-131).
 * Directory blobs -> MXC_ERR_UNSUPPORTED. */
#define MXC_PUSH_TREE  1
#define MXC_PUSH_CACHE 2
int mxc_push_digest(mxd_ctx* ctx, const char* basedir, const char* configfile, int flags, char** manifest_json);

/* The check phase of Client.Pull (pull.go:41-50 + pullFile :111-127) for every blob + config of
 * the manifest: state = "already exists" (local file hashes to desc.Digest), "empty"
 * (EmptyFileDigiest: create, nothing to download), "missing" (no local file), "differs".
 * -> JSON array [{"name":..,"state":..,"digest":..}] in manifest order (blobs, then config). */
int mxc_pull_check(mxd_ctx* ctx, const char* basedir, const char* manifest_json, char** report_json);

/* pkg/registry local FS store (FSRegistryStore over LocalFSProvider). Layout:
 *   <basepath>/<repository>/blobs/sha256/<hex>        blob bytes          (store.go:56-61)
 *   <basepath>/<repository>/blobs/sha256/<hex>.meta   {"contentType","contentLength"} indented JSON (fs_local.go:155-169)
 *   <basepath>/<repository>/manifests/<reference>     types.Manifest JSON (store.go:67-69, store_fs.go:87-104)
 * verify != 0 is NEW behaviour: the stored bytes are re-hashed on the GPU and a
 * mismatch with `digest` removes the blob and returns MXC_ERR_DIGEST_INVALID; the reference stores
 * the body unverified (registry.go:144-164).  verify == 1: `digest` is the whole-file SHA-256;
 * verify == 2: `digest` is a modelx.tree.v1 root (blobs pushed by mxc_push_local_tree). */
int mxc_fs_put_blob(mxd_ctx* ctx, const char* basepath, const char* repository, const char* digest,
                    const char* content_type, const char* srcfile, int verify);
int mxc_fs_exists_blob(const char* basepath, const char* repository, const char* digest); /* 1 / 0 / <0 */
int mxc_fs_put_manifest(const char* basepath, const char* repository, const char* reference,
                        const char* content_type, const char* manifest_json);
int mxc_fs_get_manifest(const char* basepath, const char* repository, const char* reference, char** manifest_json);
int mxc_blob_digest_path(const char* repository, const char* digest, char** path); /* BlobDigestPath, store.go:56-61 */

/* Client.Push against the in-process FS store (BASELINE config 1 without the HTTP hop):
 * digest phase, then per blob PushBlob's decisions (push.go:163-194): EmptyFileDigiest -> "empty",
 * already in the store (HeadBlob) -> "exists", else PutBlob -> "done"; finally PutManifest.
 * -> JSON {"manifest":{...},"blobs":[{"name":..,"status":..}]}. */
int mxc_push_local(mxd_ctx* ctx, const char* basedir, const char* configfile, const char* basepath,
                   const char* repository, const char* version, int verify, char** report_json);
/* The read-once, tree-keyed push (NEW, not wire compatible with stock modelx clients): every blob is
 * read from disk once -- the bytes stream through the pinned ring to the GPU (modelx.tree.v1 digest) and, in the
 * same pass, into the store (mxd_tree_digest_file_tee) -- and is stored under its tree root:
 * Descriptor.Digest = "sha256:<root>", Descriptor.Annotations["modelx.digest"] =
 * "tree.v1;leaf=16384;fanout=8;chunk=8388608;chunks=<n>".  mxc_pull_check / mxc_pull_local recognise the annotation
 * and verify local files with the tree digest.  Same report shape as mxc_push_local. */
int mxc_push_local_tree(mxd_ctx* ctx, const char* basedir, const char* configfile, const char* basepath,
                        const char* repository, const char* version, char** report_json);
/* Client.Pull against the same store (pull.go:19-39, pullFile :111-143): check, then copy what is
 * missing or different out of the store with the descriptor's permission bits.
 * -> JSON array [{"name":..,"status":"already exists"|"empty"|"done"}]. */
int mxc_pull_local(mxd_ctx* ctx, const char* basepath, const char* repository, const char* version,
                   const char* into, char** report_json);

#ifdef __cplusplus
}
#endif
#endif /* MODELX_CLIENT_H */
