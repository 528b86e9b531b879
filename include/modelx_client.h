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
 * Directory blobs are packed to <basedir>/.modelx/<name>.tar.gz first (pushDirectory, push.go:102-118). */
#define MXC_PUSH_TREE  1
#define MXC_PUSH_CACHE 2
#define MXC_PUSH_FORCE_MULTIPART 4   /* part counts as if the server forced multipart (store_s3.go:273-279: 3 parts below 5 GiB); tests */
int mxc_push_digest(mxd_ctx* ctx, const char* basedir, const char* configfile, int flags, char** manifest_json);

/* Directory blobs (pkg/client/helper.go:24-83).  mxc_tgz = TGZ(ctx, dir, intofile): tar+gzip of the directory's
 * contents with owner and times cleared (archiver's ClearAttributes), written to `intofile` when given, and the
 * SHA-256 of the archive bytes taken while they are produced (io.MultiWriter into the digester, helper.go:46-50;
 * here the GPU hasher).  mxc_untgz = UnTGZ: extracts into `intodir`, refusing entries that would leave it.
 * The compressed bytes differ from the Go compressor's, so a directory digest is stable across runs of this client
 * but is not the one a stock modelx client computes for the same directory. */
int mxc_tgz(mxd_ctx* ctx, const char* dir, const char* intofile /*may be NULL*/, char** digest, uint64_t* archive_size);
int mxc_untgz(const char* archive, const char* intodir);

/* The check phase of Client.Pull (pull.go:41-50 + pullFile :111-127) for every blob + config of
 * the manifest: state = "already exists" (local file hashes to desc.Digest), "empty"
 * (EmptyFileDigiest: create, nothing to download), "missing" (no local file), "differs".
 * -> JSON array [{"name":..,"state":..,"digest":..}] in manifest order (blobs, then config). */
int mxc_pull_check(mxd_ctx* ctx, const char* basedir, const char* manifest_json, char** report_json);

/* pkg/registry local FS store (FSRegistryStore over LocalFSProvider). Layout:
 *   <basepath>/<repository>/blobs/sha256/<hex>        blob bytes          (store.go:56-61)
 *   <basepath>/<repository>/blobs/sha256/<hex>.meta   {"contentType","contentLength"} indented JSON (fs_local.go:155-169)
 *   <basepath>/<repository>/manifests/<reference>     types.Manifest JSON (store.go:67-69, store_fs.go:87-104)
 *   <basepath>/<repository>/index.json, <basepath>/index.json   types.Index (store_fs.go:145-172,262-330)
 * verify != 0 is NEW behaviour: the body is hashed on the GPU WHILE it is written to a temp file (one
 * read, teed) and only a body whose digest equals `digest` is renamed into place; a mismatch returns
 * MXC_ERR_DIGEST_INVALID and leaves the store untouched -- an existing blob is never overwritten or deleted.  The
 * reference stores the body unverified (registry.go:144-164).  verify == 1: `digest` is the whole-file SHA-256;
 * verify == 2: `digest` is a modelx.tree.v1 root (blobs pushed by mxc_push_local_tree). */
int mxc_fs_put_blob(mxd_ctx* ctx, const char* basepath, const char* repository, const char* digest,
                    const char* content_type, const char* srcfile, int verify);
int mxc_fs_exists_blob(const char* basepath, const char* repository, const char* digest); /* 1 / 0 / <0 */
int mxc_fs_put_manifest(const char* basepath, const char* repository, const char* reference,
                        const char* content_type, const char* manifest_json);
int mxc_fs_get_manifest(const char* basepath, const char* repository, const char* reference, char** manifest_json);
/* <basepath>/<repository>/index.json (types.Index: one descriptor per pushed version) or, with repository NULL/"",
 * the registry-wide <basepath>/index.json.  Both are rewritten by every mxc_fs_put_manifest, as FSRegistryStore.
 * PutManifest -> RefreshIndex -> RefreshGlobalIndex do (store_fs.go:87-104,185-238,287-330). */
int mxc_fs_get_index(const char* basepath, const char* repository, char** index_json);
int mxc_blob_digest_path(const char* repository, const char* digest, char** path); /* BlobDigestPath, store.go:56-61 */

/* Client.Push against the in-process FS store (BASELINE config 1 without the HTTP hop):
 * every blob is read from disk once (mxc_push_stream below: digest and store copy in the same pass), then PushBlob's
 * decisions (push.go:163-194) are applied with the digest in hand: EmptyFileDigiest -> "empty", already in the store
 * (HeadBlob) -> "exists", else the copy is renamed into place -> "done"; finally PutManifest (+ index.json).
 * flags: MXC_PUSH_FORCE_MULTIPART.  -> JSON as mxc_push_stream. */
int mxc_push_local(mxd_ctx* ctx, const char* basedir, const char* configfile, const char* basepath,
                   const char* repository, const char* version, int flags, char** report_json);

/* The part-upload consumer (S3Extension.Upload's role, extension_s3.go:52-89) for the read-once push: each blob is
 * read from disk ONCE; in the same GPU rounds it yields the blob's SHA-256 and the SHA-256 of every multipart part
 * (calcParts over the server's part count), while its bytes are teed, part by part, to these callbacks.
 *   begin         a blob is about to stream: its size and part ranges (1 part below the 5 GiB threshold)
 *   part_write    bytes [offset, offset+n) of the blob, all inside part `part`; called from several threads, in any
 *                 order, at most max_concurrent at a time when that is > 0 (the reference sends 3 parts at a time,
 *                 extension_s3.go:18,66).  Non-zero = this part failed: after the pass the part is re-read from the
 *                 file and re-sent, 3 attempts in all (retry(ctx, 3, ...), extension_s3.go:66-84,133-148), each
 *                 announced by part_restart (may be NULL)
 *   complete      every byte delivered: the content address "sha256:<hex>" and nparts*32 bytes of part digests
 *                 (e.g. x-amz-checksum-sha256); writes "done" / "exists" / "empty" to status
 *   abort         the push failed; drop what was received (may be NULL)
 * Return 0 for success.  mxc_push_local is mxc_push_stream with the local FS store as the uploader. */
typedef struct mxc_uploader {
    void* user;
    int max_concurrent;
    int (*begin)(void* user, uint64_t blob, const char* name, uint64_t size, const mxd_part* parts, uint64_t nparts);
    int (*part_write)(void* user, uint64_t blob, uint64_t part, uint64_t offset, const void* data, uint64_t n);
    int (*part_restart)(void* user, uint64_t blob, uint64_t part);
    int (*complete)(void* user, uint64_t blob, const char* digest, const uint8_t* part_sha256, uint64_t nparts, char status[16]);
    void (*abort)(void* user, uint64_t blob);
} mxc_uploader;
/* -> JSON {"manifest":{...},"blobs":[{"name","status","digest","size","parts":[{"offset","length","sha256"}]}],"reread_bytes":N} */
int mxc_push_stream(mxd_ctx* ctx, const char* basedir, const char* configfile, const mxc_uploader* up, int flags, char** report_json);
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
