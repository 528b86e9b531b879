/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of modelx's digest-and-chunk path, used only as the parity checker
 * (tests/, __graft_entry__.smoke()) and as the timed CPU baseline (bench.py).  See the header
 * of sha256_ref.c for provenance and pinning.  Nothing under modelx_b200/ includes this file.
 */
#ifndef MODELX_ORACLE_H
#define MODELX_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t h[8];
    uint64_t nbytes;
    uint32_t nbuf;
    uint8_t buf[64];
} orc_sha256_ctx;

/* --- FIPS 180-4 SHA-256 (sha256_ref.c) ------------------------------------------------ */
void orc_sha256_init(orc_sha256_ctx* c);
void orc_sha256_update(orc_sha256_ctx* c, const void* data, size_t n);
void orc_sha256_final(const orc_sha256_ctx* c, uint8_t out[32]);
void orc_sha256(const void* data, size_t n, uint8_t out[32]);
void orc_sha256_blocks(uint32_t h[8], const void* data, size_t nblocks);
void orc_sha256_iv(uint32_t h[8]);
int orc_sha256_set_engine(int engine); /* -1 auto, 0 portable (normative), 1 SHA-NI */
int orc_sha256_engine(void);

/* --- the reference's call sites (modelx_ref.c) ---------------------------------------- */
/* digest.FromReader(f) as called at pkg/client/push.go:160 and pull.go:116: io.Copy into the
 * hash with a 32 KiB buffer, i.e. read(2) loop until EOF.  Returns 0 or -errno. */
int orc_digest_from_reader(int fd, uint8_t out[32], uint64_t* size);
/* Client.digest, push.go:149-161: open + FromReader + close. */
int orc_client_digest(const char* path, uint8_t out[32], uint64_t* size);
/* digest.Digest string form: "sha256:" + 64 lower-case hex (71 chars + NUL). */
void orc_digest_string(const uint8_t d[32], char out[72]);
/* pull.go:120: string equality of the computed digest and desc.Digest. */
int orc_pull_file_matches(const char* path, const char* want_digest_string);

typedef struct { int64_t offset, length; } orc_part;
/* calcParts, pkg/client/extension_s3.go:99-112.  Go panics (integer divide by zero) for
 * partscount == 0 and (make with negative len) for < 0; reported here as -1. */
int orc_calc_parts(int64_t total, int64_t partscount, orc_part* out);
/* Server-side part count: pkg/registry/store_s3.go:198-203 (threshold test) + :273-279. */
int64_t orc_server_part_count(int64_t size, int force_multipart);

/* --- definitions that are NEW in modelx-b200 (no reference counterpart) ---------------
 * Spec restated on the CPU so the GPU implementation has an independent check.  Every node of
 * the tree is a plain SHA-256 of well-defined bytes. */
/* number of nodes per level; returns the number of levels (>= k+1 where chunk = leaf*fanout^k). */
int orc_tree_shape(uint64_t size, uint64_t leaf, uint32_t fanout, uint64_t chunk, uint64_t counts[], int max_levels);
/* chunk_digests: level-k nodes (each covers `chunk` bytes), may be NULL.
 * top: digest of the single node of the last level.  root: final blob identity. */
int orc_tree_digest(const void* data, uint64_t size, uint64_t leaf, uint32_t fanout, uint64_t chunk, int threads,
                    uint8_t* chunk_digests, uint64_t* nchunks, uint8_t top[32], uint8_t root[32]);
void orc_tree_root(uint64_t size, uint64_t leaf, uint32_t fanout, const uint8_t top[32], uint8_t root[32]);
/* one level of the tree: out[j] = SHA256(in[j*seg .. min((j+1)*seg, n))), threaded. */
void orc_hash_segments(const void* data, uint64_t n, uint64_t seg, int threads, uint8_t* out);
/* batch of independent messages (n spans), threaded with `threads` workers. */
typedef struct { const void* ptr; uint64_t len; } orc_span;
void orc_sha256_batch(const orc_span* spans, uint64_t n, int threads, uint8_t* out);

/* deterministic synthetic blob: 64-bit little-endian word j = splitmix64(seed + j).
 * Fills bytes [offset, offset+n) of that stream (any alignment). */
void orc_gen_fill(void* dst, uint64_t offset, uint64_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
