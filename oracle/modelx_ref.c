/*
 * oracle/modelx_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the modelx client call sites on the digest-and-chunk path.  Each function
 * cites the reference file:line it follows (paths relative to /root/reference).  Written from
 * the behaviour of that code; no source is copied.  See sha256_ref.c for pinning.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "oracle.h"

/* io.Copy's buffer when neither side implements ReaderFrom/WriterTo: 32 KiB (Go stdlib io.go).
 * digest.FromReader = Canonical.Digester() + io.Copy(hash, rd)  [go-digest v1.0.0]. */
#define GO_IO_COPY_BUF (32 * 1024)

/* pkg/client/push.go:160 / pull.go:116 */
int orc_digest_from_reader(int fd, uint8_t out[32], uint64_t* size) {
    orc_sha256_ctx c;
    uint8_t* buf = (uint8_t*)malloc(GO_IO_COPY_BUF);
    if (!buf) return -ENOMEM;
    orc_sha256_init(&c);
    for (;;) {
        ssize_t r = read(fd, buf, GO_IO_COPY_BUF);
        if (r < 0) {
            if (errno == EINTR) continue;
            int e = errno; free(buf); return -e;
        }
        if (r == 0) break; /* io.EOF ends io.Copy without error */
        orc_sha256_update(&c, buf, (size_t)r);
    }
    free(buf);
    orc_sha256_final(&c, out);
    if (size) *size = c.nbytes;
    return 0;
}

/* pkg/client/push.go:149-161 (the ctx-cancel goroutine has no CPU-oracle counterpart) */
int orc_client_digest(const char* path, uint8_t out[32], uint64_t* size) {
    int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return -errno;
    int rc = orc_digest_from_reader(fd, out, size);
    close(fd);
    return rc;
}

/* go-digest: Digest(alg + ":" + hex.EncodeToString(sum)); Canonical = SHA256 */
void orc_digest_string(const uint8_t d[32], char out[72]) {
    static const char hexd[] = "0123456789abcdef";
    memcpy(out, "sha256:", 7);
    for (int i = 0; i < 32; ++i) { out[7 + 2 * i] = hexd[d[i] >> 4]; out[8 + 2 * i] = hexd[d[i] & 15]; }
    out[71] = 0;
}

/* pkg/client/pull.go:115-123: 1 = "already exists" (skip), 0 = differs, <0 = -errno
 * (ENOENT means "not there, download" at pull.go:125). */
int orc_pull_file_matches(const char* path, const char* want) {
    uint8_t d[32]; char s[72];
    int rc = orc_client_digest(path, d, NULL);
    if (rc < 0) return rc;
    orc_digest_string(d, s);
    return strcmp(s, want) == 0;
}

/* pkg/client/extension_s3.go:99-112 */
int orc_calc_parts(int64_t total, int64_t partscount, orc_part* out) {
    if (partscount <= 0) return -1;            /* Go: runtime panic (divide by zero / makeslice) */
    int64_t partsize = total / partscount;     /* Go int64 division truncates toward zero */
    for (int64_t i = 0; i < partscount; ++i) {
        out[i].offset = i * partsize;
        out[i].length = (i == partscount - 1) ? total - out[i].offset : partsize;
    }
    return 0;
}

/* pkg/registry/store_s3.go:20 MultiPartUploadThreshold = 5 GiB; DefaultPartCount = 3.
 * :198-203  multipart iff forced or size > threshold, otherwise one presigned PUT (1 part);
 * :273-279  count = size / threshold, +1 if there is a remainder; 3 when that quotient is 0. */
int64_t orc_server_part_count(int64_t size, int force_multipart) {
    const int64_t threshold = 5LL * 1024 * 1024 * 1024;
    if (!force_multipart && !(size > threshold)) return 1;
    int64_t count = size / threshold;
    if (count != 0) {
        if (size % threshold != 0) count++;
        return count;
    }
    return 3;
}

/* ------------------------------------------------------------------------------------------
 * New in modelx-b200: tree digest "modelx.tree.v1" (DESIGN.md section 3), parameters
 * (leaf, fanout, chunk = leaf * fanout^k, k >= 1).
 *   level 0: leaf i = bytes [i*leaf, min((i+1)*leaf, size)), n0 = max(1, ceil(size/leaf))
 *   level j+1: node i = SHA256(concat of level-j digests i*fanout .. min((i+1)*fanout, n_j)-1)
 *   levels are built at least up to level k (the CHUNK digests, each covering `chunk` bytes) and
 *   further while a level has more than one node.
 *   root = SHA256(magic16 || LE64(size) || LE64(leaf) || LE32(fanout) || LE32(0) || top)
 * ---------------------------------------------------------------------------------------- */
static int tree_klevel(uint64_t leaf, uint32_t fanout, uint64_t chunk) {
    if (leaf == 0 || fanout < 2 || chunk < leaf) return -1;
    uint64_t span = leaf; int k = 0;
    while (span < chunk) { if (span > chunk / fanout) return -1; span *= fanout; ++k; }
    return (span == chunk && k >= 1) ? k : -1;
}

int orc_tree_shape(uint64_t size, uint64_t leaf, uint32_t fanout, uint64_t chunk, uint64_t counts[], int max_levels) {
    int k = tree_klevel(leaf, fanout, chunk);
    if (k < 0 || max_levels < 2) return -1;
    uint64_t n = size ? (size + leaf - 1) / leaf : 1;
    int lv = 0;
    counts[lv++] = n;
    while (lv <= k || n > 1) {
        if (lv >= max_levels) return -1;
        n = (n + fanout - 1) / fanout;
        counts[lv++] = n;
    }
    return lv;
}

typedef struct {
    const uint8_t* data; uint64_t n, seg, nseg; uint8_t* out;
    const orc_span* spans;
    uint64_t next; pthread_mutex_t mu; uint64_t grain;
} seg_job;

static void* seg_worker(void* arg) {
    seg_job* j = (seg_job*)arg;
    for (;;) {
        uint64_t lo = __atomic_fetch_add(&j->next, j->grain, __ATOMIC_RELAXED);
        if (lo >= j->nseg) break;
        uint64_t hi = lo + j->grain; if (hi > j->nseg) hi = j->nseg;
        for (uint64_t i = lo; i < hi; ++i) {
            if (j->spans) {
                orc_sha256(j->spans[i].ptr, j->spans[i].len, j->out + 32 * i);
            } else {
                uint64_t off = i * j->seg;
                uint64_t len = (off >= j->n) ? 0 : (j->n - off < j->seg ? j->n - off : j->seg);
                orc_sha256(j->data + off, len, j->out + 32 * i);
            }
        }
    }
    return NULL;
}

static void run_job(seg_job* j, int threads) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > j->nseg) threads = (int)(j->nseg ? j->nseg : 1);
    j->next = 0;
    j->grain = j->nseg / ((uint64_t)threads * 8) + 1;
    pthread_mutex_init(&j->mu, NULL);
    if (threads == 1) { seg_worker(j); }
    else {
        pthread_t* t = (pthread_t*)malloc(sizeof(pthread_t) * threads);
        for (int i = 0; i < threads; ++i) pthread_create(&t[i], NULL, seg_worker, j);
        for (int i = 0; i < threads; ++i) pthread_join(t[i], NULL);
        free(t);
    }
    pthread_mutex_destroy(&j->mu);
}

void orc_hash_segments(const void* data, uint64_t n, uint64_t seg, int threads, uint8_t* out) {
    seg_job j; memset(&j, 0, sizeof j);
    j.data = (const uint8_t*)data; j.n = n; j.seg = seg; j.out = out;
    j.nseg = n ? (n + seg - 1) / seg : 1;
    run_job(&j, threads);
}

void orc_sha256_batch(const orc_span* spans, uint64_t n, int threads, uint8_t* out) {
    if (n == 0) return;
    seg_job j; memset(&j, 0, sizeof j);
    j.spans = spans; j.nseg = n; j.out = out;
    run_job(&j, threads);
}

static void put_le(uint8_t* p, uint64_t v, int nbytes) { for (int i = 0; i < nbytes; ++i) p[i] = (uint8_t)(v >> (8 * i)); }

void orc_tree_root(uint64_t size, uint64_t leaf, uint32_t fanout, const uint8_t top[32], uint8_t root[32]) {
    uint8_t m[72];
    memcpy(m, "modelx.tree.v1\0\0", 16);
    put_le(m + 16, size, 8); put_le(m + 24, leaf, 8); put_le(m + 32, fanout, 4); put_le(m + 36, 0, 4);
    memcpy(m + 40, top, 32);
    orc_sha256(m, sizeof m, root);
}

int orc_tree_digest(const void* data, uint64_t size, uint64_t leaf, uint32_t fanout, uint64_t chunk, int threads,
                    uint8_t* chunk_digests, uint64_t* nchunks, uint8_t top[32], uint8_t root[32]) {
    uint64_t counts[80];
    int k = tree_klevel(leaf, fanout, chunk);
    int levels = orc_tree_shape(size, leaf, fanout, chunk, counts, 80);
    if (levels < 0) return -1;
    uint8_t* cur = (uint8_t*)malloc(32 * counts[0]);
    if (!cur) return -ENOMEM;
    orc_hash_segments(data, size, leaf, threads, cur);
    for (int lv = 1; lv < levels; ++lv) {
        uint8_t* nxt = (uint8_t*)malloc(32 * counts[lv]);
        if (!nxt) { free(cur); return -ENOMEM; }
        orc_hash_segments(cur, 32 * counts[lv - 1], 32ull * fanout, threads, nxt);
        free(cur); cur = nxt;
        if (lv == k) {
            if (chunk_digests) memcpy(chunk_digests, cur, 32 * counts[lv]);
            if (nchunks) *nchunks = counts[lv];
        }
    }
    if (top) memcpy(top, cur, 32);
    if (root) orc_tree_root(size, leaf, fanout, cur, root);
    free(cur);
    return 0;
}

/* splitmix64 (Steele, Lea, Flood 2014; public-domain constants) as a counter-based generator:
 * word j of the stream is mix(seed + j * golden) -- identical on CPU and GPU, any offset. */
static inline uint64_t splitmix64_at(uint64_t seed, uint64_t j) {
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void orc_gen_fill(void* dst, uint64_t offset, uint64_t n, uint64_t seed) {
    uint8_t* p = (uint8_t*)dst;
    uint64_t pos = offset, end = offset + n;
    while (pos < end) {
        uint64_t j = pos >> 3, w = splitmix64_at(seed, j);
        unsigned b = (unsigned)(pos & 7);
        if (b == 0 && end - pos >= 8) { memcpy(p, &w, 8); p += 8; pos += 8; continue; } /* little-endian host */
        *p++ = (uint8_t)(w >> (8 * b)); pos++;
    }
}
