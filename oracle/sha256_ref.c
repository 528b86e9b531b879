/*
 * oracle/sha256_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic behind modelx's blob digest:
 *   pkg/client/push.go:160, pkg/client/pull.go:116  digest.FromReader(f)
 *   pkg/client/helper.go:46                         digest.Canonical.Digester()
 * Both resolve to github.com/opencontainers/go-digest v1.0.0 (go.mod:19, not vendored in
 * /root/reference) which delegates to the Go standard library's crypto/sha256, a conforming
 * implementation of FIPS 180-4 SHA-256.  The algorithm therefore is restated here from the
 * published standard (FIPS 180-4 sections 4.1.2, 4.2.2, 5.1.1, 5.3.3, 6.2); nothing is copied
 * from Go or from the reference tree.
 *
 * Two block functions are provided:
 *   - a portable one written line by line from the standard (the normative oracle), and
 *   - an x86 SHA-NI one (what Go >= 1.21 and OpenSSL use on this class of CPU), used only so the
 *     CPU baseline in bench.py is a fair stand-in for the Go path.  tests/ check that both agree
 *     with each other, with OpenSSL-backed hashlib, and with the FIPS known answers.
 *
 * PARITY PINNING: the reference's own tests hold no digest vector for this path.
 * The only digest constant in the reference is EmptyFileDigiest = sha256("") (push.go:25); this
 * oracle is pinned to it, to the FIPS 180-4 / NIST example vectors, and to hashlib (OpenSSL).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libmodelxdigest.so) never links or calls it.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "oracle.h"

#if defined(__x86_64__)
#include <immintrin.h>
#include <cpuid.h>
#endif

/* FIPS 180-4 section 4.2.2: first 32 bits of the fractional parts of the cube roots of the
 * first 64 primes. */
static const uint32_t K256[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

/* FIPS 180-4 section 5.3.3: initial hash value. */
static const uint32_t H256_INIT[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au,
                                      0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};

static inline uint32_t rotr32(uint32_t x, unsigned n) { return (x >> n) | (x << (32u - n)); }
static inline uint32_t load_be32(const uint8_t* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}
static inline void store_be32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}

/* FIPS 180-4 section 6.2.2, steps 1-4, for `nblocks` consecutive 512-bit blocks. */
static void compress_portable(uint32_t H[8], const uint8_t* msg, size_t nblocks) {
    uint32_t W[64];
    while (nblocks--) {
        for (int t = 0; t < 16; ++t) W[t] = load_be32(msg + 4 * t);
        for (int t = 16; t < 64; ++t) {
            uint32_t s0 = rotr32(W[t - 15], 7) ^ rotr32(W[t - 15], 18) ^ (W[t - 15] >> 3);   /* sigma0, eq 4.6 */
            uint32_t s1 = rotr32(W[t - 2], 17) ^ rotr32(W[t - 2], 19) ^ (W[t - 2] >> 10);    /* sigma1, eq 4.7 */
            W[t] = s1 + W[t - 7] + s0 + W[t - 16];
        }
        uint32_t a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5], g = H[6], h = H[7];
        for (int t = 0; t < 64; ++t) {
            uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);   /* Sigma1, eq 4.5 */
            uint32_t ch = (e & f) ^ (~e & g);                             /* Ch,     eq 4.2 */
            uint32_t T1 = h + S1 + ch + K256[t] + W[t];
            uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);   /* Sigma0, eq 4.4 */
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);                    /* Maj,    eq 4.3 */
            uint32_t T2 = S0 + mj;
            h = g; g = f; f = e; e = d + T1; d = c; c = b; b = a; a = T1 + T2;
        }
        H[0] += a; H[1] += b; H[2] += c; H[3] += d; H[4] += e; H[5] += f; H[6] += g; H[7] += h;
        msg += 64;
    }
}

#if defined(__x86_64__)
/* Same function on the x86 SHA extensions: sha256rnds2 performs two rounds on the
 * (ABEF, CDGH) state split, sha256msg1/msg2 compute the sigma0/sigma1 parts of the schedule. */
__attribute__((target("sha,sse4.1,ssse3")))
static void compress_shani(uint32_t H[8], const uint8_t* msg, size_t nblocks) {
    const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i t0 = _mm_loadu_si128((const __m128i*)&H[0]);        /* a b c d (lane0 = a) */
    __m128i t1 = _mm_loadu_si128((const __m128i*)&H[4]);        /* e f g h */
    t0 = _mm_shuffle_epi32(t0, 0xB1);                           /* b a d c */
    t1 = _mm_shuffle_epi32(t1, 0x1B);                           /* h g f e */
    __m128i abef = _mm_alignr_epi8(t0, t1, 8);                  /* f e b a  */
    __m128i cdgh = _mm_blend_epi16(t1, t0, 0xF0);               /* h g d c  */
    while (nblocks--) {
        const __m128i abef0 = abef, cdgh0 = cdgh;
        __m128i m[4], x;
        for (int i = 0; i < 4; ++i)
            m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(msg + 16 * i)), bswap);
        for (int r = 0; r < 16; ++r) {                          /* 16 groups of 4 rounds */
            __m128i cur = m[r & 3];
            x = _mm_add_epi32(cur, _mm_loadu_si128((const __m128i*)&K256[4 * r]));
            cdgh = _mm_sha256rnds2_epu32(cdgh, abef, x);
            abef = _mm_sha256rnds2_epu32(abef, cdgh, _mm_shuffle_epi32(x, 0x0E));
            if (r < 12) {                                       /* W[4r+16 .. 4r+19] replaces m[r&3] */
                __m128i w = _mm_sha256msg1_epu32(cur, m[(r + 1) & 3]);
                w = _mm_add_epi32(w, _mm_alignr_epi8(m[(r + 3) & 3], m[(r + 2) & 3], 4));
                m[r & 3] = _mm_sha256msg2_epu32(w, m[(r + 3) & 3]);
            }
        }
        abef = _mm_add_epi32(abef, abef0);
        cdgh = _mm_add_epi32(cdgh, cdgh0);
        msg += 64;
    }
    t0 = _mm_shuffle_epi32(abef, 0x1B);                         /* a b e f */
    t1 = _mm_shuffle_epi32(cdgh, 0xB1);                         /* g h c d -> after shuffle: c d g h? see below */
    _mm_storeu_si128((__m128i*)&H[0], _mm_blend_epi16(t0, t1, 0xF0));
    _mm_storeu_si128((__m128i*)&H[4], _mm_alignr_epi8(t1, t0, 8));
}

static int cpu_has_shani(void) {
    unsigned a, b, c, d;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return 0;
    if (!(b & (1u << 29))) return 0;                            /* CPUID.7.0:EBX.SHA */
    if (!__get_cpuid(1, &a, &b, &c, &d)) return 0;
    return (c & (1u << 19)) && (c & (1u << 9));                 /* SSE4.1, SSSE3 */
}
#endif

static int g_engine = -1; /* -1 auto, 0 portable, 1 sha-ni */

int orc_sha256_set_engine(int engine) {
#if defined(__x86_64__)
    if (engine == 1 && !cpu_has_shani()) return -1;
#else
    if (engine == 1) return -1;
#endif
    g_engine = engine;
    return 0;
}

int orc_sha256_engine(void) {
    if (g_engine >= 0) return g_engine;
#if defined(__x86_64__)
    return cpu_has_shani() ? 1 : 0;
#else
    return 0;
#endif
}

static void compress(uint32_t H[8], const uint8_t* msg, size_t nblocks) {
#if defined(__x86_64__)
    if (orc_sha256_engine() == 1) { compress_shani(H, msg, nblocks); return; }
#endif
    compress_portable(H, msg, nblocks);
}

void orc_sha256_init(orc_sha256_ctx* c) {
    memcpy(c->h, H256_INIT, sizeof H256_INIT);
    c->nbytes = 0;
    c->nbuf = 0;
}

void orc_sha256_update(orc_sha256_ctx* c, const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    c->nbytes += n;
    if (c->nbuf) {
        size_t take = 64 - c->nbuf;
        if (take > n) take = n;
        memcpy(c->buf + c->nbuf, p, take);
        c->nbuf += (uint32_t)take; p += take; n -= take;
        if (c->nbuf < 64) return;
        compress(c->h, c->buf, 1);
        c->nbuf = 0;
    }
    if (n >= 64) {
        size_t nb = n / 64;
        compress(c->h, p, nb);
        p += nb * 64; n -= nb * 64;
    }
    if (n) { memcpy(c->buf, p, n); c->nbuf = (uint32_t)n; }
}

/* FIPS 180-4 section 5.1.1 padding: 0x80, zeros to 56 mod 64, then the 64-bit big-endian bit
 * length.  Does not modify *c (mirrors Go's hash.Hash.Sum, which works on a copy). */
void orc_sha256_final(const orc_sha256_ctx* c, uint8_t out[32]) {
    uint32_t h[8];
    uint8_t tail[128];
    memcpy(h, c->h, sizeof h);
    size_t n = c->nbuf;
    memcpy(tail, c->buf, n);
    tail[n++] = 0x80;
    size_t padded = (n <= 56) ? 64 : 128;
    memset(tail + n, 0, padded - n);
    uint64_t bits = c->nbytes * 8u;
    for (int i = 0; i < 8; ++i) tail[padded - 1 - i] = (uint8_t)(bits >> (8 * i));
    compress(h, tail, padded / 64);
    for (int i = 0; i < 8; ++i) store_be32(out + 4 * i, h[i]);
}

void orc_sha256(const void* data, size_t n, uint8_t out[32]) {
    orc_sha256_ctx c;
    orc_sha256_init(&c);
    orc_sha256_update(&c, data, n);
    orc_sha256_final(&c, out);
}

/* Raw block function on caller-held state (used to check the GPU's carried chain state). */
void orc_sha256_blocks(uint32_t h[8], const void* data, size_t nblocks) { compress(h, (const uint8_t*)data, nblocks); }
void orc_sha256_iv(uint32_t h[8]) { memcpy(h, H256_INIT, sizeof H256_INIT); }
