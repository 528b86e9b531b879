// Multi-stream SHA-256 for sm_100a: one independent message per lane, a warp hashes 32
// messages in lock step.  This is the kernel behind every digest modelx-b200 produces:
//   - leaf / chunk / upper tree levels of a blob (uniform segments of one buffer),
//   - batches of whole blobs (arbitrary spans; the reference's one-digest-per-file semantics,
//     pkg/client/push.go:149-161 and pull.go:115-123, across many files at once),
//   - chained segments of a single stream (hash.Hash-shaped incremental API, helper.go:46).
// Pure 32-bit integer work, no tensor cores.  In the throughput kernel (k_sha256_lanes) the 16-word schedule
// and the chain state live in registers and round constants are instruction immediates (no shared memory);
// the latency kernel for few long chains (k_sha256_chains_coop) passes W+K between two warps through 16 KB
// of shared memory.
#include "kernels.h"
#include "sha256_device.cuh"
#include <cstdlib>

namespace mxd {

namespace {

// 64-thread CTAs: measured identical to 128 threads at 12.5, 20 and 100 GB (profiles/r01_quick_bench_variants.txt);
// kept for the finer work quantum per CTA (1 MiB of leaves).
constexpr int kThreads = 64;

__device__ __forceinline__ uint4 ldg128(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// 64 message bytes at p (any alignment) -> 16 big-endian words.
// a = p & 3; q = p - a is 4-byte aligned.  Little-endian word k of the block is
// funnelshift_r(q[k], q[k+1], 8a).  q[16] is only touched when a != 0, in which case it holds
// the block's last byte(s), so it is inside the message.
__device__ __forceinline__ void load_block_unaligned(const uint8_t* p, uint32_t (&w)[16]) {
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p - a);
    if (a == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) w[k] = bswap32(__ldg(q + k));
    } else {
        uint32_t lo = __ldg(q);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t hi = __ldg(q + k + 1);
            w[k] = bswap32(__funnelshift_r(lo, hi, 8 * a));
            lo = hi;
        }
    }
}

__device__ __forceinline__ void unpack_block(const uint4& v0, const uint4& v1, const uint4& v2, const uint4& v3,
                                             uint32_t (&w)[16]) {
    w[0] = bswap32(v0.x);  w[1] = bswap32(v0.y);  w[2] = bswap32(v0.z);  w[3] = bswap32(v0.w);
    w[4] = bswap32(v1.x);  w[5] = bswap32(v1.y);  w[6] = bswap32(v1.z);  w[7] = bswap32(v1.w);
    w[8] = bswap32(v2.x);  w[9] = bswap32(v2.y);  w[10] = bswap32(v2.z); w[11] = bswap32(v2.w);
    w[12] = bswap32(v3.x); w[13] = bswap32(v3.y); w[14] = bswap32(v3.z); w[15] = bswap32(v3.w);
}

// MINB = resident CTAs per SM the register allocator must allow (12 x 64 threads -> 80 registers, 16 -> 63).
template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_sha256_lanes(const MsgJob j) {
    const uint64_t m = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = m < j.nmsg;
    const uint32_t one = j.one;

    // ---- locate this lane's message ------------------------------------------------------
    const uint8_t* ptr = nullptr;
    uint64_t len = 0;
    if (valid) {
        if (j.base != nullptr) {
            const uint64_t off = m * j.seg;
            ptr = j.base + off;
            len = (off < j.nbytes) ? ((j.nbytes - off < j.seg) ? j.nbytes - off : j.seg) : 0;
        } else {
            const DevSpan sp = reinterpret_cast<const DevSpan*>(j.spans)[m];
            ptr = static_cast<const uint8_t*>(sp.ptr);
            len = sp.len;
        }
    }
    uint32_t h[8];
    if (valid && j.state != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = j.state[8 * m + i];
    } else {
        sha256_iv(h);
    }
    const uint64_t prefix = (j.prefix != nullptr && valid) ? j.prefix[m] : j.prefix_all;
    const uint64_t nfull = len >> 6;
    const uint32_t r = (uint32_t)(len & 63u);
    // blocks this lane compresses: the full ones, then (when finalizing) the padded tail block
    // and, if the 64-bit length does not fit behind the tail, one more (FIPS 180-4 section 5.1.1).
    int fin = j.finalize;
    bool live = valid;
    if (j.ctl != nullptr && valid) { const uint8_t c = j.ctl[m]; fin = (c == 1); live = (c != 2); }
    const uint64_t nblk = live ? nfull + (fin ? (r >= 56 ? 2u : 1u) : 0u) : 0;
    const uint64_t bits = (prefix + len) << 3;

    // Hot loop: when every lane of the warp has a 16-byte aligned message (always true for tree
    // levels) the full blocks run in a straight-line loop of their own: 4 x LDG.128 per block, the
    // next block prefetched into registers while this one is compressed.  Keeping this loop free of
    // control-flow merges matters: after a merge ptxas must wait for every load that any incoming
    // path may have in flight, which would serialise the prefetch with the compress (7.7 % of all
    // stall samples in the first profile, profiles/r01_ncu_leaf_v1_summary.txt).
    // Lane i streams its own message, so a request touches 32 different lines but consumes whole
    // 32-byte sectors: DRAM traffic equals the algorithmic bytes.
    const bool warp_aligned = __all_sync(0xffffffffu, (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0);
    uint32_t w[16];
    uint64_t b = 0;
    if (warp_aligned && live && nfull) {
        const uint4* p4 = reinterpret_cast<const uint4*>(ptr);
        uint4 v0 = ldg128(p4), v1 = ldg128(p4 + 1), v2 = ldg128(p4 + 2), v3 = ldg128(p4 + 3);
        for (; b < nfull; ++b) {
            unpack_block(v0, v1, v2, v3, w);
            p4 += (b + 1 < nfull) ? 4 : 0;   // last iteration re-reads its own block (L1 hit) instead of branching
            v0 = ldg128(p4); v1 = ldg128(p4 + 1); v2 = ldg128(p4 + 2); v3 = ldg128(p4 + 3);
            sha256_compress(h, w, one);
        }
    }
    // Everything else: unaligned full blocks, the padded tail block and the length block.
    for (; b < nblk; ++b) {
        if (b < nfull) {
            load_block_unaligned(ptr + (b << 6), w);
        } else if (b == nfull) {
            // r tail bytes, the 0x80 marker, zeros; the length too when it fits (r < 56)
            const uint8_t* t = ptr + (nfull << 6);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                uint32_t word = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t idx = 4 * k + q;
                    if (idx < r) word |= (uint32_t)__ldg(t + idx) << (24 - 8 * q);
                    else if (idx == r) word |= 0x80u << (24 - 8 * q);
                }
                w[k] = word;
            }
            if (r < 56) { w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits; }
        } else {
#pragma unroll
            for (int k = 0; k < 14; ++k) w[k] = 0;
            w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits;
        }
        sha256_compress(h, w, one);
    }

    if (!live) return;
    if (!fin) {
#pragma unroll
        for (int i = 0; i < 8; ++i) j.state[8 * m + i] = h[i];
        return;
    }
    uint4 lo, hi;
    lo.x = bswap32(h[0]); lo.y = bswap32(h[1]); lo.z = bswap32(h[2]); lo.w = bswap32(h[3]);
    hi.x = bswap32(h[4]); hi.y = bswap32(h[5]); hi.z = bswap32(h[6]); hi.w = bswap32(h[7]);
    uint4* o = reinterpret_cast<uint4*>(j.out + 32 * m);
    o[0] = lo; o[1] = hi;
}



// =====================================================================================================
// k_sha256_chains_coop: few, long chains.  A SHA-256 chain is serial, so when a launch has only a few
// thousand messages (a push/pull of a few hundred files, one streamed file, a ring slot) the lanes
// kernel above is latency bound: a lone warp needs ~3,300 clk per block because the 480-instruction
// message schedule and the loads sit in the same instruction stream as the 64 dependent rounds.
// Here every 32 chains get two warps on two different SM sub-partitions:
//   warp 1 (producer)  loads/pads block b of its 32 messages (next block prefetched into registers), expands
//                      the schedule and stores W[t]+K[t] (t = 0..63) to shared memory, one stage ahead;
//   warp 0 (chain)     runs only the 64 rounds, reading W[t]+K[t] with LDS.128 (16 per block), with the
//                      round written so that a single addition follows Sigma1 on the e-chain.
// Measured 69-73 MB/s per chain against 38 MB/s in the lanes kernel (profiles/r01_batch_bench_coop_v2.txt).  Same MsgJob contract as the lanes kernel
// (spans or segments, chained state, per-message control bytes), same results bit for bit.
// =====================================================================================================
constexpr int kCoopStages = 2;
#ifndef MXD_COOP_SHORT_CHAIN
#define MXD_COOP_SHORT_CHAIN 1
#endif

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

__global__ void __launch_bounds__(64) k_sha256_chains_coop(const MsgJob j) {
    // wk[stage][t/4][lane] = {W+K for rounds 4g..4g+3} of that lane's current block
    __shared__ uint4 wk[kCoopStages][16][32];
    const int lane = threadIdx.x & 31;
    const int role = threadIdx.x >> 5;            // 0 = chain warp, 1 = producer warp
    const uint64_t m = (uint64_t)blockIdx.x * 32 + lane;
    const bool valid = m < j.nmsg;
    const uint32_t one = j.one;

    const uint8_t* ptr = nullptr;
    uint64_t len = 0;
    if (valid) {
        if (j.base != nullptr) {
            const uint64_t off = m * j.seg;
            ptr = j.base + off;
            len = (off < j.nbytes) ? ((j.nbytes - off < j.seg) ? j.nbytes - off : j.seg) : 0;
        } else {
            const DevSpan sp = reinterpret_cast<const DevSpan*>(j.spans)[m];
            ptr = static_cast<const uint8_t*>(sp.ptr);
            len = sp.len;
        }
    }
    const uint64_t prefix = (j.prefix != nullptr && valid) ? j.prefix[m] : j.prefix_all;
    const uint64_t nfull = len >> 6;
    const uint32_t r = (uint32_t)(len & 63u);
    int fin = j.finalize;
    bool live = valid;
    if (j.ctl != nullptr && valid) { const uint8_t c = j.ctl[m]; fin = (c == 1); live = (c != 2); }
    const uint64_t nblk = live ? nfull + (fin ? (r >= 56 ? 2u : 1u) : 0u) : 0;
    const uint64_t bits = (prefix + len) << 3;
    // both warps iterate to the longest chain of the 32; shorter lanes idle through the barriers
    uint64_t nmax = nblk;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const uint64_t other = __shfl_xor_sync(0xffffffffu, nmax, o); nmax = other > nmax ? other : nmax; }

    constexpr int kFull0 = 1, kEmpty0 = 1 + kCoopStages;   // named barrier ids (0 is __syncthreads)

    if (role == 1) {
        // ---------------- producer: load / pad, expand, publish W+K ---------------------------------
        // expands w[16] to the 64 schedule words, adds the round constants and publishes them for `lane`
        auto expand_store = [&](uint32_t (&w)[16], int st) {
            constexpr K256Table K = k256_table();
#pragma unroll
            for (int g = 0; g < 16; ++g) {
                uint32_t o4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = 4 * g + q;
                    if (t >= 16) {
                        uint32_t x = add_fma(w[t & 15], small_sigma0(w[(t + 1) & 15]), one);
                        x = add_fma(x, w[(t + 9) & 15], one);
                        w[t & 15] = add_fma(x, small_sigma1(w[(t + 14) & 15]), one);
                    }
                    o4[q] = add_fma(w[t & 15], K.v[t], one);
                }
                wk[st][g][lane] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            }
        };
        const bool warp_aligned = __all_sync(0xffffffffu, (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0);
        uint32_t w[16];
        uint64_t b = 0;
        if (warp_aligned) {
            // Hot loop (aligned messages): the next block is prefetched into registers while this one is expanded,
            // so the chain warp never waits for DRAM.  Runs while ANY lane still has full blocks; lanes that ran
            // out keep re-reading their last block (or nothing) and publish nothing.
            // Lanes without a message (tail of the batch, finished files) shadow a live lane so that they do not
            // force the whole warp onto the slow path: same address, same trip count, nothing of theirs is consumed.
            const unsigned have = __ballot_sync(0xffffffffu, nblk > 0 && nfull > 0);
            const int src = have ? (__ffs(have) - 1) : 0;
            const uint64_t src_ptr = __shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(ptr), src);
            const uint64_t src_nfull = __shfl_sync(0xffffffffu, nfull, src);
            const bool shadow = (nblk == 0);      // only lanes the chain warp will never read for
            const uint8_t* hot_ptr = shadow ? reinterpret_cast<const uint8_t*>(src_ptr) : ptr;
            const uint64_t hot_nfull = shadow ? src_nfull : nfull;
            uint64_t nfull_min = have ? hot_nfull : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const uint64_t other = __shfl_xor_sync(0xffffffffu, nfull_min, o); nfull_min = other < nfull_min ? other : nfull_min; }
            if (nfull_min > 0) {
                const uint4* p4 = reinterpret_cast<const uint4*>(hot_ptr);
                uint4 v0 = ldg128(p4), v1 = ldg128(p4 + 1), v2 = ldg128(p4 + 2), v3 = ldg128(p4 + 3);
                for (; b < nfull_min; ++b) {          // every lane has a full block b here: no divergence, no merges
                    const int st = (int)(b % kCoopStages);
                    if (b >= (uint64_t)kCoopStages) named_bar_sync(kEmpty0 + st, 64);
                    unpack_block(v0, v1, v2, v3, w);
                    p4 += (b + 1 < hot_nfull) ? 4 : 0;
                    v0 = ldg128(p4); v1 = ldg128(p4 + 1); v2 = ldg128(p4 + 2); v3 = ldg128(p4 + 3);
                    expand_store(w, st);
                    named_bar_arrive(kFull0 + st, 64);
                }
            }
        }
        // Everything else: ragged tails of the batch, unaligned messages, padding and length blocks.
        for (; b < nmax; ++b) {
            const int st = (int)(b % kCoopStages);
            if (b >= (uint64_t)kCoopStages) named_bar_sync(kEmpty0 + st, 64);   // chain warp has drained this stage
            if (b < nblk) {
                if (b < nfull) {
                    load_block_unaligned(ptr + (b << 6), w);
                } else if (b == nfull) {
                    const uint8_t* t = ptr + (nfull << 6);
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        uint32_t word = 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t idx = 4 * k + q;
                            if (idx < r) word |= (uint32_t)__ldg(t + idx) << (24 - 8 * q);
                            else if (idx == r) word |= 0x80u << (24 - 8 * q);
                        }
                        w[k] = word;
                    }
                    if (r < 56) { w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits; }
                } else {
#pragma unroll
                    for (int k = 0; k < 14; ++k) w[k] = 0;
                    w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits;
                }
                expand_store(w, st);
            }
            named_bar_arrive(kFull0 + st, 64);
        }
        return;
    }

    // ---------------- chain warp: 64 rounds per block on W+K from shared memory ------------------------
    uint32_t h[8];
    if (valid && j.state != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = j.state[8 * m + i];
    } else {
        sha256_iv(h);
    }
    for (uint64_t b = 0; b < nmax; ++b) {
        const int st = (int)(b % kCoopStages);
        named_bar_sync(kFull0 + st, 64);
        if (b < nblk) {
            uint32_t s[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = h[i];
#pragma unroll
            for (int g = 0; g < 16; ++g) {
                const uint4 v = wk[st][g][lane];
                const uint32_t wkq[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = 4 * g + q;
                    uint32_t& a = s[(0 - t) & 7]; uint32_t& bb = s[(1 - t) & 7]; uint32_t& c = s[(2 - t) & 7];
                    uint32_t& d = s[(3 - t) & 7]; uint32_t& e = s[(4 - t) & 7]; uint32_t& f = s[(5 - t) & 7];
                    uint32_t& gg = s[(6 - t) & 7]; uint32_t& hh = s[(7 - t) & 7];
#if MXD_COOP_SHORT_CHAIN
                    // latency-bound warp: keep only one addition behind Sigma1 on the e-chain (one extra IMAD per round)
                    uint32_t y = add_fma(wkq[q], hh, one);
                    y = add_fma(y, ch(e, f, gg), one);
                    const uint32_t s1 = big_sigma1(e);
                    const uint32_t x = add_fma(d, y, one);
                    const uint32_t z = add_fma(y, maj(a, bb, c), one);
                    d = add_fma(x, s1, one);
                    const uint32_t z2 = add_fma(z, s1, one);
                    hh = add_fma(z2, big_sigma0(a), one);
#else
                    uint32_t t1 = add_fma(wkq[q], hh, one);
                    t1 = add_fma(t1, ch(e, f, gg), one);
                    t1 = add_fma(t1, big_sigma1(e), one);
                    d = add_fma(d, t1, one);
                    const uint32_t t2 = add_fma(big_sigma0(a), maj(a, bb, c), one);
                    hh = add_fma(t1, t2, one);
#endif
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = add_fma(h[i], s[i], one);
        }
        if (b + kCoopStages < nmax) named_bar_arrive(kEmpty0 + st, 64);
    }
    if (!live) return;
    if (!fin) {
#pragma unroll
        for (int i = 0; i < 8; ++i) j.state[8 * m + i] = h[i];
        return;
    }
    uint4 lo, hi;
    lo.x = bswap32(h[0]); lo.y = bswap32(h[1]); lo.z = bswap32(h[2]); lo.w = bswap32(h[3]);
    hi.x = bswap32(h[4]); hi.y = bswap32(h[5]); hi.z = bswap32(h[6]); hi.w = bswap32(h[7]);
    uint4* o = reinterpret_cast<uint4*>(j.out + 32 * m);
    o[0] = lo; o[1] = hi;
}

// One thread: the 72-byte root message of modelx.tree.v1 (two blocks with padding).
__global__ void k_tree_root(uint64_t size, uint64_t leaf, uint32_t fanout, const uint8_t* __restrict__ top,
                            uint8_t* __restrict__ root, uint32_t one) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint8_t msg[128];
    const char magic[16] = {'m', 'o', 'd', 'e', 'l', 'x', '.', 't', 'r', 'e', 'e', '.', 'v', '1', 0, 0};
    for (int i = 0; i < 16; ++i) msg[i] = (uint8_t)magic[i];
    for (int i = 0; i < 8; ++i) { msg[16 + i] = (uint8_t)(size >> (8 * i)); msg[24 + i] = (uint8_t)(leaf >> (8 * i)); }
    for (int i = 0; i < 4; ++i) { msg[32 + i] = (uint8_t)(fanout >> (8 * i)); msg[36 + i] = 0; }
    for (int i = 0; i < 32; ++i) msg[40 + i] = top[i];
    msg[72] = 0x80;
    for (int i = 73; i < 128; ++i) msg[i] = 0;
    msg[126] = (uint8_t)((72 * 8) >> 8); msg[127] = (uint8_t)(72 * 8);
    uint32_t h[8]; sha256_iv(h);
    uint32_t w[16];
    for (int b = 0; b < 2; ++b) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint8_t* q = msg + 64 * b + 4 * k;
            w[k] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
        }
        sha256_compress(h, w, one);
    }
    for (int i = 0; i < 8; ++i) {
        root[4 * i] = (uint8_t)(h[i] >> 24); root[4 * i + 1] = (uint8_t)(h[i] >> 16);
        root[4 * i + 2] = (uint8_t)(h[i] >> 8); root[4 * i + 3] = (uint8_t)h[i];
    }
}

__global__ void k_compare(const uint8_t* __restrict__ got, const uint8_t* __restrict__ want, uint64_t n,
                          uint8_t* __restrict__ ok) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t diff = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) diff |= (uint32_t)(got[32 * i + k] ^ want[32 * i + k]);
    ok[i] = diff == 0;
}

__device__ __forceinline__ uint64_t splitmix64_at(uint64_t seed, uint64_t jdx) {
    uint64_t z = seed + (jdx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void k_gen_fill(uint64_t* __restrict__ dst, uint64_t first_word, uint64_t nwords, uint64_t seed) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride)
        dst[i] = splitmix64_at(seed, first_word + i);
}

}  // namespace

// Occupancy: 24 warps per SM at 80 registers is the default (16..24 warps measure the same, 1.01 TB/s); 32 warps
// at 63 registers is 11 % slower (profiles/r01_quick_bench_v3.txt).  MXD_TUNE_MINB=8 keeps the 32-warp build
// selectable for A/B profiling.
static int g_minb = [] { const char* e = getenv("MXD_TUNE_MINB"); const int v = e ? atoi(e) : 6; return (v == 8 || v == 4) ? v : 6; }();

static long g_coop_max = [] { const char* e = getenv("MXD_TUNE_COOP"); return e ? atol(e) : 32768L; }();

cudaError_t launch_sha256(const MsgJob& job, cudaStream_t stream) {
    if (job.nmsg == 0) return cudaSuccess;
    const uint64_t blocks = (job.nmsg + kThreads - 1) / kThreads;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    // Few, long chains: two warps per 32 chains (see k_sha256_chains_coop).  Measured crossover with the lanes kernel
    // is between 16k chains (coop 675 vs 610 GB/s) and 64k (791 vs 834).  MXD_TUNE_COOP=0 disables, =N sets the threshold.
    if (job.nmsg <= (uint64_t)g_coop_max) {
        const uint64_t cblocks = (job.nmsg + 31) / 32;
        k_sha256_chains_coop<<<(unsigned)cblocks, 64, 0, stream>>>(job);
        return cudaGetLastError();
    }
    if (g_minb == 8)      k_sha256_lanes<16><<<(unsigned)blocks, kThreads, 0, stream>>>(job);
    else if (g_minb == 4) k_sha256_lanes<8><<<(unsigned)blocks, kThreads, 0, stream>>>(job);
    else                  k_sha256_lanes<12><<<(unsigned)blocks, kThreads, 0, stream>>>(job);
    return cudaGetLastError();
}

cudaError_t launch_compare(const uint8_t* got, const uint8_t* want, uint64_t n, uint8_t* ok, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_compare<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(got, want, n, ok);
    return cudaGetLastError();
}

cudaError_t launch_gen_fill(void* dst, uint64_t offset, uint64_t n, uint64_t seed, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if ((offset | n) & 7u || (reinterpret_cast<uintptr_t>(dst) & 7u)) return cudaErrorInvalidValue;
    const uint64_t nwords = n >> 3;
    uint64_t blocks = (nwords + 255) / 256;
    if (blocks > 148ull * 64) blocks = 148ull * 64;
    k_gen_fill<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<uint64_t*>(dst), offset >> 3, nwords, seed);
    return cudaGetLastError();
}

cudaError_t launch_tree_root(uint64_t size, uint64_t leaf, uint32_t fanout, const uint8_t* top, uint8_t* root,
                             cudaStream_t stream) {
    k_tree_root<<<1, 32, 0, stream>>>(size, leaf, fanout, top, root, 1u);
    return cudaGetLastError();
}

int sha256_kernel_regs() {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k_sha256_lanes<12>) != cudaSuccess) return -1;
    return a.numRegs;
}

}  // namespace mxd
