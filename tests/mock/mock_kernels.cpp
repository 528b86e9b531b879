// TEST INFRASTRUCTURE, NOT PRODUCT CODE: CPU stand-ins for the kernel launchers declared in
// modelx_b200/csrc/kernels.h, hashing with the oracle's SHA-256 (oracle/sha256_ref.c).  Linked only into
// tests/mock/_build/libmodelxdigest_mock.so (see tests/mock/include/cuda_runtime.h for why that exists).
// They honour the same job contracts (descriptors, chained state, control bits) as the CUDA kernels, so the host
// code above them runs unchanged.
#include "../../modelx_b200/csrc/kernels.h"
#include "../../oracle/oracle.h"

#include <atomic>
#include <cstdio>
#include <vector>

namespace mockcuda {
Registry& registry() { static Registry r; return r; }
int device_count() {
    const char* e = getenv("MOCK_CUDA_DEVICES");
    const int n = e ? atoi(e) : 1;
    return n < 0 ? 0 : n;
}
int& current_device() { static thread_local int d = 0; return d; }
void* alloc(size_t n, int kind, int device) {
    void* p = nullptr;
    if (posix_memalign(&p, 256, n ? n : 1) != 0) return nullptr;
    Registry& r = registry();
    std::lock_guard<std::mutex> lk(r.mu);
    r.blocks[reinterpret_cast<uintptr_t>(p)] = {n ? n : 1, kind | (device << 8)};
    return p;
}
void release(void* p) {
    if (!p) return;
    Registry& r = registry();
    { std::lock_guard<std::mutex> lk(r.mu); r.blocks.erase(reinterpret_cast<uintptr_t>(p)); }
    free(p);
}
int lookup(const void* p, int* device) {
    Registry& r = registry();
    std::lock_guard<std::mutex> lk(r.mu);
    auto it = r.blocks.upper_bound(reinterpret_cast<uintptr_t>(p));
    if (it == r.blocks.begin()) return 0;
    --it;
    if (reinterpret_cast<uintptr_t>(p) >= it->first + it->second.first) return 0;
    if (device) *device = it->second.second >> 8;
    return it->second.second & 0xff;
}
}  // namespace mockcuda

namespace mxd {

namespace {
// absorb `len` bytes at p into chain state h (len a multiple of 64 unless finalizing); finalize pads with the total length
void chain(uint32_t h[8], const uint8_t* p, uint64_t len, uint64_t prefix, bool fin, uint8_t* out) {
    const uint64_t nfull = len / 64;
    if (nfull) orc_sha256_blocks(h, p, nfull);
    if (!fin) return;
    uint8_t tail[128] = {0};
    const uint64_t r = len % 64;
    memcpy(tail, p + nfull * 64, r);
    tail[r] = 0x80;
    const uint64_t bits = (prefix + len) * 8;
    const size_t tb = r >= 56 ? 128 : 64;
    for (int i = 0; i < 8; ++i) tail[tb - 1 - i] = (uint8_t)(bits >> (8 * i));
    orc_sha256_blocks(h, tail, tb / 64);
    for (int i = 0; i < 8; ++i) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
}
}  // namespace

static std::atomic<uint64_t> g_launches{0};
uint64_t kernel_launch_count() { return g_launches.load(); }

cudaError_t launch_sha256(const MsgJob& j, cudaStream_t) {
    if (j.one != 1) return cudaErrorInvalidValue;
    ++g_launches;
    for (uint64_t m = 0; m < j.nmsg; ++m) {
        const uint8_t* ptr; uint64_t len, prefix = j.prefix_all; uint64_t sidx = m, oidx = m; bool fin = j.finalize != 0, live = true, load = false;
        if (j.descs) {
            const LaneDesc& d = j.descs[m];
            ptr = static_cast<const uint8_t*>(d.ptr); len = d.len; prefix = d.prefix; sidx = d.lane; oidx = d.oidx;
            fin = d.ctl & kFinalize; live = !(d.ctl & kSkip); load = live && !(d.ctl & kFresh);
        } else if (j.base) {
            const uint64_t off = m * j.seg;
            ptr = j.base + off; len = off < j.nbytes ? (j.nbytes - off < j.seg ? j.nbytes - off : j.seg) : 0; load = j.state != nullptr;
        } else {
            const DevSpan& sp = static_cast<const DevSpan*>(j.spans)[m];
            ptr = static_cast<const uint8_t*>(sp.ptr); len = sp.len; load = j.state != nullptr;
        }
        if (!live) continue;
        if (!fin && (len % 64)) { fprintf(stderr, "mock launch_sha256: non-final round of %llu bytes\n", (unsigned long long)len); return cudaErrorInvalidValue; }
        uint32_t h[8];
        if (load) memcpy(h, j.state + 8 * sidx, 32); else orc_sha256_iv(h);
        chain(h, ptr, len, prefix, fin, fin ? j.out + 32 * oidx : nullptr);
        if (!fin) memcpy(j.state + 8 * sidx, h, 32);
    }
    return cudaSuccess;
}

uint32_t leaf_fusable_levels(uint32_t fanout, uint32_t want) {
    uint32_t lv = 0; uint64_t span = 1;
    while (lv < want && span * fanout <= 64 && 64 % (span * fanout) == 0) { span *= fanout; ++lv; }
    return lv;
}
uint64_t leaf_sched_bytes(uint64_t n0) { return 64 + (n0 + 63) / 64 * 4; }
bool leaf_kernel_selected() { const char* f = getenv("MXD_TUNE_FUSE"); return f && atoi(f) > 0; }

cudaError_t launch_tree_leaves(const LeafJob& j, cudaStream_t st) {
    std::vector<uint8_t> cur(j.n0 * 32);
    MsgJob m{}; m.base = j.base; m.nbytes = j.nbytes; m.seg = j.leaf; m.nmsg = j.n0; m.out = cur.data(); m.finalize = 1; m.one = 1;
    cudaError_t e = launch_sha256(m, st);
    uint64_t n = j.n0;
    for (uint32_t lv = 0; lv < j.fused && e == cudaSuccess; ++lv) {
        // groups never straddle a 64-leaf unit because fanout^fused divides 64
        const uint64_t nn = (n + j.fanout - 1) / j.fanout;
        std::vector<uint8_t> next(nn * 32);
        MsgJob u{}; u.base = cur.data(); u.nbytes = n * 32; u.seg = 32ull * j.fanout; u.nmsg = nn; u.out = next.data(); u.finalize = 1; u.one = 1;
        e = launch_sha256(u, st);
        cur.swap(next); n = nn;
    }
    if (e == cudaSuccess) memcpy(j.out, cur.data(), n * 32);
    return e;
}

uint64_t tree_top_scratch_bytes(uint64_t n, uint32_t fanout) { return 2 * (((n + fanout - 1) / fanout) * 32) + 64; }

cudaError_t launch_tree_top(const uint8_t* digests, uint64_t n, uint32_t fanout, uint64_t size, uint64_t leaf,
                            uint8_t*, uint8_t* root, cudaStream_t st) {
    std::vector<uint8_t> cur(digests, digests + n * 32);
    while (n > 1) {
        const uint64_t nn = (n + fanout - 1) / fanout;
        std::vector<uint8_t> next(nn * 32);
        MsgJob u{}; u.base = cur.data(); u.nbytes = n * 32; u.seg = 32ull * fanout; u.nmsg = nn; u.out = next.data(); u.finalize = 1; u.one = 1;
        cudaError_t e = launch_sha256(u, st);
        if (e != cudaSuccess) return e;
        cur.swap(next); n = nn;
    }
    orc_tree_root(size, leaf, fanout, cur.data(), root);
    return cudaSuccess;
}

cudaError_t launch_compare(const uint8_t* got, const uint8_t* want, uint64_t n, uint8_t* ok, cudaStream_t) {
    for (uint64_t i = 0; i < n; ++i) ok[i] = memcmp(got + 32 * i, want + 32 * i, 32) == 0;
    return cudaSuccess;
}
cudaError_t launch_gen_fill(void* dst, uint64_t offset, uint64_t n, uint64_t seed, cudaStream_t) { orc_gen_fill(dst, offset, n, seed); return cudaSuccess; }
int sha256_kernel_regs() { return 0; }

}  // namespace mxd
