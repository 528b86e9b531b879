// C ABI of libmodelxdigest.so (include/modelx_digest.h): contexts, the pinned-host ring that
// streams host/file data to the GPUs, tree assembly, and the integer split helpers.
// Host side of the hot path of kubegems/modelx push/pull; each entry point cites the reference
// call site it replaces in the header.  There is deliberately no CPU hashing in this file: if
// CUDA is unavailable every digest call fails.
#include "../../include/modelx_digest.h"
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <mutex>
#include <sched.h>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

thread_local std::string g_last_error;

int fail(int status, const std::string& msg) {
    g_last_error = msg;
    return status;
}

#define MXD_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(MXD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

// MXD_DEBUG_TIMING=1: print host-side phase timings of the streaming calls to stderr (developer aid)
static const bool g_dbg_timing = getenv("MXD_DEBUG_TIMING") != nullptr;
struct PhaseTimer {
    const char* what; timespec t0;
    explicit PhaseTimer(const char* w) : what(w) { if (g_dbg_timing) clock_gettime(CLOCK_MONOTONIC, &t0); }
    ~PhaseTimer() {
        if (!g_dbg_timing) return;
        timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "[mxd] %-28s %9.3f ms\n", what, (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
    }
};

constexpr uint64_t kDefaultRingBytes = 256ull << 20;
const uint32_t kIVHost[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
constexpr int kSlots = 4;

// ---- NUMA locality ------------------------------------------------------------------------------------
// On an 8-GPU HGX board four GPUs hang off each CPU socket.  Pinned buffers that end up on the other
// socket cross the inter-socket link on every H2D copy (41 instead of 54 GB/s per GPU at N=8 in the first
// runs), so pinned allocations and the slot-filler threads are bound to the CPUs that are local to the
// device (sysfs local_cpulist of its PCI function).  Best effort: any failure leaves affinity untouched.
bool device_local_cpus(int ordinal, cpu_set_t* set) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, ordinal) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char* p = bus; *p; ++p) *p = (char)tolower(*p);
    std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist");
    std::string list;
    if (!f || !std::getline(f, list) || list.empty()) return false;
    CPU_ZERO(set);
    int count = 0;
    size_t i = 0;
    while (i < list.size()) {
        char* end = nullptr;
        long a = strtol(list.c_str() + i, &end, 10), b = a;
        if (end == list.c_str() + i) break;
        i = (size_t)(end - list.c_str());
        if (i < list.size() && list[i] == '-') { b = strtol(list.c_str() + i + 1, &end, 10); i = (size_t)(end - list.c_str()); }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++count; }
        if (i < list.size() && list[i] == ',') ++i;
    }
    return count > 0;
}

// Scoped: run the enclosed allocations (first touch + pin) on the device's local CPUs, then restore.
struct LocalCpuScope {
    cpu_set_t old; bool active = false;
    explicit LocalCpuScope(int ordinal) {
        cpu_set_t want;
        if (getenv("MXD_NO_NUMA_BIND") || !device_local_cpus(ordinal, &want)) return;
        if (sched_getaffinity(0, sizeof old, &old) != 0) return;
        cpu_set_t both; CPU_AND(&both, &old, &want);            // stay inside whatever the container allows
        if (CPU_COUNT(&both) == 0) return;
        active = sched_setaffinity(0, sizeof both, &both) == 0;
    }
    ~LocalCpuScope() { if (active) sched_setaffinity(0, sizeof old, &old); }
};

// Small persistent worker pool that fills pinned ring slots (pread / memcpy) in parallel: one
// thread reads the page cache at 2-4 GB/s, far below the 55 GB/s a PCIe Gen5 x16 link moves.
class StagePool {
public:
    explicit StagePool(int nthreads) {
        for (int i = 0; i < nthreads; ++i) workers_.emplace_back([this] { run(); });
    }
    ~StagePool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // run fn(i) for i in [0, n) on the pool plus the calling thread; returns when all are done
    void parallel_for(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        if (n == 1 || workers_.empty()) { for (int i = 0; i < n; ++i) fn(i); return; }
        Batch b; b.fn = &fn; b.n = n; b.next = 0; b.done = 0;
        { std::lock_guard<std::mutex> lk(mu_); queue_.push_back(&b); }
        cv_.notify_all();
        for (;;) {  // the caller works too
            int i = b.next.fetch_add(1);
            if (i >= n) break;
            fn(i);
            b.done.fetch_add(1);
        }
        std::unique_lock<std::mutex> lk(mu_);
        for (auto it = queue_.begin(); it != queue_.end(); ++it) if (*it == &b) { queue_.erase(it); break; }
        done_cv_.wait(lk, [&] { return b.done.load() >= n && b.active == 0; });
    }
private:
    struct Batch { const std::function<void(int)>* fn; int n; std::atomic<int> next, done; int active = 0; };
    void run() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || !queue_.empty(); });
            if (stop_) return;
            Batch* b = queue_.front();
            if (b->next.load() >= b->n) { queue_.erase(queue_.begin()); continue; }
            b->active++;
            lk.unlock();
            for (;;) {
                int i = b->next.fetch_add(1);
                if (i >= b->n) break;
                (*b->fn)(i);
                b->done.fetch_add(1);
            }
            lk.lock();
            b->active--;
            done_cv_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::vector<Batch*> queue_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    bool stop_ = false;
};

struct DevState {
    int ordinal = -1;
    cudaStream_t compute = nullptr, copy = nullptr;
    uint64_t slot_bytes = 0;
    uint8_t* h_ring = nullptr;  // kSlots * slot_bytes, pinned
    uint8_t* d_ring = nullptr;  // kSlots * slot_bytes
    cudaEvent_t ev_copied[kSlots] = {}, ev_done[kSlots] = {};
    std::mutex mu;               // one streaming operation per device at a time
    StagePool* pool = nullptr;   // slot fillers for this device
    uint8_t* h_desc = nullptr;   // pinned per-round descriptors of lock-step batches (grown on demand, under mu)
    uint64_t h_desc_bytes = 0;
};

}  // namespace

struct mxd_ctx {
    std::vector<DevState*> devs;
    std::atomic<uint64_t> launches{0}, bytes_hashed{0}, h2d{0}, d2h{0};
    std::atomic<int> canceled{0};
    std::atomic<uint32_t> rr{0};  // round-robin device pick for single-device calls
    // live timing of leaf-level launches (mxd_prof_*)
    std::atomic<int> prof_on{0};
    std::mutex prof_mu;
    struct ProfRec { cudaEvent_t a, b; uint64_t bytes; int ordinal; };
    std::vector<ProfRec> prof;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int ordinal) { cudaGetDevice(&prev); cudaSetDevice(ordinal); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Resolved tree parameters: chunk = leaf * fanout^klevel.
struct Tree {
    uint64_t chunk = 8ull << 20, leaf = 16ull << 10, fanout = 8;
    int klevel = 3;
};

bool tree_resolve(const mxd_tree_params* tp, Tree* t) {
    if (tp) { t->chunk = tp->chunk; t->leaf = tp->leaf; t->fanout = tp->fanout; }
    if (t->leaf < 64 || (t->leaf % 64) != 0 || t->fanout < 2 || t->chunk < t->leaf) return false;
    if (tp && tp->reserved != 0) return false;
    uint64_t span = t->leaf;
    int k = 0;
    while (span < t->chunk) {
        if (span > t->chunk / t->fanout) return false;   // overflow / not a power
        span *= t->fanout; ++k;
    }
    if (span != t->chunk || k < 1) return false;
    t->klevel = k;
    return true;
}

// ---- kernel enqueue helpers ------------------------------------------------------------------
int enqueue_segments(mxd_ctx* c, const uint8_t* d_data, uint64_t nbytes, uint64_t seg, uint8_t* d_out, cudaStream_t st,
                     bool leaf_level = false) {
    mxd_ctx::ProfRec rec{};
    bool prof = leaf_level && c->prof_on.load();
    if (prof) { std::lock_guard<std::mutex> lk(c->prof_mu); if (c->prof.size() >= (1u << 16)) prof = false; }   // bounded
    if (prof) {
        cudaGetDevice(&rec.ordinal);
        MXD_CUDA(cudaEventCreate(&rec.a));
        MXD_CUDA(cudaEventCreate(&rec.b));
        rec.bytes = nbytes;
        MXD_CUDA(cudaEventRecord(rec.a, st));
    }
    mxd::MsgJob j{};
    j.base = d_data; j.nbytes = nbytes; j.seg = seg;
    j.nmsg = nbytes ? (nbytes + seg - 1) / seg : 1;
    j.out = d_out; j.finalize = 1; j.one = 1;
    if (j.base == nullptr) {  // empty message: any non-null base keeps the kernel in segment mode
        j.base = reinterpret_cast<const uint8_t*>(d_out);
    }
    MXD_CUDA(mxd::launch_sha256(j, st));
    c->launches++; c->bytes_hashed += nbytes;
    if (prof) {
        MXD_CUDA(cudaEventRecord(rec.b, st));
        std::lock_guard<std::mutex> lk(c->prof_mu);
        c->prof.push_back(rec);
    }
    return MXD_OK;
}

// digests of one tree level -> the next: groups of `fanout` digests
int enqueue_level(mxd_ctx* c, const uint8_t* d_in, uint64_t n, uint64_t fanout, uint8_t* d_out, cudaStream_t st) {
    return enqueue_segments(c, d_in, n * 32, 32 * fanout, d_out, st);
}

// leaf digests (level 0, n0 of them in d_leaves, clobbered as scratch) -> chunk digests (level k)
int enqueue_leaves_to_chunks(mxd_ctx* c, const Tree& t, uint8_t* d_leaves, uint64_t n0, uint8_t* d_chunks, cudaStream_t st) {
    // ping-pong between the leaf buffer and one scratch buffer of the level-1 size
    uint64_t n = n0;
    const uint64_t n1 = (n0 + t.fanout - 1) / t.fanout;
    uint8_t* scratch = nullptr;
    if (t.klevel > 1) MXD_CUDA(cudaMallocAsync(&scratch, n1 * 32, st));
    uint8_t* cur = d_leaves;
    int rc = MXD_OK;
    for (int lv = 1; lv <= t.klevel && rc == MXD_OK; ++lv) {
        const uint64_t nn = (n + t.fanout - 1) / t.fanout;
        uint8_t* dst = (lv == t.klevel) ? d_chunks : ((cur == d_leaves) ? scratch : d_leaves);
        rc = enqueue_level(c, cur, n, t.fanout, dst, st);
        cur = dst; n = nn;
    }
    if (scratch) cudaFreeAsync(scratch, st);
    return rc;
}

// leaves -> chunk digests for one piece resident in device memory
int enqueue_tree_chunks(mxd_ctx* c, const Tree& t, const uint8_t* d_piece, uint64_t nbytes, uint8_t* d_chunks, cudaStream_t st) {
    const uint64_t n0 = nbytes ? (nbytes + t.leaf - 1) / t.leaf : 1;
    uint8_t* ws = nullptr;
    MXD_CUDA(cudaMallocAsync(&ws, n0 * 32, st));
    int rc = enqueue_segments(c, d_piece, nbytes, t.leaf, ws, st, /*leaf_level=*/true);
    if (rc == MXD_OK) rc = enqueue_leaves_to_chunks(c, t, ws, n0, d_chunks, st);
    cudaFreeAsync(ws, st);
    return rc;
}

// levels above the chunk list, then the root message
int enqueue_tree_finish(mxd_ctx* c, const Tree& t, const uint8_t* d_chunks, uint64_t nchunks, uint64_t size,
                        uint8_t* d_root, cudaStream_t st) {
    const uint8_t* cur = d_chunks;
    uint64_t n = nchunks;
    uint8_t* bufs[2] = {nullptr, nullptr};
    int rc = MXD_OK, which = 0;
    if (n > 1) {
        const uint64_t n1 = (n + t.fanout - 1) / t.fanout;
        for (int i = 0; i < 2 && rc == MXD_OK; ++i) {
            cudaError_t e = cudaMallocAsync(&bufs[i], n1 * 32, st);
            if (e != cudaSuccess) rc = fail(MXD_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
        }
    }
    while (n > 1 && rc == MXD_OK) {
        rc = enqueue_level(c, cur, n, t.fanout, bufs[which], st);
        cur = bufs[which]; which ^= 1;
        n = (n + t.fanout - 1) / t.fanout;
    }
    if (rc == MXD_OK) {
        cudaError_t e = mxd::launch_tree_root(size, t.leaf, (uint32_t)t.fanout, cur, d_root, st);
        if (e != cudaSuccess) rc = fail(MXD_ERR_CUDA, std::string("launch_tree_root: ") + cudaGetErrorString(e));
        else c->launches++;
    }
    for (int i = 0; i < 2; ++i) if (bufs[i]) cudaFreeAsync(bufs[i], st);
    return rc;
}

enum class MemKind { Pageable, Pinned, Device };

MemKind classify(const void* p, int* device_ordinal) {
    cudaPointerAttributes a{};
    if (p == nullptr || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return MemKind::Pageable; }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) {
        if (device_ordinal) *device_ordinal = a.device;
        return MemKind::Device;
    }
    if (a.type == cudaMemoryTypeHost) return MemKind::Pinned;
    return MemKind::Pageable;
}

int dev_index_of(const mxd_ctx* c, int ordinal) {
    for (size_t i = 0; i < c->devs.size(); ++i)
        if (c->devs[i]->ordinal == ordinal) return (int)i;
    return -1;
}

// ---- data sources for the streaming ring -------------------------------------------------------
struct Source {
    const uint8_t* mem = nullptr;  // host memory source
    bool pinned = false;
    int fd = -1;                   // file source
    uint64_t base = 0;             // offset of this source's byte 0 inside the file
    mxd_sink_fn sink = nullptr;    // optional tee: every streamed byte is also handed to this callback
    void* sink_user = nullptr;
    uint64_t sink_base = 0;        // logical offset of this source's byte 0 for the sink
};

// hand [off, off+n) (already staged at `data`) to the tee in <= 4 MiB pieces, in parallel on the pool
int sink_slot(const Source& s, uint64_t off, uint64_t n, const uint8_t* data, StagePool* pool) {
    if (!s.sink) return MXD_OK;
    constexpr uint64_t kPiece = 4ull << 20;
    const int pieces = (int)((n + kPiece - 1) / kPiece);
    std::atomic<int> bad{0};
    auto put = [&](int i) {
        const uint64_t p0 = (uint64_t)i * kPiece, pn = std::min(kPiece, n - p0);
        if (s.sink(s.sink_user, s.sink_base + off + p0, data + p0, pn) != 0) bad.store(1);
    };
    if (pool && pieces > 1) pool->parallel_for(pieces, put); else for (int i = 0; i < pieces; ++i) put(i);
    return bad.load() ? fail(MXD_ERR_IO, "sink refused data") : MXD_OK;
}

// Fill `n` bytes at logical offset `off` of the source into pinned `dst`; returns the pointer the
// H2D copy should read from (dst, or the caller's own memory when that is already pinned).
// Large fills are split into 4 MiB pieces across the device's StagePool.
int source_stage(const Source& s, uint64_t off, uint64_t n, uint8_t* dst, const uint8_t** from, StagePool* pool = nullptr) {
    if (s.fd < 0 && s.pinned) { *from = s.mem + off; return MXD_OK; }
    *from = dst;
    static const uint64_t kPiece = [] { const char* e = getenv("MXD_STAGE_PIECE"); uint64_t v = e ? strtoull(e, nullptr, 10) : 0; return v >= 4096 ? v : (4ull << 20); }();
    const int pieces = (int)((n + kPiece - 1) / kPiece);
    std::atomic<int> err{0};
    auto fill = [&](int i) {
        const uint64_t p0 = (uint64_t)i * kPiece, pn = std::min(kPiece, n - p0);
        if (s.fd >= 0) {
            uint64_t got = 0;
            while (got < pn) {
                ssize_t r = pread(s.fd, dst + p0 + got, pn - got, (off_t)(s.base + off + p0 + got));
                if (r < 0) { if (errno == EINTR) continue; err.store(errno ? errno : EIO); return; }
                if (r == 0) { err.store(-1); return; }
                got += (uint64_t)r;
            }
        } else {
            memcpy(dst + p0, s.mem + off + p0, pn);
        }
    };
    if (pool && pieces > 1) pool->parallel_for(pieces, fill);
    else for (int i = 0; i < pieces; ++i) fill(i);
    const int e = err.load();
    if (e == -1) return fail(MXD_ERR_IO, "pread: file shrank while hashing");
    if (e) return fail(MXD_ERR_IO, std::string("pread: ") + strerror(e));
    return MXD_OK;
}

// Stream [0, nbytes) of `src` through device `d`'s ring and hash it as uniform segments of `seg`
// bytes into d_out (device, ceil(nbytes/seg) digests).  slot_bytes is a multiple of seg.
// H2D copies run on the copy stream, kernels on the compute stream; a slot is refilled only
// after the kernel that read it has finished, so copy k+1.. overlap kernel k.
int stream_segments(mxd_ctx* c, DevState* d, const Source& src, uint64_t nbytes, uint64_t seg, uint8_t* d_out) {
    const uint64_t per_slot = (d->slot_bytes / seg) * seg;
    if (per_slot == 0) return fail(MXD_ERR_INVALID, "ring slot smaller than one segment; raise ring_bytes");
    if (nbytes == 0) return enqueue_segments(c, nullptr, 0, seg, d_out, d->compute);
    uint64_t off = 0;
    for (uint64_t i = 0; off < nbytes; ++i) {
        if (c->canceled.load()) return fail(MXD_ERR_CANCELED, "canceled");
        const int s = (int)(i % kSlots);
        const uint64_t n = (nbytes - off < per_slot) ? nbytes - off : per_slot;
        if (i >= (uint64_t)kSlots) MXD_CUDA(cudaEventSynchronize(d->ev_done[s]));
        uint8_t* h_slot = d->h_ring + (uint64_t)s * d->slot_bytes;
        uint8_t* d_slot = d->d_ring + (uint64_t)s * d->slot_bytes;
        const uint8_t* from = nullptr;
        int rc;
        { PhaseTimer pt("  fill slot"); rc = source_stage(src, off, n, h_slot, &from, d->pool); }
        if (rc != MXD_OK) return rc;
        MXD_CUDA(cudaMemcpyAsync(d_slot, from, n, cudaMemcpyHostToDevice, d->copy));
        MXD_CUDA(cudaEventRecord(d->ev_copied[s], d->copy));
        MXD_CUDA(cudaStreamWaitEvent(d->compute, d->ev_copied[s], 0));
        rc = enqueue_segments(c, d_slot, n, seg, d_out + (off / seg) * 32, d->compute, /*leaf_level=*/true);
        if (rc != MXD_OK) return rc;
        MXD_CUDA(cudaEventRecord(d->ev_done[s], d->compute));
        rc = sink_slot(src, off, n, from, d->pool);     // the tee runs while the copy engine and the SMs work on this slot
        if (rc != MXD_OK) return rc;
        c->h2d += n;
        off += n;
    }
    return MXD_OK;
}

// chunk digests of a host/file piece on one device; result left in device memory d_chunks
int stream_tree_chunks(mxd_ctx* c, DevState* d, const Tree& t, const Source& src, uint64_t nbytes, uint8_t* d_chunks) {
    const uint64_t n0 = nbytes ? (nbytes + t.leaf - 1) / t.leaf : 1;
    uint8_t* d_leaves = nullptr;
    { PhaseTimer pt("alloc leaf digests"); MXD_CUDA(cudaMallocAsync(&d_leaves, n0 * 32, d->compute)); }
    int rc;
    { PhaseTimer pt("stream_segments (enqueue)"); rc = stream_segments(c, d, src, nbytes, t.leaf, d_leaves); }
    if (rc == MXD_OK) rc = enqueue_leaves_to_chunks(c, t, d_leaves, n0, d_chunks, d->compute);
    if (rc != MXD_OK) cudaStreamSynchronize(d->copy);   // nothing may still be reading the caller's (pinned) memory
    cudaError_t e;
    { PhaseTimer pt("sync compute"); e = cudaStreamSynchronize(d->compute); }
    if (rc == MXD_OK && e != cudaSuccess) rc = fail(MXD_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
    cudaFreeAsync(d_leaves, d->compute);
    return rc;
}

// Chunk digests of a host/file blob using every device of the context: device g takes the
// contiguous chunk range [g*n/G, (g+1)*n/G) (sequential reads per device, no data-path
// collective).  Results land in host memory `out` (nchunks*32).
int host_tree_chunks_all(mxd_ctx* c, const Tree& t, const Source& src, uint64_t nbytes, uint8_t* out) {
    const uint64_t chunk = t.chunk;
    const uint64_t nchunks = nbytes ? (nbytes + chunk - 1) / chunk : 1;
    const int G = (int)std::min<uint64_t>(c->devs.size(), nchunks);
    std::vector<int> rcs(G, MXD_OK);
    std::vector<std::string> errs(G);
    auto work = [&](int g) {
        DevState* d = c->devs[g];
        std::lock_guard<std::mutex> lk(d->mu);
        DeviceGuard guard(d->ordinal);
        const uint64_t c0 = nchunks * g / G, c1 = nchunks * (g + 1) / G;
        const uint64_t b0 = c0 * chunk, b1 = std::min<uint64_t>(c1 * chunk, nbytes);
        Source piece = src;
        if (piece.fd >= 0) piece.base += b0; else piece.mem += b0;
        piece.sink_base += b0;
        uint8_t* d_chunks = nullptr;
        cudaError_t e = cudaMallocAsync(&d_chunks, (c1 - c0) * 32, d->compute);
        if (e != cudaSuccess) { rcs[g] = MXD_ERR_CUDA; errs[g] = cudaGetErrorString(e); return; }
        int rc = stream_tree_chunks(c, d, t, piece, b1 - b0, d_chunks);
        if (rc == MXD_OK) {
            PhaseTimer pt("chunk digests D2H");
            e = cudaMemcpyAsync(out + c0 * 32, d_chunks, (c1 - c0) * 32, cudaMemcpyDeviceToHost, d->compute);
            if (e == cudaSuccess) e = cudaStreamSynchronize(d->compute);
            if (e != cudaSuccess) { rc = MXD_ERR_CUDA; g_last_error = cudaGetErrorString(e); }
            c->d2h += (c1 - c0) * 32;
        }
        cudaFreeAsync(d_chunks, d->compute);
        rcs[g] = rc;
        if (rc != MXD_OK) errs[g] = g_last_error;
    };
    if (G == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) th.emplace_back(work, g);
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < G; ++g)
        if (rcs[g] != MXD_OK) return fail(rcs[g], errs[g]);
    return MXD_OK;
}

// upper levels + root from a host-resident chunk list (tiny: 32 B per chunk)
int host_tree_finish(mxd_ctx* c, DevState* d, const Tree& t, const uint8_t* chunks, uint64_t nchunks, uint64_t size,
                     uint8_t root[32]) {
    PhaseTimer pt("tree finish (levels + root)");
    std::lock_guard<std::mutex> lk(d->mu);
    DeviceGuard guard(d->ordinal);
    uint8_t* d_buf = nullptr;
    MXD_CUDA(cudaMallocAsync(&d_buf, nchunks * 32 + 32, d->compute));
    int rc = MXD_OK;
    cudaError_t e = cudaMemcpyAsync(d_buf, chunks, nchunks * 32, cudaMemcpyHostToDevice, d->compute);
    if (e != cudaSuccess) rc = fail(MXD_ERR_CUDA, cudaGetErrorString(e));
    if (rc == MXD_OK) rc = enqueue_tree_finish(c, t, d_buf, nchunks, size, d_buf + nchunks * 32, d->compute);
    if (rc == MXD_OK) {
        e = cudaMemcpyAsync(root, d_buf + nchunks * 32, 32, cudaMemcpyDeviceToHost, d->compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d->compute);
        if (e != cudaSuccess) rc = fail(MXD_ERR_CUDA, cudaGetErrorString(e));
        c->h2d += nchunks * 32; c->d2h += 32;
    } else {
        cudaStreamSynchronize(d->compute);
    }
    cudaFreeAsync(d_buf, d->compute);
    return rc;
}

DevState* pick_device(mxd_ctx* c) { return c->devs[c->rr++ % c->devs.size()]; }

// ---- lock-step hashing of n whole messages from host sources (files or host spans) -------------
// Round k moves bytes [k*S, (k+1)*S) of every still-running message into one ring slot
// (message i at slot offset i*S) and advances all chains together; each lane carries its chain
// state in device memory between rounds.  This is the reference's one-digest-per-file result
// (push.go:149-161) for many files at once.
struct LockstepInput {
    std::vector<Source> src;
    std::vector<uint64_t> len;
};

int lockstep_digest(mxd_ctx* c, DevState* d, const LockstepInput& in, uint8_t* out) {
    const uint64_t n = in.src.size();
    if (n == 0) return MXD_OK;
    std::lock_guard<std::mutex> lk(d->mu);
    DeviceGuard guard(d->ordinal);
    if ((d->slot_bytes / n) < 64) return fail(MXD_ERR_INVALID, "too many messages for the ring slot; raise ring_bytes or split the batch");
    // Per round every still-running message advances by S bytes, S = slot / (running messages) rounded down to
    // 64 and capped at 8 MiB; finished messages give their share of the slot to the rest, and only the
    // occupied part of the slot is copied.
    std::vector<uint64_t> done(n, 0);
    std::vector<uint8_t> finished(n, 0);
    uint64_t running = n;

    // per-round descriptors live in pinned memory, double-buffered per slot
    const uint64_t desc_bytes = (n * (sizeof(mxd::DevSpan) + sizeof(uint64_t) + 1) + 255) & ~255ull;
    // Scratch: pinned descriptors are kept in the device state (cudaMallocHost costs ~0.5 ms, too much for small
    // calls); device buffers come from the stream-ordered pool, which mxd_open told to keep freed memory cached.
    if (d->h_desc_bytes < desc_bytes * kSlots) {
        if (d->h_desc) cudaFreeHost(d->h_desc);
        d->h_desc = nullptr; d->h_desc_bytes = 0;
        const uint64_t want = std::max<uint64_t>(desc_bytes * kSlots, 64 << 10);
        MXD_CUDA(cudaHostAlloc(&d->h_desc, want, cudaHostAllocPortable));
        d->h_desc_bytes = want;
    }
    uint8_t *h_desc = d->h_desc, *d_desc = nullptr, *d_out = nullptr;
    uint32_t* d_state = nullptr;
    int rc = MXD_OK;
    cudaError_t e;
    if ((e = cudaMallocAsync(&d_desc, desc_bytes * kSlots, d->compute)) != cudaSuccess ||
        (e = cudaMallocAsync(&d_out, n * 32, d->compute)) != cudaSuccess ||
        (e = cudaMallocAsync(&d_state, n * 32, d->compute)) != cudaSuccess) {
        rc = fail(MXD_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    }
    if (rc == MXD_OK) {  // every chain starts from the FIPS 180-4 initial hash value
        std::vector<uint32_t> iv(n * 8);
        for (uint64_t i = 0; i < n; ++i) memcpy(&iv[8 * i], kIVHost, 32);
        e = cudaMemcpyAsync(d_state, iv.data(), n * 32, cudaMemcpyHostToDevice, d->compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d->compute);   // iv is a stack/heap temporary; also orders the pool allocations before the copy stream uses them
        if (e != cudaSuccess) rc = fail(MXD_ERR_CUDA, cudaGetErrorString(e));
    }
    for (uint64_t k = 0; running > 0 && rc == MXD_OK; ++k) {
        if (c->canceled.load()) { rc = fail(MXD_ERR_CANCELED, "canceled"); break; }
        const int s = (int)(k % kSlots);
        if (k >= (uint64_t)kSlots) {
            e = cudaEventSynchronize(d->ev_done[s]);
            if (e != cudaSuccess) { rc = fail(MXD_ERR_CUDA, cudaGetErrorString(e)); break; }
        }
        uint8_t* h_slot = d->h_ring + (uint64_t)s * d->slot_bytes;
        uint8_t* d_slot = d->d_ring + (uint64_t)s * d->slot_bytes;
        uint8_t* hd = h_desc + desc_bytes * s;
        auto* spans = reinterpret_cast<mxd::DevSpan*>(hd);
        auto* prefix = reinterpret_cast<uint64_t*>(hd + n * sizeof(mxd::DevSpan));
        uint8_t* ctl = hd + n * (sizeof(mxd::DevSpan) + sizeof(uint64_t));
        const uint64_t S = std::min<uint64_t>(8ull << 20, (d->slot_bytes / running) & ~63ull);
        uint64_t moved = 0, pos = 0;
        struct Fill { uint64_t i, done, take, pos; };
        std::vector<Fill> fills;
        for (uint64_t i = 0; i < n; ++i) {
            if (finished[i]) { spans[i] = {d_slot, 0}; prefix[i] = in.len[i]; ctl[i] = 2; continue; }
            const uint64_t take = std::min(in.len[i] - done[i], S);
            const bool last = done[i] + take == in.len[i];     // finalised in the round that reaches its end
            spans[i] = {d_slot + pos * S, take};
            prefix[i] = done[i];
            ctl[i] = last ? 1 : 0;
            if (take) { fills.push_back({i, done[i], take, pos}); ++pos; }   // empty messages occupy no slot space
            moved += take; done[i] += take;
            if (last) finished[i] = 1;
        }
        const uint64_t occupied = pos * S;
        running = 0;
        for (uint64_t i = 0; i < n; ++i) running += finished[i] ? 0 : 1;
        // gather this round's bytes of every running message into the slot, files read in parallel
        std::vector<int> frc(fills.size(), MXD_OK);
        std::vector<std::string> ferr(fills.size());
        d->pool->parallel_for((int)fills.size(), [&](int f) {
            Source one = in.src[fills[f].i];
            one.pinned = false;   // always copy into the slot so the whole round is one H2D transfer
            const uint8_t* from = nullptr;
            frc[f] = source_stage(one, fills[f].done, fills[f].take, h_slot + fills[f].pos * S, &from);
            if (frc[f] != MXD_OK) ferr[f] = g_last_error;
        });
        for (size_t f = 0; f < fills.size(); ++f) if (frc[f] != MXD_OK) { rc = fail(frc[f], ferr[f]); break; }
        if (rc != MXD_OK) break;
        const uint64_t span_bytes = occupied;
        e = span_bytes ? cudaMemcpyAsync(d_slot, h_slot, span_bytes, cudaMemcpyHostToDevice, d->copy) : cudaSuccess;
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_desc + desc_bytes * s, hd, desc_bytes, cudaMemcpyHostToDevice, d->copy);
        if (e == cudaSuccess) e = cudaEventRecord(d->ev_copied[s], d->copy);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(d->compute, d->ev_copied[s], 0);
        if (e != cudaSuccess) { rc = fail(MXD_ERR_CUDA, cudaGetErrorString(e)); break; }
        mxd::MsgJob j{};
        uint8_t* dd = d_desc + desc_bytes * s;
        j.spans = dd; j.nmsg = n; j.out = d_out; j.state = d_state;
        j.prefix = reinterpret_cast<const uint64_t*>(dd + n * sizeof(mxd::DevSpan));
        j.ctl = dd + n * (sizeof(mxd::DevSpan) + sizeof(uint64_t));
        j.finalize = 0; j.one = 1;
        e = mxd::launch_sha256(j, d->compute);
        if (e == cudaSuccess) e = cudaEventRecord(d->ev_done[s], d->compute);
        if (e != cudaSuccess) { rc = fail(MXD_ERR_CUDA, cudaGetErrorString(e)); break; }
        c->launches++; c->bytes_hashed += moved; c->h2d += span_bytes + desc_bytes;
    }
    if (rc == MXD_OK) {
        e = cudaMemcpyAsync(out, d_out, n * 32, cudaMemcpyDeviceToHost, d->compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d->compute);
        if (e != cudaSuccess) rc = fail(MXD_ERR_CUDA, cudaGetErrorString(e));
        c->d2h += n * 32;
    } else {
        cudaStreamSynchronize(d->compute); cudaStreamSynchronize(d->copy);
    }
    if (d_state) cudaFreeAsync(d_state, d->compute);
    if (d_out) cudaFreeAsync(d_out, d->compute);
    if (d_desc) cudaFreeAsync(d_desc, d->compute);
    return rc;
}

// Spread a batch of whole messages over every device of the context (largest first onto the least
// loaded device), one lock-step batch per device running concurrently; results keep the caller's order.
int lockstep_digest_all(mxd_ctx* c, const LockstepInput& in, uint8_t* out) {
    const uint64_t n = in.src.size();
    const size_t G = std::min<uint64_t>(c->devs.size(), std::max<uint64_t>(n, 1));
    if (G <= 1) return lockstep_digest(c, pick_device(c), in, out);
    std::vector<uint64_t> order(n);
    for (uint64_t i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return in.len[a] > in.len[b]; });
    std::vector<std::vector<uint64_t>> group(G);
    std::vector<uint64_t> load(G, 0);
    for (uint64_t idx : order) {
        size_t g = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
        group[g].push_back(idx); load[g] += in.len[idx] + 1;
    }
    std::vector<int> rcs(G, MXD_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g) th.emplace_back([&, g] {
        LockstepInput part;
        for (uint64_t idx : group[g]) { part.src.push_back(in.src[idx]); part.len.push_back(in.len[idx]); }
        std::vector<uint8_t> o(32 * part.src.size());
        rcs[g] = lockstep_digest(c, c->devs[g], part, o.data());
        if (rcs[g] != MXD_OK) { errs[g] = g_last_error; return; }
        for (size_t k = 0; k < group[g].size(); ++k) memcpy(out + 32 * group[g][k], &o[32 * k], 32);
    });
    for (auto& t : th) t.join();
    for (size_t g = 0; g < G; ++g) if (rcs[g] != MXD_OK) return fail(rcs[g], errs[g]);
    return MXD_OK;
}

int open_files(const char* const* paths, uint64_t n, LockstepInput* in, std::vector<int>* fds) {
    in->src.resize(n); in->len.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
        if (fd < 0) { int e = errno; for (int f : *fds) close(f); fds->clear(); errno = e;
                      return fail(MXD_ERR_IO, std::string("open ") + paths[i] + ": " + strerror(e)); }
        fds->push_back(fd);
        struct stat st;
        if (fstat(fd, &st) != 0) { int e = errno; for (int f : *fds) close(f); fds->clear(); errno = e;
                                   return fail(MXD_ERR_IO, std::string("fstat ") + paths[i] + ": " + strerror(e)); }
        in->src[i].fd = fd; in->len[i] = (uint64_t)st.st_size;
    }
    return MXD_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int mxd_abi_version(void) { return MXD_ABI_VERSION; }

const char* mxd_strerror(int status) {
    switch (status) {
        case MXD_OK: return "ok";
        case MXD_ERR_INVALID: return "invalid argument";
        case MXD_ERR_NO_DEVICE: return "no CUDA device available (modelx-b200 has no CPU fallback)";
        case MXD_ERR_CUDA: return "CUDA error";
        case MXD_ERR_IO: return "I/O error";
        case MXD_ERR_NOMEM: return "out of memory";
        case MXD_ERR_CANCELED: return "canceled";
        case MXD_ERR_DIV_ZERO: return "integer divide by zero (calcParts with 0 parts)";
        default: return "unknown status";
    }
}

const char* mxd_last_error(void) { return g_last_error.c_str(); }

int mxd_prof_enable(mxd_ctx* c, int on) {
    if (!c) return fail(MXD_ERR_INVALID, "prof_enable: null");
    c->prof_on.store(on ? 1 : 0);
    return MXD_OK;
}

int mxd_prof_read(mxd_ctx* c, double* kernel_ms, uint64_t* launches, uint64_t* bytes) {
    if (!c) return fail(MXD_ERR_INVALID, "prof_read: null");
    std::vector<mxd_ctx::ProfRec> recs;
    { std::lock_guard<std::mutex> lk(c->prof_mu); recs.swap(c->prof); }
    double ms = 0; uint64_t nb = 0;
    int rc = MXD_OK;
    for (auto& r : recs) {
        DeviceGuard guard(r.ordinal);
        float t = 0;
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
        if (e != cudaSuccess) rc = fail(MXD_ERR_CUDA, std::string("prof_read: ") + cudaGetErrorString(e));
        ms += t; nb += r.bytes;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    if (kernel_ms) *kernel_ms = ms;
    if (launches) *launches = recs.size();
    if (bytes) *bytes = nb;
    return rc;
}

int mxd_open(mxd_ctx** out, const int* devices, int ndev, uint64_t ring_bytes) {
    if (!out || ndev < 0 || (ndev > 0 && !devices)) return fail(MXD_ERR_INVALID, "mxd_open: bad arguments");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(MXD_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + (e == cudaSuccess ? "0 devices" : cudaGetErrorString(e)));
    }
    std::vector<int> ords;
    if (ndev == 0) for (int i = 0; i < count; ++i) ords.push_back(i);
    else for (int i = 0; i < ndev; ++i) {
        if (devices[i] < 0 || devices[i] >= count) return fail(MXD_ERR_INVALID, "mxd_open: device ordinal out of range");
        ords.push_back(devices[i]);
    }
    if (const char* env = getenv("MXD_RING_BYTES")) { uint64_t v = strtoull(env, nullptr, 10); if (v) ring_bytes = v; }
    if (ring_bytes == 0) ring_bytes = kDefaultRingBytes;
    uint64_t slot = (ring_bytes / kSlots) & ~((1ull << 20) - 1);
    if (slot < (1ull << 20)) slot = 1ull << 20;

    auto* c = new mxd_ctx();
    int prev = -1; cudaGetDevice(&prev);
    int rc = MXD_OK;
    // slot-filler threads per device: MXD_STAGE_THREADS, default min(16, hw threads / devices), at least 1
    // (page-cache pread runs at 2-4 GB/s per thread; 16 threads gave 43.5 GB/s from a tmpfs file, 32 only 30.5)
    int stage_threads = 0;
    if (const char* env = getenv("MXD_STAGE_THREADS")) stage_threads = atoi(env);
    if (stage_threads <= 0) {
        const unsigned hw = std::thread::hardware_concurrency();
        stage_threads = (int)std::min<unsigned>(16, std::max<unsigned>(1, hw / (unsigned)ords.size()));
    }
    for (int ord : ords) {
        auto* d = new DevState();
        d->ordinal = ord; d->slot_bytes = slot;
        c->devs.push_back(d);
        if ((e = cudaSetDevice(ord)) != cudaSuccess) break;
        LocalCpuScope numa(ord);     // the pool's threads inherit this affinity; the pinned ring is first-touched here
        d->pool = new StagePool(stage_threads - 1);
        if ((e = cudaStreamCreateWithFlags(&d->compute, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaStreamCreateWithFlags(&d->copy, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaHostAlloc(&d->h_ring, slot * kSlots, cudaHostAllocPortable)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d->d_ring, slot * kSlots)) != cudaSuccess) break;
        for (int s = 0; s < kSlots && e == cudaSuccess; ++s) {
            e = cudaEventCreateWithFlags(&d->ev_copied[s], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_done[s], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) break;
        // keep stream-ordered workspace allocations cached instead of returning them to the OS
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ord) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) {
        rc = fail(MXD_ERR_CUDA, std::string("mxd_open: ") + cudaGetErrorString(e));
        mxd_close(c);
        return rc;
    }
    *out = c;
    return MXD_OK;
}

void mxd_close(mxd_ctx* c) {
    if (!c) return;
    int prev = -1; cudaGetDevice(&prev);
    for (DevState* d : c->devs) {
        if (d->ordinal >= 0) cudaSetDevice(d->ordinal);
        if (d->compute) { cudaStreamSynchronize(d->compute); cudaStreamDestroy(d->compute); }
        if (d->copy) { cudaStreamSynchronize(d->copy); cudaStreamDestroy(d->copy); }
        for (int s = 0; s < kSlots; ++s) {
            if (d->ev_copied[s]) cudaEventDestroy(d->ev_copied[s]);
            if (d->ev_done[s]) cudaEventDestroy(d->ev_done[s]);
        }
        if (d->h_ring) cudaFreeHost(d->h_ring);
        if (d->d_ring) cudaFree(d->d_ring);
        if (d->h_desc) cudaFreeHost(d->h_desc);
        delete d->pool;
        delete d;
    }
    if (prev >= 0) cudaSetDevice(prev);
    delete c;
}

int mxd_device_count(const mxd_ctx* c) { return c ? (int)c->devs.size() : 0; }
void mxd_cancel(mxd_ctx* c) { if (c) c->canceled.store(1); }
void mxd_reset_cancel(mxd_ctx* c) { if (c) c->canceled.store(0); }

int mxd_get_stats(const mxd_ctx* c, mxd_stats* out) {
    if (!c || !out) return fail(MXD_ERR_INVALID, "mxd_get_stats: null");
    memset(out, 0, sizeof *out);
    out->kernel_launches = c->launches.load(); out->bytes_hashed = c->bytes_hashed.load();
    out->h2d_bytes = c->h2d.load(); out->d2h_bytes = c->d2h.load();
    return MXD_OK;
}

// ---- integer split: extension_s3.go:99-112, store_s3.go:198-203,273-279 --------------------------
int mxd_calc_parts(int64_t total, int64_t partscount, mxd_part* out) {
    if (partscount == 0) return fail(MXD_ERR_DIV_ZERO, "calcParts: partscount == 0 (reference panics: integer divide by zero)");
    if (partscount < 0 || !out) return fail(MXD_ERR_INVALID, "calcParts: negative part count");
    const int64_t partsize = total / partscount;
    for (int64_t i = 0; i < partscount; ++i) {
        out[i].offset = i * partsize;
        out[i].length = (i == partscount - 1) ? total - out[i].offset : partsize;
    }
    return MXD_OK;
}

int64_t mxd_server_part_count(int64_t size, int force_multipart) {
    const int64_t kThreshold = 5ll << 30;   // MultiPartUploadThreshold
    const int64_t kDefaultParts = 3;        // DefaultPartCount
    if (!force_multipart && size <= kThreshold) return 1;
    int64_t count = size / kThreshold;
    if (count == 0) return kDefaultParts;
    return (size % kThreshold) ? count + 1 : count;
}

// ---- digest strings -------------------------------------------------------------------------------
void mxd_digest_string(const uint8_t d[32], char out[72]) {
    static const char* hex = "0123456789abcdef";
    memcpy(out, "sha256:", 7);
    for (int i = 0; i < 32; ++i) { out[7 + 2 * i] = hex[d[i] >> 4]; out[8 + 2 * i] = hex[d[i] & 15]; }
    out[71] = 0;
}

int mxd_digest_parse(const char* s, uint8_t out[32]) {
    if (!s || strncmp(s, "sha256:", 7) != 0 || strlen(s) != 71) return fail(MXD_ERR_INVALID, "digest: want sha256:<64 lower hex>");
    for (int i = 0; i < 32; ++i) {
        int v = 0;
        for (int k = 0; k < 2; ++k) {
            const char ch = s[7 + 2 * i + k];
            int x = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : -1;
            if (x < 0) return fail(MXD_ERR_INVALID, "digest: invalid hex (go-digest accepts lower case only)");
            v = v * 16 + x;
        }
        if (out) out[i] = (uint8_t)v;
    }
    return MXD_OK;
}

// ---- tree ---------------------------------------------------------------------------------------------
int mxd_tree_shape(uint64_t size, const mxd_tree_params* tp, uint64_t* counts, int max_levels, int* chunk_level) {
    Tree t;
    if (!tree_resolve(tp, &t) || !counts || max_levels < 2) return fail(MXD_ERR_INVALID, "tree: chunk must be leaf * fanout^k (k >= 1), leaf a multiple of 64, fanout >= 2");
    uint64_t n = size ? (size + t.leaf - 1) / t.leaf : 1;
    int lv = 0;
    counts[lv++] = n;
    while (lv <= t.klevel || n > 1) {
        if (lv >= max_levels) return fail(MXD_ERR_INVALID, "tree: too many levels for counts[]");
        n = (n + t.fanout - 1) / t.fanout;
        counts[lv++] = n;
    }
    if (chunk_level) *chunk_level = t.klevel;
    return lv;
}

int mxd_dev_sha256_segments(mxd_ctx* c, int dev, const void* d_data, uint64_t nbytes, uint64_t seg, void* d_out, void* stream) {
    if (!c || dev < 0 || dev >= (int)c->devs.size() || seg == 0 || !d_out) return fail(MXD_ERR_INVALID, "dev_sha256_segments: bad arguments");
    DeviceGuard guard(c->devs[dev]->ordinal);
    return enqueue_segments(c, static_cast<const uint8_t*>(d_data), nbytes, seg, static_cast<uint8_t*>(d_out), (cudaStream_t)stream);
}

int mxd_dev_sha256_batch(mxd_ctx* c, int dev, const mxd_span* d_spans, uint64_t n, void* d_out, void* stream) {
    if (!c || dev < 0 || dev >= (int)c->devs.size() || (n && (!d_spans || !d_out))) return fail(MXD_ERR_INVALID, "dev_sha256_batch: bad arguments");
    if (n == 0) return MXD_OK;
    DeviceGuard guard(c->devs[dev]->ordinal);
    mxd::MsgJob j{};
    j.spans = d_spans; j.nmsg = n; j.out = static_cast<uint8_t*>(d_out); j.finalize = 1; j.one = 1;
    MXD_CUDA(mxd::launch_sha256(j, (cudaStream_t)stream));
    c->launches++;
    return MXD_OK;
}

int mxd_dev_tree_chunks(mxd_ctx* c, int dev, const void* d_piece, uint64_t nbytes, const mxd_tree_params* tp,
                        void* d_chunk_digests, void* stream) {
    Tree t;
    if (!c || dev < 0 || dev >= (int)c->devs.size() || !tree_resolve(tp, &t) || !d_chunk_digests)
        return fail(MXD_ERR_INVALID, "dev_tree_chunks: bad arguments");
    DeviceGuard guard(c->devs[dev]->ordinal);
    return enqueue_tree_chunks(c, t, static_cast<const uint8_t*>(d_piece), nbytes, static_cast<uint8_t*>(d_chunk_digests),
                               (cudaStream_t)stream);
}

int mxd_dev_tree_finish(mxd_ctx* c, int dev, const void* d_chunk_digests, uint64_t nchunks, uint64_t size,
                        const mxd_tree_params* tp, void* d_root, void* stream) {
    Tree t;
    if (!c || dev < 0 || dev >= (int)c->devs.size() || !tree_resolve(tp, &t) || !d_chunk_digests || !d_root || nchunks == 0)
        return fail(MXD_ERR_INVALID, "dev_tree_finish: bad arguments");
    DeviceGuard guard(c->devs[dev]->ordinal);
    return enqueue_tree_finish(c, t, static_cast<const uint8_t*>(d_chunk_digests), nchunks, size,
                               static_cast<uint8_t*>(d_root), (cudaStream_t)stream);
}

int mxd_dev_tree_digest(mxd_ctx* c, int dev, const void* d_data, uint64_t size, const mxd_tree_params* tp,
                        void* d_chunk_digests, void* d_root, void* stream) {
    Tree t;
    if (!c || dev < 0 || dev >= (int)c->devs.size() || !tree_resolve(tp, &t) || !d_root)
        return fail(MXD_ERR_INVALID, "dev_tree_digest: bad arguments");
    DeviceGuard guard(c->devs[dev]->ordinal);
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t nchunks = size ? (size + t.chunk - 1) / t.chunk : 1;
    uint8_t* chunks = static_cast<uint8_t*>(d_chunk_digests);
    uint8_t* owned = nullptr;
    if (!chunks) { MXD_CUDA(cudaMallocAsync(&owned, nchunks * 32, st)); chunks = owned; }
    int rc = enqueue_tree_chunks(c, t, static_cast<const uint8_t*>(d_data), size, chunks, st);
    if (rc == MXD_OK) rc = enqueue_tree_finish(c, t, chunks, nchunks, size, static_cast<uint8_t*>(d_root), st);
    if (owned) cudaFreeAsync(owned, st);
    return rc;
}

int mxd_dev_compare(mxd_ctx* c, int dev, const void* d_got, const void* d_want, uint64_t n, void* d_ok, void* stream) {
    if (!c || dev < 0 || dev >= (int)c->devs.size() || (n && (!d_got || !d_want || !d_ok))) return fail(MXD_ERR_INVALID, "dev_compare: bad arguments");
    DeviceGuard guard(c->devs[dev]->ordinal);
    MXD_CUDA(mxd::launch_compare(static_cast<const uint8_t*>(d_got), static_cast<const uint8_t*>(d_want), n,
                                 static_cast<uint8_t*>(d_ok), (cudaStream_t)stream));
    if (n) c->launches++;
    return MXD_OK;
}

int mxd_dev_gen_fill(mxd_ctx* c, int dev, void* d_dst, uint64_t offset, uint64_t n, uint64_t seed, void* stream) {
    if (!c || dev < 0 || dev >= (int)c->devs.size() || (n && !d_dst)) return fail(MXD_ERR_INVALID, "dev_gen_fill: bad arguments");
    DeviceGuard guard(c->devs[dev]->ordinal);
    MXD_CUDA(mxd::launch_gen_fill(d_dst, offset, n, seed, (cudaStream_t)stream));
    if (n) c->launches++;
    return MXD_OK;
}

int mxd_tree_chunks(mxd_ctx* c, const void* piece, uint64_t nbytes, const mxd_tree_params* tp, uint8_t* out) {
    Tree t;
    if (!c || !tree_resolve(tp, &t) || !out || (nbytes && !piece)) return fail(MXD_ERR_INVALID, "tree_chunks: bad arguments");
    const uint64_t nchunks = nbytes ? (nbytes + t.chunk - 1) / t.chunk : 1;
    int ord = -1;
    const MemKind kind = classify(piece, &ord);
    if (kind == MemKind::Device) {
        const int di = dev_index_of(c, ord);
        if (di < 0) return fail(MXD_ERR_INVALID, "tree_chunks: data lives on a device this context does not drive");
        DevState* d = c->devs[di];
        std::lock_guard<std::mutex> lk(d->mu);
        DeviceGuard guard(d->ordinal);
        uint8_t* d_chunks = nullptr;
        MXD_CUDA(cudaMallocAsync(&d_chunks, nchunks * 32, d->compute));   // cudaMalloc/cudaFree cost ~6 ms a pair on this box
        int rc = enqueue_tree_chunks(c, t, static_cast<const uint8_t*>(piece), nbytes, d_chunks, d->compute);
        if (rc == MXD_OK) {
            cudaError_t e = cudaMemcpyAsync(out, d_chunks, nchunks * 32, cudaMemcpyDeviceToHost, d->compute);
            if (e == cudaSuccess) e = cudaStreamSynchronize(d->compute);
            if (e != cudaSuccess) rc = fail(MXD_ERR_CUDA, cudaGetErrorString(e));
            c->d2h += nchunks * 32;
        }
        cudaFreeAsync(d_chunks, d->compute);
        return rc;
    }
    Source src; src.mem = static_cast<const uint8_t*>(piece); src.pinned = (kind == MemKind::Pinned);
    return host_tree_chunks_all(c, t, src, nbytes, out);
}

int mxd_tree_finish(mxd_ctx* c, const uint8_t* chunk_digests, uint64_t nchunks, uint64_t size, const mxd_tree_params* tp,
                    uint8_t root[32]) {
    Tree t;
    if (!c || !tree_resolve(tp, &t) || !chunk_digests || !root || nchunks == 0) return fail(MXD_ERR_INVALID, "tree_finish: bad arguments");
    const uint64_t expect = size ? (size + t.chunk - 1) / t.chunk : 1;
    if (expect != nchunks) return fail(MXD_ERR_INVALID, "tree_finish: nchunks does not match size/chunk");
    return host_tree_finish(c, c->devs[0], t, chunk_digests, nchunks, size, root);
}

int mxd_tree_digest(mxd_ctx* c, const void* data, uint64_t size, const mxd_tree_params* tp, uint8_t* chunk_digests,
                    uint64_t* nchunks_out, uint8_t root[32]) {
    Tree t;
    if (!c || !tree_resolve(tp, &t) || !root || (size && !data)) return fail(MXD_ERR_INVALID, "tree_digest: bad arguments");
    const uint64_t nchunks = size ? (size + t.chunk - 1) / t.chunk : 1;
    std::vector<uint8_t> tmp;
    uint8_t* chunks = chunk_digests;
    if (!chunks) { tmp.resize(nchunks * 32); chunks = tmp.data(); }
    int rc = mxd_tree_chunks(c, data, size, tp, chunks);
    if (rc != MXD_OK) return rc;
    if (nchunks_out) *nchunks_out = nchunks;
    return host_tree_finish(c, c->devs[0], t, chunks, nchunks, size, root);
}

int mxd_tree_digest_file(mxd_ctx* c, const char* path, const mxd_tree_params* tp, uint8_t* chunk_digests,
                         uint64_t cap_chunks, uint64_t* nchunks_out, uint64_t* size_out, uint8_t root[32]) {
    return mxd_tree_digest_file_tee(c, path, tp, chunk_digests, cap_chunks, nchunks_out, size_out, root, nullptr, nullptr);
}

int mxd_tree_digest_file_tee(mxd_ctx* c, const char* path, const mxd_tree_params* tp, uint8_t* chunk_digests,
                             uint64_t cap_chunks, uint64_t* nchunks_out, uint64_t* size_out, uint8_t root[32],
                             mxd_sink_fn sink, void* user) {
    Tree t;
    if (!c || !path || !tree_resolve(tp, &t) || !root) return fail(MXD_ERR_INVALID, "tree_digest_file: bad arguments");
    int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return fail(MXD_ERR_IO, std::string("open ") + path + ": " + strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) { int e = errno; close(fd); errno = e; return fail(MXD_ERR_IO, std::string("fstat: ") + strerror(e)); }
    const uint64_t size = (uint64_t)st.st_size;
    const uint64_t nchunks = size ? (size + t.chunk - 1) / t.chunk : 1;
    if (size_out) *size_out = size;
    if (nchunks_out) *nchunks_out = nchunks;
    if (chunk_digests && cap_chunks < nchunks) { close(fd); return fail(MXD_ERR_INVALID, "tree_digest_file: chunk_digests too small"); }
    std::vector<uint8_t> tmp;
    uint8_t* chunks = chunk_digests;
    if (!chunks) { tmp.resize(nchunks * 32); chunks = tmp.data(); }
    Source src; src.fd = fd; src.sink = sink; src.sink_user = user;
    int rc = host_tree_chunks_all(c, t, src, size, chunks);
    close(fd);
    if (rc != MXD_OK) return rc;
    return host_tree_finish(c, c->devs[0], t, chunks, nchunks, size, root);
}

// ---- whole-message digests ---------------------------------------------------------------------------
int mxd_sha256_batch(mxd_ctx* c, const mxd_span* spans, uint64_t n, uint8_t* out) {
    if (!c || (n && (!spans || !out))) return fail(MXD_ERR_INVALID, "sha256_batch: bad arguments");
    if (n == 0) return MXD_OK;
    // device-resident spans: one launch, no staging
    int ord = -1, first_ord = -1;
    bool all_dev = true, any_dev = false;
    for (uint64_t i = 0; i < n; ++i) {
        if (spans[i].len && !spans[i].ptr) return fail(MXD_ERR_INVALID, "sha256_batch: null span with non-zero length");
        if (spans[i].len == 0) continue;
        if (classify(spans[i].ptr, &ord) == MemKind::Device) {
            any_dev = true;
            if (first_ord < 0) first_ord = ord; else if (ord != first_ord) return fail(MXD_ERR_INVALID, "sha256_batch: spans on different devices");
        } else all_dev = false;
    }
    if (any_dev && !all_dev) return fail(MXD_ERR_INVALID, "sha256_batch: mixing host and device spans");
    if (any_dev) {
        const int di = dev_index_of(c, first_ord);
        if (di < 0) return fail(MXD_ERR_INVALID, "sha256_batch: data lives on a device this context does not drive");
        DevState* d = c->devs[di];
        std::lock_guard<std::mutex> lk(d->mu);
        DeviceGuard guard(d->ordinal);
        uint8_t* d_buf = nullptr;
        MXD_CUDA(cudaMallocAsync(&d_buf, n * (sizeof(mxd_span) + 32), d->compute));
        int rc = MXD_OK;
        uint64_t total = 0; for (uint64_t i = 0; i < n; ++i) total += spans[i].len;
        cudaError_t e = cudaMemcpyAsync(d_buf, spans, n * sizeof(mxd_span), cudaMemcpyHostToDevice, d->compute);
        if (e == cudaSuccess) {
            mxd::MsgJob j{};
            j.spans = d_buf; j.nmsg = n; j.out = d_buf + n * sizeof(mxd_span); j.finalize = 1; j.one = 1;
            e = mxd::launch_sha256(j, d->compute);
            c->launches++; c->bytes_hashed += total;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_buf + n * sizeof(mxd_span), n * 32, cudaMemcpyDeviceToHost, d->compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d->compute);
        if (e != cudaSuccess) rc = fail(MXD_ERR_CUDA, cudaGetErrorString(e));
        c->d2h += n * 32;
        cudaFreeAsync(d_buf, d->compute);
        return rc;
    }
    LockstepInput in;
    in.src.resize(n); in.len.resize(n);
    for (uint64_t i = 0; i < n; ++i) { in.src[i].mem = static_cast<const uint8_t*>(spans[i].ptr); in.len[i] = spans[i].len; }
    return lockstep_digest_all(c, in, out);
}

int mxd_sha256(mxd_ctx* c, const void* data, uint64_t n, uint8_t out[32]) {
    if (!c || !out || (n && !data)) return fail(MXD_ERR_INVALID, "sha256: bad arguments");
    mxd_span sp{data, n};
    return mxd_sha256_batch(c, &sp, 1, out);
}

int mxd_sha256_files(mxd_ctx* c, const char* const* paths, uint64_t n, uint8_t* out, uint64_t* sizes) {
    if (!c || (n && (!paths || !out))) return fail(MXD_ERR_INVALID, "sha256_files: bad arguments");
    if (n == 0) return MXD_OK;
    LockstepInput in; std::vector<int> fds;
    int rc = open_files(paths, n, &in, &fds);
    if (rc != MXD_OK) return rc;
    if (sizes) for (uint64_t i = 0; i < n; ++i) sizes[i] = in.len[i];
    rc = lockstep_digest_all(c, in, out);
    for (int fd : fds) close(fd);
    return rc;
}

int mxd_sha256_file_parts(mxd_ctx* c, const char* path, const mxd_part* parts, uint64_t n, uint8_t* out) {
    if (!c || !path || (n && (!parts || !out))) return fail(MXD_ERR_INVALID, "sha256_file_parts: bad arguments");
    if (n == 0) return MXD_OK;
    int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return fail(MXD_ERR_IO, std::string("open ") + path + ": " + strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) { int e = errno; close(fd); errno = e; return fail(MXD_ERR_IO, std::string("fstat: ") + strerror(e)); }
    LockstepInput in;
    in.src.resize(n); in.len.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        if (parts[i].offset < 0 || parts[i].length < 0 || (uint64_t)parts[i].offset + (uint64_t)parts[i].length > (uint64_t)st.st_size) {
            close(fd);
            return fail(MXD_ERR_IO, "sha256_file_parts: part " + std::to_string(i) + " lies outside the file");
        }
        in.src[i].fd = fd; in.src[i].base = (uint64_t)parts[i].offset; in.len[i] = (uint64_t)parts[i].length;
    }
    int rc = lockstep_digest_all(c, in, out);
    close(fd);
    return rc;
}

int mxd_sha256_file(mxd_ctx* c, const char* path, uint8_t out[32], uint64_t* size) {
    if (!path) return fail(MXD_ERR_INVALID, "sha256_file: null path");
    const char* paths[1] = {path};
    return mxd_sha256_files(c, paths, 1, out, size);
}

int mxd_verify_batch(mxd_ctx* c, const mxd_span* spans, const uint8_t* want, uint64_t n, uint8_t* ok) {
    if (!c || (n && (!spans || !want || !ok))) return fail(MXD_ERR_INVALID, "verify_batch: bad arguments");
    std::vector<uint8_t> got(n * 32);
    int rc = mxd_sha256_batch(c, spans, n, got.data());
    if (rc != MXD_OK) return rc;
    for (uint64_t i = 0; i < n; ++i) ok[i] = memcmp(&got[32 * i], want + 32 * i, 32) == 0;
    return MXD_OK;
}

int mxd_verify_files(mxd_ctx* c, const char* const* paths, const uint8_t* want, uint64_t n, uint8_t* ok) {
    if (!c || (n && (!paths || !want || !ok))) return fail(MXD_ERR_INVALID, "verify_files: bad arguments");
    std::vector<uint8_t> got(n * 32);
    int rc = mxd_sha256_files(c, paths, n, got.data(), nullptr);
    if (rc != MXD_OK) return rc;
    for (uint64_t i = 0; i < n; ++i) ok[i] = memcmp(&got[32 * i], want + 32 * i, 32) == 0;
    return MXD_OK;
}

// ---- pinned memory -----------------------------------------------------------------------------------
int mxd_host_alloc(mxd_ctx* c, void** out, uint64_t nbytes) {
    if (!c || !out) return fail(MXD_ERR_INVALID, "host_alloc: bad arguments");
    LocalCpuScope numa(c->devs[0]->ordinal);   // place the pages next to the (first) device that will read them
    MXD_CUDA(cudaHostAlloc(out, nbytes, cudaHostAllocPortable));
    return MXD_OK;
}
void mxd_host_free(mxd_ctx*, void* p) { if (p) cudaFreeHost(p); }
int mxd_host_register(mxd_ctx* c, void* p, uint64_t nbytes) {
    if (!c || !p) return fail(MXD_ERR_INVALID, "host_register: bad arguments");
    MXD_CUDA(cudaHostRegister(p, nbytes, cudaHostRegisterPortable));
    return MXD_OK;
}
int mxd_host_unregister(mxd_ctx* c, void* p) {
    if (!c || !p) return fail(MXD_ERR_INVALID, "host_unregister: bad arguments");
    MXD_CUDA(cudaHostUnregister(p));
    return MXD_OK;
}

}  // extern "C"

// ---- incremental hasher (hash.Hash shape, helper.go:46-49) --------------------------------------------
// Writes accumulate in a pinned buffer; every full buffer advances the chain on the GPU (state
// stays in device memory).  Sum hashes the unflushed tail with finalize on a scratch copy, so the
// running state is untouched, as Go's Sum requires.
struct mxd_hasher {
    mxd_ctx* ctx = nullptr;
    DevState* dev = nullptr;
    uint8_t* h_buf = nullptr;  // pinned
    uint8_t* d_buf = nullptr;
    uint32_t* d_state = nullptr;
    uint8_t* d_out = nullptr;
    uint64_t cap = 0, fill = 0, absorbed = 0;
    std::mutex mu;
};

namespace {
constexpr uint64_t kHasherBuf = 4ull << 20;

int hasher_launch(mxd_hasher* h, uint64_t nbytes, int finalize, uint32_t* state, uint8_t* d_out) {
    mxd_ctx* c = h->ctx;
    cudaStream_t st = h->dev->compute;
    if (nbytes) MXD_CUDA(cudaMemcpyAsync(h->d_buf, h->h_buf, nbytes, cudaMemcpyHostToDevice, st));
    mxd::MsgJob j{};
    j.base = h->d_buf; j.nbytes = nbytes; j.seg = nbytes ? nbytes : 64; j.nmsg = 1;
    j.out = d_out; j.state = state; j.prefix_all = h->absorbed; j.finalize = finalize; j.one = 1;
    MXD_CUDA(mxd::launch_sha256(j, st));
    c->launches++; c->bytes_hashed += nbytes; c->h2d += nbytes;
    return MXD_OK;
}
}  // namespace

extern "C" {

int mxd_hasher_new(mxd_ctx* c, mxd_hasher** out) {
    if (!c || !out) return fail(MXD_ERR_INVALID, "hasher_new: bad arguments");
    auto* h = new mxd_hasher();
    h->ctx = c; h->dev = pick_device(c); h->cap = kHasherBuf;
    DeviceGuard guard(h->dev->ordinal);
    cudaError_t e = cudaHostAlloc(&h->h_buf, h->cap, cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_buf, h->cap);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_state, 64);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_out, 32);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_state, kIVHost, 32, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { int rc = fail(MXD_ERR_CUDA, std::string("hasher_new: ") + cudaGetErrorString(e)); mxd_hasher_free(h); return rc; }
    *out = h;
    return MXD_OK;
}

int mxd_hasher_write(mxd_hasher* h, const void* data, uint64_t n) {
    if (!h || (n && !data)) return fail(MXD_ERR_INVALID, "hasher_write: bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev->ordinal);
    const uint8_t* p = static_cast<const uint8_t*>(data);
    while (n) {
        const uint64_t take = std::min(n, h->cap - h->fill);
        memcpy(h->h_buf + h->fill, p, take);
        h->fill += take; p += take; n -= take;
        if (h->fill == h->cap) {
            std::lock_guard<std::mutex> dl(h->dev->mu);
            int rc = hasher_launch(h, h->cap, 0, h->d_state, h->d_out);
            if (rc != MXD_OK) return rc;
            MXD_CUDA(cudaStreamSynchronize(h->dev->compute));  // h_buf is about to be overwritten
            h->absorbed += h->cap; h->fill = 0;
        }
    }
    return MXD_OK;
}

int mxd_hasher_sum(mxd_hasher* h, uint8_t out[32]) {
    if (!h || !out) return fail(MXD_ERR_INVALID, "hasher_sum: bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev->ordinal);
    std::lock_guard<std::mutex> dl(h->dev->mu);
    // finalize reads the running state but does not write it back
    int rc = hasher_launch(h, h->fill, 1, h->d_state, h->d_out);
    if (rc != MXD_OK) return rc;
    MXD_CUDA(cudaMemcpyAsync(out, h->d_out, 32, cudaMemcpyDeviceToHost, h->dev->compute));
    MXD_CUDA(cudaStreamSynchronize(h->dev->compute));
    h->ctx->d2h += 32;
    return MXD_OK;
}

int mxd_hasher_reset(mxd_hasher* h) {
    if (!h) return fail(MXD_ERR_INVALID, "hasher_reset: null");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev->ordinal);
    MXD_CUDA(cudaMemcpy(h->d_state, kIVHost, 32, cudaMemcpyHostToDevice));
    h->fill = 0; h->absorbed = 0;
    return MXD_OK;
}

uint64_t mxd_hasher_size(const mxd_hasher* h) { return h ? h->absorbed + h->fill : 0; }

void mxd_hasher_free(mxd_hasher* h) {
    if (!h) return;
    if (h->dev) { DeviceGuard guard(h->dev->ordinal);
        if (h->h_buf) cudaFreeHost(h->h_buf);
        if (h->d_buf) cudaFree(h->d_buf);
        if (h->d_state) cudaFree(h->d_state);
        if (h->d_out) cudaFree(h->d_out);
    }
    delete h;
}

}  // extern "C"
