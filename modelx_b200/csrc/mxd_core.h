// Internal structures shared by the translation units behind include/modelx_digest.h
// (mxd_api.cu: lifecycle, tree digests, device forms; mxd_lockstep.cu: the whole-message digest
// service; mxd_hasher.cu: the hash.Hash-shaped incremental hasher).  Nothing here is ABI.
#pragma once
#include "../../include/modelx_digest.h"
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <sched.h>
#include <string>
#include <thread>
#include <vector>

namespace mxdi {

int fail(int status, const std::string& msg);          // records the thread-local detail, returns status
const std::string& last_error();

#define MXD_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return ::mxdi::fail(MXD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));  \
    } while (0)

constexpr int kSlots = 4;
extern const uint32_t kIVHost[8];

bool device_local_cpus(int ordinal, cpu_set_t* set);

// Scoped: run the enclosed allocations (first touch + pin) / thread creations on the device's local CPUs.
struct LocalCpuScope {
    cpu_set_t old; bool active = false;
    explicit LocalCpuScope(int ordinal);
    ~LocalCpuScope();
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int ordinal) { cudaGetDevice(&prev); cudaSetDevice(ordinal); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Small persistent worker pool that fills pinned ring slots (pread / memcpy) in parallel: one
// thread reads the page cache at 2-4 GB/s, far below the 55 GB/s a PCIe Gen5 x16 link moves.
class StagePool {
public:
    explicit StagePool(int nthreads);
    ~StagePool();
    // run fn(i) for i in [0, n) on the pool plus the calling thread; returns when all are done
    void parallel_for(int n, const std::function<void(int)>& fn);
    int width() const { return (int)workers_.size() + 1; }
private:
    struct Batch { const std::function<void(int)>* fn; int n; std::atomic<int> next, done; int active = 0; };
    void run();
    std::vector<std::thread> workers_;
    std::vector<Batch*> queue_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    bool stop_ = false;
};

// Where streamed bytes come from: host memory or a byte range of an open file; optionally teed to a sink.
struct Source {
    const uint8_t* mem = nullptr;
    bool pinned = false;
    int fd = -1;
    uint64_t base = 0;             // offset of this source's byte 0 inside the file
    mxd_sink_fn sink = nullptr;    // every streamed byte is also handed to this callback, once
    void* sink_user = nullptr;
    uint64_t sink_base = 0;        // logical offset of this source's byte 0 for the sink
    const uint8_t* map = nullptr;  // file sources: read-only mapping of the source's byte 0 onwards (null: pread)
};

// host/stage_copy.cpp: the staging copy (streaming stores) and file mappings with a SIGBUS guard
void stage_copy(uint8_t* dst, const uint8_t* src, size_t n);
int stage_copy_mapped(uint8_t* dst, const uint8_t* src, size_t n);      // -1: the mapping faulted (file shrank)
void* file_map(int fd, uint64_t base, uint64_t nbytes, const uint8_t** map, uint64_t* handle_len);
void file_unmap(void* handle, uint64_t handle_len);
struct FileMapGuard {     // scoped mapping of a Source's file range
    void* h = nullptr; uint64_t len = 0;
    void attach(Source* s, uint64_t nbytes) { if (s->fd >= 0 && !s->map) h = file_map(s->fd, s->base, nbytes, &s->map, &len); }
    ~FileMapGuard() { file_unmap(h, len); }
};

struct LaneService;

struct DevState {
    int ordinal = -1;
    cudaStream_t compute = nullptr, copy = nullptr;
    uint64_t slot_bytes = 0;
    uint8_t* h_ring = nullptr;  // kSlots * slot_bytes, pinned
    uint8_t* d_ring = nullptr;  // kSlots * slot_bytes
    cudaEvent_t ev_copied[kSlots] = {}, ev_done[kSlots] = {};
    std::mutex mu;               // owner of the ring: one streaming operation (or the digest service) at a time
    StagePool* pool = nullptr;   // slot fillers for this device
    std::mutex svc_mu;           // guards lazy creation of svc
    LaneService* svc = nullptr;  // whole-message digest service (mxd_lockstep.cu), created on first use
};

// What mxd_open creates.  The root handle and every operation handle (mxd_op_begin) point at one Core.
struct Core {
    std::vector<DevState*> devs;
    std::atomic<uint64_t> launches{0}, bytes_hashed{0}, h2d{0}, d2h{0}, src_read{0};
    std::atomic<uint64_t> cancel_gen{0};   // bumped by mxd_cancel(root): aborts every call in flight at that moment
    std::atomic<uint32_t> rr{0};           // round-robin device pick for single-device calls
    std::atomic<int> open_fds{0};          // files the digest service holds open right now
    int fd_cap = 256;                      // bound on open_fds, from RLIMIT_NOFILE
    std::atomic<int> live_ops{0};
    // live timing of leaf-level launches (mxd_prof_*)
    std::atomic<int> prof_on{0};
    std::mutex prof_mu;
    struct ProfRec { cudaEvent_t a, b; uint64_t bytes; int ordinal; };
    std::vector<ProfRec> prof;
    // slot timeline (mxd_trace_*): host-side record of what the ring did, for overlap evidence without nsys
    std::atomic<int> trace_on{0};
    std::mutex trace_mu;
    struct TraceRec { cudaEvent_t c0, c1, k0, k1; uint64_t bytes; int ordinal, slot; double fill_ms; };
    std::vector<TraceRec> trace;
    cudaEvent_t trace_origin = nullptr; int trace_origin_ord = -1;
};

}  // namespace mxdi

// The public opaque handle: the root (parent == nullptr) or one operation on it.
struct mxd_ctx {
    mxdi::Core* core = nullptr;
    mxd_ctx* parent = nullptr;
    std::atomic<int> canceled{0};          // operation handles: sticky for the life of the operation
};

namespace mxdi {

// Cancellation scope of one ABI call: captured at entry, polled by the streaming loops.
struct CancelScope {
    const mxd_ctx* h; uint64_t gen;
    explicit CancelScope(const mxd_ctx* handle) : h(handle), gen(handle->core->cancel_gen.load()) {}
    bool canceled() const { return h->canceled.load(std::memory_order_relaxed) != 0 || h->core->cancel_gen.load(std::memory_order_relaxed) != gen; }
};

DevState* pick_device(Core* c);
int dev_index_of(const Core* c, int ordinal);

enum class MemKind { Pageable, Pinned, Device };
MemKind classify(const void* p, int* device_ordinal);

// Fill `n` bytes at logical offset `off` of the source into pinned `dst`; *from = the pointer the H2D copy should
// read (dst, or the caller's own memory when that is already pinned).  Large fills are split across the pool.
int source_stage(Core* c, const Source& s, uint64_t off, uint64_t n, uint8_t* dst, const uint8_t** from, StagePool* pool);
int sink_pieces(const Source& s, uint64_t off, uint64_t n, const uint8_t* data, StagePool* pool);

// ---- whole-message digest service (mxd_lockstep.cu) --------------------------------------------------------------
// One pass over one source feeding one or more SHA-256 chains (byte ranges of that source).
struct RangeReq { uint64_t off, len; uint8_t* out; };
struct StreamReq {
    const char* path = nullptr;        // file source: opened by the service when the stream is admitted ...
    const uint8_t* mem = nullptr;      // ... or host memory
    uint64_t size = 0;                 // bytes of the source (files: filled in by svc_run from fstat)
    std::vector<RangeReq> ranges;      // chains; an empty list means one chain over [0, size)
    uint8_t* whole_out = nullptr;      // digest of [0, size) when `ranges` is empty
    mxd_sink_fn sink = nullptr; void* sink_user = nullptr;
    int status = MXD_OK;               // per-stream result
    std::string error;
};
// Runs all streams (spread over the context's devices, coalesced with whatever other callers have in flight) and
// waits for them.  Returns MXD_OK or the first failing stream's status; per-stream status is left in the requests.
int svc_run(mxd_ctx* h, std::vector<StreamReq>& reqs);
void svc_destroy(DevState* d);

}  // namespace mxdi
