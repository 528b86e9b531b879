// Incremental hasher with the shape of Go's hash.Hash, the seam pkg/client/helper.go:46-49 needs
// (digest.Canonical.Digester() fed by io.MultiWriter while a tar.gz is produced).
//
// Writes accumulate in one of two pinned buffers; a full buffer is handed to the GPU (H2D + one chained launch on
// the hasher's own stream, the chain state stays in device memory) while the caller fills the other one, so Write
// only waits when it is two buffers ahead of the chain.  Sum hashes the unflushed tail with finalize on a launch
// that reads the running state without writing it back, as hash.Hash.Sum requires.  A single stream is a single
// SHA-256 chain (~0.09 GB/s on one GPU lane): this exists for API completeness and for callers whose producer is
// slower than that (gzip); INTEGRATION.md says when to keep the CPU hasher.
#include "mxd_core.h"

using namespace mxdi;

struct mxd_hasher {
    Core* core = nullptr;
    DevState* dev = nullptr;
    cudaStream_t st = nullptr;
    uint8_t* h_buf[2] = {nullptr, nullptr};   // pinned
    uint8_t* d_buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool busy[2] = {false, false};
    int cur = 0;
    uint32_t* d_state = nullptr;
    uint8_t* d_out = nullptr;
    uint64_t cap = 0, fill = 0, absorbed = 0;
    std::mutex mu;
};

namespace {
constexpr uint64_t kHasherBuf = 4ull << 20;

// H2D of buffer `b` (nbytes) + one chained launch.  finalize: pad and emit into d_out, state untouched.
int hasher_launch(mxd_hasher* h, int b, uint64_t nbytes, int finalize) {
    if (nbytes) MXD_CUDA(cudaMemcpyAsync(h->d_buf[b], h->h_buf[b], nbytes, cudaMemcpyHostToDevice, h->st));
    mxd::MsgJob j{};
    j.base = h->d_buf[b]; j.nbytes = nbytes; j.seg = nbytes ? nbytes : 64; j.nmsg = 1;
    j.out = h->d_out; j.state = h->d_state; j.prefix_all = h->absorbed; j.finalize = finalize; j.one = 1;
    MXD_CUDA(mxd::launch_sha256(j, h->st));
    h->core->launches++; h->core->bytes_hashed += nbytes; h->core->h2d += nbytes;
    return MXD_OK;
}
}  // namespace

extern "C" {

int mxd_hasher_new(mxd_ctx* c, mxd_hasher** out) {
    if (!c || !c->core || !out) return fail(MXD_ERR_INVALID, "hasher_new: bad arguments");
    auto* h = new mxd_hasher();
    h->core = c->core; h->dev = pick_device(c->core); h->cap = kHasherBuf;
    DeviceGuard guard(h->dev->ordinal);
    cudaError_t e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        e = cudaHostAlloc(&h->h_buf[b], h->cap, cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_buf[b], h->cap);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev[b], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMalloc(&h->d_state, 64);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_out, 32);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_state, kIVHost, 32, cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
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
        memcpy(h->h_buf[h->cur] + h->fill, p, take);
        h->fill += take; p += take; n -= take;
        if (h->fill == h->cap) {
            int rc = hasher_launch(h, h->cur, h->cap, 0);
            if (rc != MXD_OK) return rc;
            MXD_CUDA(cudaEventRecord(h->ev[h->cur], h->st));
            h->busy[h->cur] = true;
            h->absorbed += h->cap; h->fill = 0;
            h->cur ^= 1;
            if (h->busy[h->cur]) {               // the other buffer is still being copied / hashed: wait for it, not for this one
                MXD_CUDA(cudaEventSynchronize(h->ev[h->cur]));
                h->busy[h->cur] = false;
            }
        }
    }
    return MXD_OK;
}

int mxd_hasher_sum(mxd_hasher* h, uint8_t out[32]) {
    if (!h || !out) return fail(MXD_ERR_INVALID, "hasher_sum: bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev->ordinal);
    // finalize reads the running state but does not write it back; the tail stays in the pinned buffer
    int rc = hasher_launch(h, h->cur, h->fill, 1);
    if (rc != MXD_OK) return rc;
    MXD_CUDA(cudaMemcpyAsync(out, h->d_out, 32, cudaMemcpyDeviceToHost, h->st));
    MXD_CUDA(cudaStreamSynchronize(h->st));
    h->busy[0] = h->busy[1] = false;
    h->core->d2h += 32;
    return MXD_OK;
}

int mxd_hasher_reset(mxd_hasher* h) {
    if (!h) return fail(MXD_ERR_INVALID, "hasher_reset: null");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev->ordinal);
    MXD_CUDA(cudaStreamSynchronize(h->st));
    MXD_CUDA(cudaMemcpyAsync(h->d_state, kIVHost, 32, cudaMemcpyHostToDevice, h->st));
    MXD_CUDA(cudaStreamSynchronize(h->st));
    h->busy[0] = h->busy[1] = false;
    h->fill = 0; h->absorbed = 0;
    return MXD_OK;
}

uint64_t mxd_hasher_size(const mxd_hasher*) { return 32; }          // hash.Hash.Size()
uint64_t mxd_hasher_block_size(const mxd_hasher*) { return 64; }    // hash.Hash.BlockSize()
uint64_t mxd_hasher_written(const mxd_hasher* h) { return h ? h->absorbed + h->fill : 0; }

void mxd_hasher_free(mxd_hasher* h) {
    if (!h) return;
    if (h->dev) {
        DeviceGuard guard(h->dev->ordinal);
        if (h->st) cudaStreamSynchronize(h->st);
        for (int b = 0; b < 2; ++b) {
            if (h->h_buf[b]) cudaFreeHost(h->h_buf[b]);
            if (h->d_buf[b]) cudaFree(h->d_buf[b]);
            if (h->ev[b]) cudaEventDestroy(h->ev[b]);
        }
        if (h->d_state) cudaFree(h->d_state);
        if (h->d_out) cudaFree(h->d_out);
        if (h->st) cudaStreamDestroy(h->st);
    }
    delete h;
}

}  // extern "C"
