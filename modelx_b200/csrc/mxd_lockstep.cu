// The whole-message digest service: the reference's one-SHA-256-per-file semantics
// (pkg/client/push.go:149-161, pull.go:115-123) for any number of files / buffers at once.
//
// One SHA-256 message is one serial chain, so a single digest can use one GPU lane (~0.09 GB/s against ~1.4 GB/s
// on a SHA-NI core); the GPU only wins through width.  Every device therefore runs ONE service thread that
// advances all messages currently in flight together, in rounds: round k stages the next S bytes of every running
// stream into one pinned ring slot, copies the slot to the device and launches one kernel in which every chain
// absorbs its share; chain states (8 words) stay in device memory between rounds.  Callers only enqueue and wait:
//   - concurrent callers coalesce: the 3 goroutines of PullPushConcurrency (push.go:27) calling mxd_sha256_file
//     at the same time become 3 lanes of the same rounds instead of three launches one after another; a stream
//     joins at the next round boundary and leaves when it is done, releasing its share of the slot;
//   - cancellation is per call (push.go:150-159): a canceled stream leaves at the next round, the others go on;
//   - at most Core::fd_cap files are open at once and at most lane_cap chains run at once; the rest wait in the
//     queue, so batches of any size work (the reference holds 3 files open);
//   - one stream may feed several chains: SHA-256 of any byte ranges of a file (the whole file and each multipart
//     part, extension_s3.go:99-112) in ONE pass over it, optionally teed to a sink (an uploader / store writer).
// No hashing happens on the CPU here.
#include "mxd_core.h"

#include <cerrno>
#include <deque>
#include <fcntl.h>
#include <memory>
#include <sys/stat.h>
#include <unistd.h>

namespace mxdi {

namespace {

constexpr uint32_t kLaneCap = 16384;          // chains per device at once
constexpr uint64_t kMaxShare = 8ull << 20;    // bytes one stream advances per round at most

struct Group {                                // one svc_run call
    std::mutex mu; std::condition_variable cv; uint64_t remaining = 0;
};

struct Chain {
    uint64_t start = 0, end = 0, cur = 0;     // byte range of the stream; next byte to absorb
    uint8_t* out = nullptr;
    int lane = -1;
    bool fresh = true, done = false;
};

struct Stream {
    StreamReq* req = nullptr;
    Group* group = nullptr;
    CancelScope cs;
    Source src;
    int fd = -1;
    uint64_t size = 0;
    FileMapGuard fmap;                        // read-only mapping of the file (null: pread)
    std::vector<Chain> chains;
    uint64_t pos = 0;                         // bytes below pos have been staged and handed to the sink
    int unfinished = 0;                       // chains that have not had their final round yet
    int pending_retire = 0;                   // chains finalized in rounds whose digests are not back yet
    int status = MXD_OK; std::string error;
    bool active = false, completed = false;
    // per-round plan
    uint64_t a = 0, b = 0, slot_off = 0; uint32_t d0 = 0, d1 = 0;
    std::atomic<int> fill_err{0};
    explicit Stream(const mxd_ctx* h) : cs(h) {}
    bool wants_more() const { return unfinished > 0 || (src.sink != nullptr && pos < size); }
};

}  // namespace

struct LaneService {
    Core* core; DevState* d;
    std::thread worker;
    std::mutex mu; std::condition_variable cv;
    std::deque<Stream*> queue;
    bool stop = false;
    std::atomic<uint64_t> pending_bytes{0};

    // ---- worker-private -----------------------------------------------------------------------
    std::vector<Stream*> active;
    std::vector<int> free_lanes;
    uint32_t lane_cap = kLaneCap, max_streams = kLaneCap;
    mxd::LaneDesc* h_desc = nullptr; mxd::LaneDesc* d_desc = nullptr;     // kSlots * lane_cap
    uint32_t* d_state = nullptr;                                           // lane_cap * 8
    uint8_t* d_fin = nullptr; uint8_t* h_fin = nullptr;                    // kSlots * lane_cap * 32
    struct Fin { Stream* s; uint32_t chain; };
    std::vector<Fin> slotfin[kSlots];
    bool slot_busy[kSlots] = {false, false, false, false};
    uint64_t round = 0;
    bool holding_ring = false;
    std::string init_error;

    LaneService(Core* c, DevState* dev) : core(c), d(dev) {
        max_streams = (uint32_t)std::min<uint64_t>(lane_cap, d->slot_bytes / 256);
        DeviceGuard guard(d->ordinal);
        LocalCpuScope numa(d->ordinal);
        cudaError_t e = cudaHostAlloc(&h_desc, sizeof(mxd::LaneDesc) * kSlots * lane_cap, cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaHostAlloc(&h_fin, 32ull * kSlots * lane_cap, cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaMalloc(&d_desc, sizeof(mxd::LaneDesc) * kSlots * lane_cap);
        if (e == cudaSuccess) e = cudaMalloc(&d_state, 32ull * lane_cap);
        if (e == cudaSuccess) e = cudaMalloc(&d_fin, 32ull * kSlots * lane_cap);
        if (e != cudaSuccess) init_error = std::string("digest service: ") + cudaGetErrorString(e);
        for (int i = (int)lane_cap - 1; i >= 0; --i) free_lanes.push_back(i);
        worker = std::thread([this] { run(); });
    }
    ~LaneService() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        if (worker.joinable()) worker.join();
        DeviceGuard guard(d->ordinal);
        if (h_desc) cudaFreeHost(h_desc);
        if (h_fin) cudaFreeHost(h_fin);
        if (d_desc) cudaFree(d_desc);
        if (d_state) cudaFree(d_state);
        if (d_fin) cudaFree(d_fin);
    }

    void submit(Stream* s) {
        pending_bytes += s->size + 1;
        { std::lock_guard<std::mutex> lk(mu); queue.push_back(s); }
        cv.notify_all();
    }

    // ---- completion ---------------------------------------------------------------------------
    void maybe_complete(Stream* s) {
        if (s->completed || s->active || s->pending_retire > 0) return;
        s->completed = true;
        if (s->fd >= 0) { close(s->fd); s->fd = -1; core->open_fds--; }
        pending_bytes -= s->size + 1;
        s->req->status = s->status; s->req->error = s->error; s->req->size = s->size;
        Group* g = s->group;                   // after the decrement below the caller may free s
        std::lock_guard<std::mutex> lk(g->mu);
        g->remaining--;
        g->cv.notify_all();
    }
    void fail_stream(Stream* s, int rc, const std::string& msg) {
        if (s->status == MXD_OK) { s->status = rc; s->error = msg; }
        for (auto& ch : s->chains) if (!ch.done && ch.lane >= 0) { free_lanes.push_back(ch.lane); ch.lane = -1; ch.done = true; }
        s->unfinished = 0;
        s->active = false;
        maybe_complete(s);
    }
    void retire(int slot) {
        for (size_t i = 0; i < slotfin[slot].size(); ++i) {
            Stream* s = slotfin[slot][i].s;
            if (s->status == MXD_OK) memcpy(s->chains[slotfin[slot][i].chain].out, h_fin + 32ull * ((uint64_t)slot * lane_cap + i), 32);
            s->pending_retire--;
            maybe_complete(s);
        }
        slotfin[slot].clear();
        slot_busy[slot] = false;
    }
    bool drain() {           // wait for every round in flight, oldest first
        bool ok = true;
        for (uint64_t k = 0; k < kSlots; ++k) {
            const int s = (int)((round + k) % kSlots);
            if (!slot_busy[s]) continue;
            if (cudaEventSynchronize(d->ev_done[s]) != cudaSuccess) ok = false;
            retire(s);
        }
        return ok;
    }

    // ---- admission ----------------------------------------------------------------------------
    // Opens the file, resolves the chains and hands out lanes.  Returns false when resources (lanes, fds) are short
    // and the stream has to stay queued.
    bool admit(Stream* s) {
        StreamReq* rq = s->req;
        const size_t want_lanes = rq->ranges.empty() ? 1 : rq->ranges.size();
        if (want_lanes > lane_cap) { s->status = MXD_ERR_INVALID; s->error = "more ranges than chains per device (" + std::to_string(lane_cap) + ")"; finish_unadmitted(s); return true; }
        if (free_lanes.size() < want_lanes || active.size() >= max_streams) return false;
        if (s->cs.canceled()) { s->status = MXD_ERR_CANCELED; s->error = "canceled"; finish_unadmitted(s); return true; }
        if (rq->path) {
            int cur = core->open_fds.load();           // reserve a slot of the open-file budget (never overshoots, even transiently)
            do { if (cur >= core->fd_cap) return false; } while (!core->open_fds.compare_exchange_weak(cur, cur + 1));
            int fd = open(rq->path, O_RDONLY | O_CLOEXEC);
            struct stat st;
            if (fd < 0 || fstat(fd, &st) != 0) {
                const int e = errno;
                if (fd >= 0) close(fd);
                core->open_fds--;
                s->status = MXD_ERR_IO; s->error = std::string(fd < 0 ? "open " : "fstat ") + rq->path + ": " + strerror(e);
                finish_unadmitted(s); return true;
            }
            if (S_ISDIR(st.st_mode)) {      // os.Open succeeds on a directory, the read then fails with EISDIR (pull.go:116-119)
                close(fd); core->open_fds--;
                s->status = MXD_ERR_IO; s->error = std::string("read ") + rq->path + ": is a directory";
                finish_unadmitted(s); return true;
            }
            s->fd = fd; s->src.fd = fd;
            pending_bytes -= s->size; s->size = (uint64_t)st.st_size; pending_bytes += s->size;
            s->fmap.attach(&s->src, s->size);
        } else {
            s->src.mem = rq->mem;
        }
        s->src.sink = rq->sink; s->src.sink_user = rq->sink_user;
        if (rq->ranges.empty()) {
            Chain ch; ch.start = 0; ch.end = s->size; ch.cur = 0; ch.out = rq->whole_out;
            s->chains.push_back(ch);
        } else {
            for (auto& r : rq->ranges) {
                if (r.off > s->size || r.len > s->size - r.off) {
                    s->status = MXD_ERR_IO; s->error = "range [" + std::to_string(r.off) + ", +" + std::to_string(r.len) + ") lies outside the " + std::to_string(s->size) + "-byte source";
                    finish_unadmitted(s); return true;
                }
                Chain ch; ch.start = r.off; ch.end = r.off + r.len; ch.cur = r.off; ch.out = r.out;
                s->chains.push_back(ch);
            }
        }
        for (auto& ch : s->chains) { ch.lane = free_lanes.back(); free_lanes.pop_back(); }
        s->unfinished = (int)s->chains.size();
        s->active = true;
        active.push_back(s);
        return true;
    }
    void finish_unadmitted(Stream* s) {
        if (s->fd >= 0) { close(s->fd); s->fd = -1; core->open_fds--; }
        s->active = false;
        maybe_complete(s);
    }

    // ---- one round ------------------------------------------------------------------------------
    int round_once() {
        const int slot = (int)(round % kSlots);
        if (slot_busy[slot]) {
            MXD_CUDA(cudaEventSynchronize(d->ev_done[slot]));
            retire(slot);
        }
        // NB: a stream may be freed by its caller as soon as maybe_complete() has run, so it is dropped from `active`
        // at the very point it can complete and never dereferenced afterwards.
        {
            std::vector<Stream*> keep;
            for (Stream* s : active) { if (s->cs.canceled()) fail_stream(s, MXD_ERR_CANCELED, "canceled"); else keep.push_back(s); }
            active.swap(keep);
        }
        if (active.empty()) return MXD_OK;

        const uint64_t n = active.size();
        uint64_t share = ((d->slot_bytes - 16 * n) / n) & ~63ull;
        share = std::min<uint64_t>(share, kMaxShare);
        uint8_t* h_slot = d->h_ring + (uint64_t)slot * d->slot_bytes;
        uint8_t* d_slot = d->d_ring + (uint64_t)slot * d->slot_bytes;
        mxd::LaneDesc* hd = h_desc + (uint64_t)slot * lane_cap;
        uint32_t ndesc = 0, nfin = 0;
        uint64_t used = 0, hashed = 0;

        struct Fill { Stream* s; uint64_t off, n, dst; };      // source bytes [off, off+n) -> h_slot + dst
        std::vector<Fill> fills;
        constexpr uint64_t kPiece = 1ull << 20;
        for (Stream* s : active) {
            // window: from the lowest byte any running chain (or the sink) still needs, one share long
            uint64_t a = ~0ull;
            for (auto& ch : s->chains) if (!ch.done && ch.cur < ch.end) a = std::min(a, ch.cur);
            if (s->src.sink && s->pos < s->size) a = std::min(a, s->pos);
            if (a == ~0ull) a = s->size;                            // only empty chains left
            const uint64_t b = std::min(s->size, a + share);
            s->a = a; s->b = b; s->slot_off = used; s->d0 = ndesc;
            for (uint32_t ci = 0; ci < s->chains.size(); ++ci) {
                Chain& ch = s->chains[ci];
                if (ch.done) continue;
                const bool empty_left = ch.cur == ch.end;
                if (!empty_left && (ch.cur < a || ch.cur >= b)) continue;
                const uint64_t stop_at = empty_left ? ch.cur : std::min(ch.end, b);
                uint64_t take = stop_at - ch.cur;
                const bool final = stop_at == ch.end;
                if (!final) take &= ~63ull;
                if (take == 0 && !final) continue;
                mxd::LaneDesc& ld = hd[ndesc++];
                ld.ptr = d_slot + used + (empty_left ? 0 : ch.cur - a);
                ld.len = take; ld.prefix = ch.cur - ch.start; ld.lane = (uint32_t)ch.lane; ld.oidx = final ? nfin : 0;
                ld.ctl = (final ? mxd::kFinalize : 0u) | (ch.fresh ? mxd::kFresh : 0u); ld.pad = 0;
                ch.cur += take; ch.fresh = false; hashed += take;
                if (final) {
                    slotfin[slot].push_back({s, ci}); ++nfin;
                    ch.done = true; s->unfinished--; s->pending_retire++;
                    free_lanes.push_back(ch.lane);               // stream order protects the state slot: a later round re-seeds it (kFresh)
                }
            }
            s->d1 = ndesc;
            if (b > a) {
                for (uint64_t p = 0; p < b - a; p += kPiece) fills.push_back({s, a + p, std::min(kPiece, b - a - p), used + p});
                used += (b - a + 15) & ~15ull;
            }
        }

        // gather this round's bytes of every stream into the slot (files read in parallel) and feed the sinks
        std::vector<std::string> ferr(fills.size());
        d->pool->parallel_for((int)fills.size(), [&](int f) {
            const Fill& fl = fills[f];
            Stream* s = fl.s;
            uint8_t* dst = h_slot + fl.dst;
            if (s->src.map) {
                if (stage_copy_mapped(dst, s->src.map + fl.off, fl.n) != 0) { ferr[f] = "read: file shrank while hashing"; s->fill_err.store(1); return; }
            } else if (s->fd >= 0) {
                uint64_t got = 0;
                while (got < fl.n) {
                    ssize_t r = pread(s->fd, dst + got, fl.n - got, (off_t)(fl.off + got));
                    if (r < 0) { if (errno == EINTR) continue; ferr[f] = std::string("pread: ") + strerror(errno); s->fill_err.store(1); return; }
                    if (r == 0) { ferr[f] = "pread: file shrank while hashing"; s->fill_err.store(1); return; }
                    got += (uint64_t)r;
                }
            } else {
                stage_copy(dst, s->src.mem + fl.off, fl.n);
            }
            if (s->src.sink) {                                  // hand each byte to the tee exactly once
                const uint64_t lo = std::max(fl.off, s->pos), hi = fl.off + fl.n;
                if (hi > lo && s->src.sink(s->src.sink_user, lo, dst + (lo - fl.off), hi - lo) != 0) { ferr[f] = "sink refused data"; s->fill_err.store(1); }
            }
        });
        uint64_t staged = 0;
        for (auto& fl : fills) staged += fl.n;
        core->src_read += staged;
        {
            std::vector<Stream*> failed;
            for (size_t f = 0; f < fills.size(); ++f) {
                Stream* s = fills[f].s;
                if (ferr[f].empty() || std::find(failed.begin(), failed.end(), s) != failed.end()) continue;
                failed.push_back(s);
                for (uint32_t k = s->d0; k < s->d1; ++k) hd[k].ctl = mxd::kSkip;     // its chains sit this launch out
                s->status = MXD_ERR_IO; s->error = ferr[f];
            }
            if (!failed.empty()) {
                std::vector<Stream*> keep;
                for (Stream* s : active) if (std::find(failed.begin(), failed.end(), s) == failed.end()) keep.push_back(s);
                active.swap(keep);
                for (Stream* s : failed) fail_stream(s, MXD_ERR_IO, s->error);
            }
        }
        for (Stream* s : active) s->pos = std::max(s->pos, s->b);

        if (ndesc) {
            if (used) MXD_CUDA(cudaMemcpyAsync(d_slot, h_slot, used, cudaMemcpyHostToDevice, d->copy));
            mxd::LaneDesc* dd = d_desc + (uint64_t)slot * lane_cap;
            MXD_CUDA(cudaMemcpyAsync(dd, hd, sizeof(mxd::LaneDesc) * ndesc, cudaMemcpyHostToDevice, d->copy));
            MXD_CUDA(cudaEventRecord(d->ev_copied[slot], d->copy));
            MXD_CUDA(cudaStreamWaitEvent(d->compute, d->ev_copied[slot], 0));
            mxd::MsgJob j{};
            j.descs = dd; j.nmsg = ndesc; j.state = d_state; j.out = d_fin + 32ull * slot * lane_cap; j.one = 1;
            MXD_CUDA(mxd::launch_sha256(j, d->compute));
            if (nfin) MXD_CUDA(cudaMemcpyAsync(h_fin + 32ull * slot * lane_cap, j.out, 32ull * nfin, cudaMemcpyDeviceToHost, d->compute));
            MXD_CUDA(cudaEventRecord(d->ev_done[slot], d->compute));
            slot_busy[slot] = true;
            core->launches++; core->bytes_hashed += hashed;
            core->h2d += used + sizeof(mxd::LaneDesc) * ndesc; core->d2h += 32ull * nfin;
        }
        ++round;

        {
            std::vector<Stream*> keep;
            for (Stream* s : active) { if (s->wants_more()) keep.push_back(s); else { s->active = false; maybe_complete(s); } }
            active.swap(keep);
        }
        // digests of earlier rounds that have already landed: hand them back now rather than when their slot comes round
        for (uint64_t k = 0; k < kSlots; ++k) {
            const int s2 = (int)((round + k) % kSlots);
            if (!slot_busy[s2]) continue;
            if (cudaEventQuery(d->ev_done[s2]) != cudaSuccess) { cudaGetLastError(); break; }
            retire(s2);
        }
        return MXD_OK;
    }

    void run() {
        cudaSetDevice(d->ordinal);
        for (;;) {
            if (active.empty()) {            // idle: hand back every digest still in flight and release the ring before sleeping
                drain();
                if (holding_ring) { d->mu.unlock(); holding_ring = false; }
            }
            {
                std::unique_lock<std::mutex> lk(mu);
                if (active.empty()) cv.wait(lk, [&] { return stop || !queue.empty(); });
                if (stop) {
                    while (!queue.empty()) { Stream* s = queue.front(); queue.pop_front(); lk.unlock(); s->status = MXD_ERR_CANCELED; s->error = "engine closed"; finish_unadmitted(s); lk.lock(); }
                    if (active.empty()) break;
                }
                while (!queue.empty()) {
                    Stream* s = queue.front();
                    lk.unlock();
                    bool taken;
                    if (!init_error.empty()) { s->status = MXD_ERR_CUDA; s->error = init_error; finish_unadmitted(s); taken = true; }
                    else taken = admit(s);
                    lk.lock();
                    if (!taken) break;
                    queue.pop_front();        // only this thread pops, so the front is still s
                }
            }
            if (active.empty()) {
                bool waiting;
                { std::lock_guard<std::mutex> lk(mu); waiting = !queue.empty(); }
                if (waiting) std::this_thread::sleep_for(std::chrono::microseconds(200));   // short of fds held by another device's service
                continue;
            }
            if (!holding_ring) { d->mu.lock(); holding_ring = true; }
            int rc = round_once();
            if (rc != MXD_OK) {              // a CUDA call failed: nothing in flight can be trusted
                const std::string msg = last_error();
                cudaStreamSynchronize(d->compute); cudaStreamSynchronize(d->copy);
                for (int s = 0; s < kSlots; ++s) {
                    for (auto& f : slotfin[s]) if (f.s->status == MXD_OK) { f.s->status = rc; f.s->error = msg; }
                    retire(s);
                }
                for (Stream* s : std::vector<Stream*>(active)) fail_stream(s, rc, msg);
                active.clear();
            }
        }
        drain();
        if (holding_ring) { d->mu.unlock(); holding_ring = false; }
    }
};

void svc_destroy(DevState* d) {
    std::lock_guard<std::mutex> lk(d->svc_mu);
    delete d->svc;
    d->svc = nullptr;
}

static LaneService* service_of(Core* c, DevState* d) {
    std::lock_guard<std::mutex> lk(d->svc_mu);
    if (!d->svc) d->svc = new LaneService(c, d);
    return d->svc;
}

int svc_run(mxd_ctx* h, std::vector<StreamReq>& reqs) {
    Core* c = h->core;
    Group g;
    std::vector<std::unique_ptr<Stream>> streams;
    // sizes first (a stat per file) so the streams can be spread over the devices by bytes, largest first
    for (auto& rq : reqs) {
        rq.status = MXD_OK; rq.error.clear();
        if (rq.path) {
            struct stat st;
            if (stat(rq.path, &st) != 0) { rq.status = MXD_ERR_IO; rq.error = std::string("open ") + rq.path + ": " + strerror(errno); continue; }
            rq.size = (uint64_t)st.st_size;
        } else if (rq.size && !rq.mem) { rq.status = MXD_ERR_INVALID; rq.error = "null buffer with non-zero length"; continue; }
        auto s = std::unique_ptr<Stream>(new Stream(h));
        s->req = &rq; s->group = &g; s->size = rq.size;
        streams.push_back(std::move(s));
    }
    if (!streams.empty()) {
        std::vector<Stream*> order;
        for (auto& s : streams) order.push_back(s.get());
        if (c->devs.size() > 1) std::stable_sort(order.begin(), order.end(), [](Stream* a, Stream* b) { return a->size > b->size; });
        g.remaining = order.size();
        std::vector<LaneService*> svcs;
        for (DevState* d : c->devs) svcs.push_back(service_of(c, d));
        std::vector<uint64_t> planned(svcs.size(), 0);
        std::vector<LaneService*> target;
        for (Stream* s : order) {      // least loaded device (bytes already queued there + what this call has planned)
            size_t best = 0; uint64_t best_load = ~0ull;
            for (size_t i = 0; i < svcs.size(); ++i) {
                const uint64_t load = svcs[i]->pending_bytes.load() + planned[i];
                if (load < best_load) { best_load = load; best = i; }
            }
            planned[best] += s->size + 1;
            target.push_back(svcs[best]);
        }
        for (size_t i = 0; i < order.size(); ++i) target[i]->submit(order[i]);
        std::unique_lock<std::mutex> lk(g.mu);
        g.cv.wait(lk, [&] { return g.remaining == 0; });
    }
    for (auto& rq : reqs) if (rq.status != MXD_OK) return fail(rq.status, rq.error);
    return MXD_OK;
}

}  // namespace mxdi
