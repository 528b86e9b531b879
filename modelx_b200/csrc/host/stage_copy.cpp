// The staging copy: page cache / pageable memory -> a pinned ring slot.
//
// Once the GPU side is fed at PCIe rate this copy is what bounds file -> digest throughput: every
// byte the CPU copies crosses DRAM as a source read, a read-for-ownership of the destination line, its write-back, and the
// copy engine's read.  Measured on the bench box (tools/ubench/copy_probe.c, profiles/r02_ring_sweep_and_copy_probe.txt),
// 16 threads, tmpfs file: pread 36.6 GB/s, memcpy out of an mmap 40 GB/s, AVX2 loads + NON-TEMPORAL stores out of an mmap
// 56.8 GB/s (no RFO, no cache pollution; 4 threads: 18 / 19 / 27-31 GB/s).  In a bare process, that is.  Inside the library
// the picture flips (profiles/r02_stage_copy_ab.txt, 24 GB file): pread 47.5 GB/s, mapped + streaming stores 32 GB/s,
// mapped + memcpy 31 GB/s with 16 filler threads.  Why is open: pre-filling the page tables of every piece with
// madvise(MADV_POPULATE_READ) changes nothing (33.0 vs 32.4 GB/s, profiles/r02_stage_populate_ab.txt), so it is not the
// per-page fault cost.  Only when CPUs are scarce does the cheaper copy win (2 CPUs: 11.4 vs 9.7 GB/s, 4 CPUs: 19.8 vs 18.8).  So the DEFAULT stays pread + memcpy, and
// MXD_STAGE_MMAP=1 selects the mapped streaming-store copy for CPU-starved hosts.
// A mapping can fault if the file is truncated while it is being hashed (pread would return a short read): the mapped copy
// runs under a SIGBUS guard that turns the fault into "file shrank while hashing" instead of killing the host process;
// the handler chains to whatever was installed before for faults that are not ours, and is installed with SA_ONSTACK so
// that it is safe under the Go runtime.
#include <atomic>
#include <csetjmp>
#include <csignal>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
#include <mutex>
#include <sys/mman.h>
#include <unistd.h>

namespace mxdi {

namespace {

__attribute__((target("avx2"))) void copy_nt_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
    size_t i = 0;
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;     // streaming stores want 32-byte aligned lines
    if (head && head <= n) { memcpy(dst, src, head); i = head; }
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
    _mm_sfence();            // the streamed lines are globally visible before the slot is handed to the copy engine
}

const bool g_mmap = [] { const char* e = getenv("MXD_STAGE_MMAP"); return e && e[0] == '1'; }();
const bool g_avx2 = g_mmap && __builtin_cpu_supports("avx2") && getenv("MXD_STAGE_NO_NT") == nullptr;

thread_local sigjmp_buf* t_guard = nullptr;
struct sigaction g_prev_bus;
std::once_flag g_bus_once;

void on_sigbus(int sig, siginfo_t* info, void* uctx) {
    if (t_guard) siglongjmp(*t_guard, 1);                       // a fault inside stage_copy_mapped: unwind to it
    if (g_prev_bus.sa_flags & SA_SIGINFO) { if (g_prev_bus.sa_sigaction) { g_prev_bus.sa_sigaction(sig, info, uctx); return; } }
    else if (g_prev_bus.sa_handler != SIG_DFL && g_prev_bus.sa_handler != SIG_IGN) { g_prev_bus.sa_handler(sig); return; }
    signal(SIGBUS, SIG_DFL);                                   // not ours and nobody else's: die the default way
    raise(SIGBUS);
}

void install_guard() {
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = on_sigbus;
    sa.sa_flags = SA_SIGINFO | SA_ONSTACK | SA_NODEFER;
    sigemptyset(&sa.sa_mask);
    sigaction(SIGBUS, &sa, &g_prev_bus);
}

}  // namespace

// memory -> slot
void stage_copy(uint8_t* dst, const uint8_t* src, size_t n) {
    if (g_avx2 && n >= 4096) copy_nt_avx2(dst, src, n); else memcpy(dst, src, n);
}

// file mapping -> slot; -1 when the mapping faulted (the file shrank under us)
int stage_copy_mapped(uint8_t* dst, const uint8_t* src, size_t n) {
    std::call_once(g_bus_once, install_guard);
    sigjmp_buf jb;
    t_guard = &jb;
    int rc = 0;
    if (sigsetjmp(jb, 1) == 0) stage_copy(dst, src, n);
    else rc = -1;
    t_guard = nullptr;
    return rc;
}

bool stage_mmap_enabled() {
    return g_mmap;
}

// Map bytes [base, base + nbytes) of fd read-only; *map = address of byte `base`.  Returns the handle for file_unmap or null.
void* file_map(int fd, uint64_t base, uint64_t nbytes, const uint8_t** map, uint64_t* handle_len) {
    *map = nullptr; *handle_len = 0;
    if (!stage_mmap_enabled() || nbytes == 0) return nullptr;
    const long page = sysconf(_SC_PAGESIZE);
    if (page <= 0) return nullptr;
    const uint64_t lead = base % (uint64_t)page;
    void* p = mmap(nullptr, nbytes + lead, PROT_READ, MAP_SHARED, fd, (off_t)(base - lead));
    if (p == MAP_FAILED) return nullptr;
    madvise(p, nbytes + lead, MADV_SEQUENTIAL);
    *map = static_cast<const uint8_t*>(p) + lead;
    *handle_len = nbytes + lead;
    return p;
}
void file_unmap(void* handle, uint64_t handle_len) { if (handle) munmap(handle, handle_len); }

}  // namespace mxdi
