// CPU-only probe: how fast can T threads move a page-cache resident (tmpfs) file into a staging buffer?
//   pread     what the ring fillers do today (kernel copy_to_user)
//   memcpy    glibc memcpy out of an mmap of the file
//   ntcopy    AVX2 loads + non-temporal stores out of the mmap (no read-for-ownership of the destination lines)
// The staging copy is what bounds file -> GPU throughput once the container's CPU quota is the limit.
// Build: gcc -O2 -mavx2 -pthread -o build/copy_probe tools/ubench/copy_probe.c ; run: build/copy_probe <GiB> <threads>
#define _GNU_SOURCE
#include <fcntl.h>
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

static const size_t PIECE = 4u << 20, SLOT = 64u << 20;
static int g_fd, g_mode, g_threads;
static size_t g_size;
static const uint8_t* g_map;
static uint8_t* g_slots;          // one 64 MiB slot per thread group, reused like the ring
static _Atomic size_t g_next;

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }

static void ntcopy(uint8_t* dst, const uint8_t* src, size_t n) {
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a); _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c); _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
    _mm_sfence();
}

static void* work(void* arg) {
    (void)arg;
    for (;;) {
        size_t p = g_next++;
        size_t off = p * PIECE;
        if (off >= g_size) break;
        size_t n = g_size - off < PIECE ? g_size - off : PIECE;
        uint8_t* dst = g_slots + (off % (4 * SLOT));          // a 4-slot ring, like the library
        if (g_mode == 0) { size_t got = 0; while (got < n) { ssize_t r = pread(g_fd, dst + got, n - got, off + got); if (r <= 0) break; got += r; } }
        else if (g_mode == 1) memcpy(dst, g_map + off, n);
        else ntcopy(dst, g_map + off, n);
    }
    return NULL;
}

int main(int argc, char** argv) {
    size_t gib = argc > 1 ? atol(argv[1]) : 8;
    g_threads = argc > 2 ? atoi(argv[2]) : 16;
    g_size = gib << 30;
    char path[128]; snprintf(path, sizeof path, "/dev/shm/copy_probe_%d.bin", getpid());
    g_fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0600);
    if (g_fd < 0 || ftruncate(g_fd, g_size)) { perror("file"); return 1; }
    uint8_t* w = mmap(NULL, g_size, PROT_READ | PROT_WRITE, MAP_SHARED, g_fd, 0);
    for (size_t i = 0; i < g_size; i += 4096) w[i] = (uint8_t)i;      // materialise the pages
    munmap(w, g_size);
    g_slots = aligned_alloc(1 << 21, 4 * SLOT);
    memset(g_slots, 1, 4 * SLOT);
    const char* names[3] = {"pread", "memcpy(mmap)", "ntcopy(mmap)"};
    for (int rep = 0; rep < 2; ++rep)
        for (g_mode = 0; g_mode < 3; ++g_mode) {
            g_map = mmap(NULL, g_size, PROT_READ, MAP_SHARED, g_fd, 0);   // a FRESH mapping each time: page faults included
            madvise((void*)g_map, g_size, MADV_SEQUENTIAL);
            g_next = 0;
            pthread_t th[256];
            double t0 = now();
            for (int i = 0; i < g_threads; ++i) pthread_create(&th[i], NULL, work, NULL);
            for (int i = 0; i < g_threads; ++i) pthread_join(th[i], NULL);
            double dt = now() - t0;
            printf("threads=%d %-14s %6.1f GB/s (%.2f s, fresh mapping, rep %d)\n", g_threads, names[g_mode], g_size / dt / 1e9, dt, rep);
            munmap((void*)g_map, g_size);
        }
    close(g_fd); unlink(path);
    return 0;
}
