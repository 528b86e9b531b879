// Integer-pipe issue-rate microbenchmark for sm_100a.
//
// SHA-256 is 32-bit rotate/xor/add only, so its ceiling on a B200 is set by how many
// SHF / LOP3 / IADD3 (ALU pipe) and IMAD (FMA pipe) warp-instructions an SM sub-partition
// can issue per clock, alone and mixed.  This tool measures exactly that, so DESIGN.md can
// quote a measured INT-issue roofline instead of an estimate.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int CHAINS = 8;      // independent dependency chains per thread (covers the 4-cycle ALU latency)
constexpr int INNER  = 32;     // unrolled ops per chain per loop trip

enum Mix { SHF = 0, LOP3, IADD3, IMADADD, IMADWIDE, IMADSHL, SHF_IMAD, LOP3_IMAD, SHF_LOP3, SHF_IMADWIDE,
           SHA_MIX, PRMT, IMADWIDE_IMM, SHF_IMAD_IMAD, IMADHI, IMADHI_LOP3, IMADHI_SHF_LOP3, ROT_FMA_LOP3, NMIX };
static const char* mix_name[NMIX] = {
  "SHF.R.W (rotate)", "LOP3 (xor3)", "IADD3", "IMAD (x*1+y, fma pipe)", "1 IMAD.WIDE.U32 (x*reg) : 1 LOP3",
  "IMAD.SHL (x*2^k imm)", "1 SHF : 1 IMAD", "1 LOP3 : 1 IMAD", "1 SHF : 1 LOP3", "1 SHF : 1 IMAD.WIDE : 1 LOP3",
  "sha-like 6 SHF : 3 LOP3 : 2 IADD3 : 1 IMAD", "PRMT", "1 IMAD.WIDE.U32 (x*imm) : 1 LOP3", "1 SHF : 2 IMAD",
  "IMAD.HI.U32 (x*reg hi)", "1 IMAD.HI.U32 : 1 LOP3", "1 IMAD.HI : 1 SHF : 1 LOP3", "rot via IMAD.HI+IMAD : 1 LOP3 (2 fma : 1 alu)" };
// instructions issued per "op slot" of the inner loop, for rate accounting
static const int mix_instr[NMIX] = {1, 1, 1, 1, 2, 1, 2, 2, 2, 3, 12, 1, 2, 3, 1, 2, 3, 3};

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
template <int MIX>
__global__ void __launch_bounds__(256) k_pipe(uint32_t* out, int iters, uint32_t one, uint32_t mulc, long long* cycles) {
  unsigned long long g0 = gtime();
  uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x * 2654435761u + c; y[c] = blockIdx.x + c * 40503u; }
  asm volatile("mov.u32 %0, %0;" : "+r"(one));
  asm volatile("mov.u32 %0, %0;" : "+r"(mulc));
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < INNER; ++j) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
        if (MIX == SHF) {
          asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[c]));
        } else if (MIX == LOP3) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(one));
        } else if (MIX == IADD3) {
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x[c]) : "r"(y[c]), "r"(one));
        } else if (MIX == IMADADD) {
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(one), "r"(y[c]));
        } else if (MIX == IMADWIDE) {
          uint64_t w;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[c]), "r"(mulc));
          asm volatile("xor.b32 %0, %1, %2;" : "=r"(x[c]) : "r"((uint32_t)w), "r"((uint32_t)(w >> 32)));
        } else if (MIX == IMADWIDE_IMM) {
          uint64_t w;
          asm volatile("mul.wide.u32 %0, %1, 67108865;" : "=l"(w) : "r"(x[c]));
          asm volatile("xor.b32 %0, %1, %2;" : "=r"(x[c]) : "r"((uint32_t)w), "r"((uint32_t)(w >> 32)));
        } else if (MIX == IMADSHL) {
          asm volatile("mad.lo.u32 %0, %0, 128, %1;" : "+r"(x[c]) : "r"(y[c]));
        } else if (MIX == PRMT) {
          asm volatile("prmt.b32 %0, %0, %1, 0x0123;" : "+r"(x[c]) : "r"(y[c]));
        } else if (MIX == SHF_IMAD) {
          asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[c]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(one), "r"(x[c]));
        } else if (MIX == SHF_IMAD_IMAD) {
          asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[c]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(one), "r"(x[c]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(mulc), "r"(x[c]));
        } else if (MIX == LOP3_IMAD) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(one));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(one), "r"(x[c]));
        } else if (MIX == SHF_LOP3) {
          asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[c]));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[c]) : "r"(x[c]), "r"(one));
        } else if (MIX == SHF_IMADWIDE) {
          uint64_t w;
          asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[c]));
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(y[c]), "r"(mulc));
          y[c] = (uint32_t)w ^ (uint32_t)(w >> 32);   // one extra LOP3 (counted as part of the 2; slight under-count)
        } else if (MIX == IMADHI) {
          asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(mulc));
          asm volatile("" : "+r"(x[c]));
        } else if (MIX == IMADHI_LOP3) {
          asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(y[c]) : "r"(x[c]), "r"(mulc));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(one));
        } else if (MIX == IMADHI_SHF_LOP3) {
          uint32_t t;
          asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(x[c]), "r"(mulc));
          asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(y[c]));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(t));
        } else if (MIX == ROT_FMA_LOP3) {
          uint32_t t;
          asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(x[c]), "r"(mulc));          // x >> 6
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(x[c]), "r"(mulc));      // + x << 26  => rotr(x, 6)
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(t));
        } else if (MIX == SHA_MIX) {
          uint32_t a, b, d;
          asm volatile("shf.r.wrap.b32 %0, %1, %1, 6;"  : "=r"(a) : "r"(x[c]));
          asm volatile("shf.r.wrap.b32 %0, %1, %1, 11;" : "=r"(b) : "r"(x[c]));
          asm volatile("shf.r.wrap.b32 %0, %1, %1, 25;" : "=r"(d) : "r"(x[c]));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(d));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(b) : "r"(x[c]), "r"(y[c]), "r"(one));
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(y[c]) : "r"(a), "r"(b));
          asm volatile("shf.r.wrap.b32 %0, %1, %1, 2;"  : "=r"(a) : "r"(y[c]));
          asm volatile("shf.r.wrap.b32 %0, %1, %1, 13;" : "=r"(b) : "r"(y[c]));
          asm volatile("shf.r.wrap.b32 %0, %1, %1, 22;" : "=r"(d) : "r"(y[c]));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(d));
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x[c]) : "r"(a), "r"(one));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(one), "r"(y[c]));
        }
      }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc ^= x[c] ^ y[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) { cycles[2 * blockIdx.x] = t1 - t0; cycles[2 * blockIdx.x + 1] = (long long)(gtime() - g0); }
}

template <int MIX>
int run(int nsm, int warps_per_sm, uint32_t* out, long long* cyc_d, int sm_khz) {
  const int threads = 256;
  const int blocks_per_sm = warps_per_sm * 32 / threads;
  const int grid = nsm * blocks_per_sm;
  const int iters = 4000;
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pipe<MIX>, threads, 0));
  if (occ < blocks_per_sm) { printf("%-46s warps/SM=%2d  SKIP: only %d blocks/SM resident\n", mix_name[MIX], warps_per_sm, occ); return 0; }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_pipe<MIX><<<grid, threads>>>(out, 10, 1u, 67108864u, cyc_d);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_pipe<MIX><<<grid, threads>>>(out, iters, 1u, 67108864u, cyc_d);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  static long long cyc_h[8192];
  CK(cudaMemcpy(cyc_h, cyc_d, sizeof(long long) * grid * 2, cudaMemcpyDeviceToHost));
  double avg = 0, avgns = 0; for (int i = 0; i < grid; ++i) { avg += (double)cyc_h[2 * i]; avgns += (double)cyc_h[2 * i + 1]; }
  avg /= grid; avgns /= grid;
  const double slots = (double)iters * INNER * CHAINS;                  // op slots per thread
  const double winstr_per_sm = slots * mix_instr[MIX] * warps_per_sm;   // warp-instructions per SM
  const double per_clk_sm = winstr_per_sm / avg;                        // warp-instr / clk / SM (by clock64)
  printf("%-46s warps/SM=%2d  %6.3f warp-instr/clk/SM (%5.1f lane-ops/clk/SM)  %7.3f ms  %6.3f G warp-instr/s/SM  SM clock %.0f MHz (clock64/globaltimer)\n",
         mix_name[MIX], warps_per_sm, per_clk_sm, per_clk_sm * 32, ms, winstr_per_sm / (ms * 1e-3) / 1e9, avg / avgns * 1e3);
  (void)sm_khz;
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s  SMs=%d  cc=%d.%d  clock=%d kHz  L2=%d MB  smem/SM=%zu\n", p.name, p.multiProcessorCount,
         p.major, p.minor, p.clockRate, p.l2CacheSize >> 20, p.sharedMemPerMultiprocessor);
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(uint32_t) * 148 * 2048 * 2)); CK(cudaMalloc(&cyc, sizeof(long long) * 8192));
  k_pipe<SHA_MIX><<<148 * 2, 256>>>(out, 20000, 1u, 3u, cyc); CK(cudaDeviceSynchronize());  // clock warm-up
  const int nsm = p.multiProcessorCount;
  for (int w : {16, 32}) {
    run<SHF>(nsm, w, out, cyc, p.clockRate);
    run<LOP3>(nsm, w, out, cyc, p.clockRate);
    run<IADD3>(nsm, w, out, cyc, p.clockRate);
    run<PRMT>(nsm, w, out, cyc, p.clockRate);
    run<IMADADD>(nsm, w, out, cyc, p.clockRate);
    run<IMADSHL>(nsm, w, out, cyc, p.clockRate);
    run<IMADWIDE>(nsm, w, out, cyc, p.clockRate);
    run<IMADWIDE_IMM>(nsm, w, out, cyc, p.clockRate);
    run<SHF_IMAD>(nsm, w, out, cyc, p.clockRate);
    run<SHF_IMAD_IMAD>(nsm, w, out, cyc, p.clockRate);
    run<LOP3_IMAD>(nsm, w, out, cyc, p.clockRate);
    run<SHF_LOP3>(nsm, w, out, cyc, p.clockRate);
    run<SHF_IMADWIDE>(nsm, w, out, cyc, p.clockRate);
    run<SHA_MIX>(nsm, w, out, cyc, p.clockRate);
    run<IMADHI>(nsm, w, out, cyc, p.clockRate);
    run<IMADHI_LOP3>(nsm, w, out, cyc, p.clockRate);
    run<IMADHI_SHF_LOP3>(nsm, w, out, cyc, p.clockRate);
    run<ROT_FMA_LOP3>(nsm, w, out, cyc, p.clockRate);
    printf("\n");
  }
  return 0;
}
