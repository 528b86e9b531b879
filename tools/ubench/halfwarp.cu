// Does a warp with fewer active lanes issue integer ALU instructions faster on sm_100a?
// The ALU pipe is 16 lanes wide per SM sub-partition (a 32-lane warp instruction occupies it for 2 clk,
// profiles/r01_pipes_ubench.txt).  If a warp whose upper half is inactive took 1 clk, few-chain SHA-256 launches
// (k_sha256_chains_coop) should run 16 chains per warp.  This measures clk per SHF / LOP3 / IMAD for ONE warp per
// sub-partition with 32, 16, 8 and 1 active lanes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/halfwarp tools/ubench/halfwarp.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CH = 8, INNER = 64;

template <int OP>
__global__ void k(uint32_t* out, int iters, int active, uint32_t one, long long* cyc) {
    if ((threadIdx.x & 31) >= active) return;
    uint32_t x[CH], y = blockIdx.x * 7 + 1;
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 2654435761u + c;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < INNER; ++j) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (OP == 0) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[c]));
                else if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(one));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(one), "r"(y));
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

int main() {
    uint32_t* out; long long* cyc;
    const int warps_per_cta = 4, ctas = 148;      // one warp per sub-partition
    cudaMalloc(&out, ctas * warps_per_cta * 32 * 4);
    cudaMalloc(&cyc, ctas * warps_per_cta * 8);
    const char* names[3] = {"SHF", "LOP3", "IMAD"};
    for (int op = 0; op < 3; ++op)
        for (int wpc : {4, 8})                      // 1 or 2 warps per sub-partition
            for (int active : {32, 16, 8, 1}) {
                const int iters = 200;
                for (int rep = 0; rep < 2; ++rep) {
                    if (op == 0) k<0><<<ctas, wpc * 32>>>(out, iters, active, 1, cyc);
                    if (op == 1) k<1><<<ctas, wpc * 32>>>(out, iters, active, 1, cyc);
                    if (op == 2) k<2><<<ctas, wpc * 32>>>(out, iters, active, 1, cyc);
                    cudaDeviceSynchronize();
                }
                long long h[8];
                cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
                const double per = (double)h[0] / ((double)iters * INNER * CH);
                printf("%-5s warps/subpartition=%d active lanes=%2d  %.3f clk per warp-instruction (8 independent chains)\n",
                       names[op], wpc / 4, active, per);
            }
    return 0;
}
