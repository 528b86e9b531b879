// Internal (non-ABI) interface between the CUDA kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mxd {

// One launch hashes `nmsg` independent messages, one SHA-256 chain per lane.
struct MsgJob {
    // --- where message m lives -------------------------------------------------------------
    // uniform segments of one buffer (tree levels, chunk ranges): ptr = base + m*seg,
    // len = min(seg, nbytes - m*seg); nbytes == 0 means one empty message.
    const uint8_t* base;
    uint64_t nbytes;
    uint64_t seg;
    // arbitrary spans (batches of blobs): spans[m] = {ptr, len}, device memory. Used when base == nullptr.
    const void* spans;
    uint64_t nmsg;
    // --- results ---------------------------------------------------------------------------
    uint8_t* out;            // nmsg * 32 bytes, written when finalize != 0
    // --- chained mode (streaming a message through several launches) -------------------------
    uint32_t* state;         // nullable; 8 words per message, read at start, written when !finalize
    const uint64_t* prefix;  // nullable; bytes already absorbed per message (multiple of 64)
    uint64_t prefix_all;     // used when prefix == nullptr
    int finalize;            // 1: pad and emit digest; 0: len must be a multiple of 64
    const uint8_t* ctl;      // nullable; per message: 0 absorb only, 1 absorb + finalize, 2 skip (overrides finalize)
    uint32_t one;            // must be 1: opaque multiplier that steers additions to the FMA pipe
};

struct DevSpan { const void* ptr; uint64_t len; };

cudaError_t launch_sha256(const MsgJob& job, cudaStream_t stream);
// ok[i] = (memcmp(got + 32 i, want + 32 i, 32) == 0)
cudaError_t launch_compare(const uint8_t* got, const uint8_t* want, uint64_t n, uint8_t* ok, cudaStream_t stream);
// bytes [offset, offset+n) of the splitmix64 counter stream (offset and n multiples of 8, dst 8-byte aligned)
cudaError_t launch_gen_fill(void* dst, uint64_t offset, uint64_t n, uint64_t seed, cudaStream_t stream);
// root = SHA256("modelx.tree.v1\0\0" || LE64(size) || LE64(leaf) || LE32(fanout) || LE32(0) || top[32])
cudaError_t launch_tree_root(uint64_t size, uint64_t leaf, uint32_t fanout, const uint8_t* top, uint8_t* root,
                             cudaStream_t stream);
int sha256_kernel_regs();

}  // namespace mxd
