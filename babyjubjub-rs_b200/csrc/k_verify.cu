// verify_batch / verify_compressed_batch kernels (reference: verify, src/lib.rs:395-412; decompress_signature
// + decompress_point, src/lib.rs:260-268, 192-224).  The lane bodies are in lanes.cuh.
//
// One verification is a short pipeline of kernels, so that no kernel carries another phase's registers
// or instruction footprint:
//   (bjj_cuda.cu: k_decompress_prepare / k_batch_inverse / k_decompress_finish for compressed input) ->
//   k_verify_hash -> k_verify_split (EdDSA) -> k_verify_ec || k_verify_exact
#include "kernels.h"

using namespace bjj;

// resident CTAs per SM the compiler must make room for (register cap = 65536 / (128 * MINB)); the carry
// chains are dependent instruction streams, so the fma pipe needs >= 3-4 warps per SMSP to stay busy
#ifndef BJJ_VERIFY_HASH_MINB
#define BJJ_VERIFY_HASH_MINB 2
#endif
#define BJJ_VERIFY_EXACT_BLOCK 64

#if BJJ_VERIFY_HASH_MINB > 0
__global__ void __launch_bounds__(BJJ_BLOCK, BJJ_VERIFY_HASH_MINB) k_verify_hash(
#else
__global__ void __launch_bounds__(BJJ_BLOCK) k_verify_hash(
#endif
    size_t n, const uint8_t* r8x, const uint8_t* r8y, const uint8_t* ax, const uint8_t* ay, const uint8_t* msg,
    const uint8_t* s_base, size_t s_stride, size_t s_off, const uint8_t* skip, uint8_t* hm, size_t plane, uint8_t* ok,
    int gate, ExactQueue qa, ExactQueue qr, uint32_t* gflags, int mode, int split, uint8_t* msg_status,
    unsigned long long* work) {
    BJJ_FLAGS_BEGIN
    BJJ_CLAIM_LOOP(n, work)
    if (i < n)
    lane_verify_hash(r8x, r8y, ax, ay, msg, s_base, s_stride, s_off, skip, hm, plane, ok, i, gate != 0, qa, qr, flags, mode,
                     split != 0, msg_status);
    BJJ_FLAGS_END(gflags)
}

// half-size scalars (EdDSA lanes): pure integer-ALU work, small code, many resident warps
__global__ void __launch_bounds__(BJJ_BLOCK, 4) k_verify_split(size_t n, const uint8_t* s_base, size_t s_stride, size_t s_off,
                                                              uint8_t* hm, size_t plane, const uint8_t* ok) {
    BJJ_LANE_LOOP(n) lane_verify_split(s_base, s_stride, s_off, hm, plane, ok, i);
}

// Register budget left to the compiler (239 registers, 2 CTAs per SM): capping it at 168 for a third CTA
// costs spills inside the Straus loop (measured 15.4 ms vs 11.7 ms per 2^18 lanes).  The register file is
// partitioned per SMSP (16,384 registers each), so the exact-lane kernel cannot co-reside with this one
// whatever the cap; it runs on a side stream and fills the tail instead.
__global__ void __launch_bounds__(BJJ_BLOCK, 2) k_verify_ec(size_t n, const uint8_t* r8x, const uint8_t* r8y,
                                                         const uint8_t* ax, const uint8_t* ay, const uint8_t* hm,
                                                         size_t plane, uint8_t* ok, U128* table, const CombEntry* comb,
                                                         int mode, unsigned long long* work) {
    // two per-thread radix-16 tables (multiples of 8A and of R8), back to back
    const LaneTable tbl_a = thread_table(table);
    const LaneTable tbl_r = thread_table(table + (size_t)BJJ_TABLE_U128_PER_LANE * gridDim.x * blockDim.x);
    BJJ_CLAIM_LOOP(n, work)
    if (i < n) lane_verify_ec(r8x, r8y, ax, ay, hm, plane, ok, i, tbl_a, tbl_r, comb, mode);
}

// exact lanes: off-curve inputs replay the reference sequence (rare; fed by the queues of k_verify_hash).
// One launch serves both queues: the first half of the grid takes the "A off the curve" queue, the second
// half the "only R8 off the curve" queue, so no warp ever mixes the two ladders.
__global__ void __launch_bounds__(BJJ_VERIFY_EXACT_BLOCK) k_verify_exact(const uint8_t* r8x, const uint8_t* r8y, const uint8_t* s,
                                                                         const uint8_t* ax, const uint8_t* ay, const uint8_t* hm,
                                                                         uint8_t* ok, ExactQueue qa, ExactQueue qr,
                                                                         const CombEntry* comb, int mode) {
    const uint32_t half = gridDim.x / 2;
    if (blockIdx.x < half) {
        for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, cnt = *qa.count; j < cnt; j += half * blockDim.x)
            lane_verify_exact<true>(r8x, r8y, s, ax, ay, hm, ok, qa.list[j], comb, mode);
    } else {
        for (uint32_t j = (blockIdx.x - half) * blockDim.x + threadIdx.x, cnt = *qr.count; j < cnt; j += half * blockDim.x)
            lane_verify_exact<false>(r8x, r8y, s, ax, ay, hm, ok, qr.list[j], comb, mode);
    }
}

namespace bjjk {

static int occ(const void* k, int block) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm;
}
int verify_hash_blocks_per_sm() { return occ((const void*)k_verify_hash, BJJ_BLOCK); }
int verify_ec_blocks_per_sm() { return occ((const void*)k_verify_ec, BJJ_BLOCK); }
int verify_split_blocks_per_sm() { return occ((const void*)k_verify_split, BJJ_BLOCK); }

void verify_hash(int grid, cudaStream_t st, size_t n, const uint8_t* r8x, const uint8_t* r8y, const uint8_t* ax,
                 const uint8_t* ay, const uint8_t* msg, const uint8_t* s_base, size_t s_stride, size_t s_off,
                 const uint8_t* skip, uint8_t* hm, size_t plane, uint8_t* ok, bool gate, ExactQueue qa, ExactQueue qr,
                 uint32_t* gflags, int mode, bool split, uint8_t* msg_status, unsigned long long* work) {
    k_verify_hash<<<grid, BJJ_BLOCK, 0, st>>>(n, r8x, r8y, ax, ay, msg, s_base, s_stride, s_off, skip, hm, plane, ok, gate ? 1 : 0,
                                              qa, qr, gflags, mode, split ? 1 : 0, msg_status, work);
}
void verify_split(int grid, cudaStream_t st, size_t n, const uint8_t* s_base, size_t s_stride, size_t s_off, uint8_t* hm,
                  size_t plane, const uint8_t* ok) {
    k_verify_split<<<grid, BJJ_BLOCK, 0, st>>>(n, s_base, s_stride, s_off, hm, plane, ok);
}
void verify_ec(int grid, cudaStream_t st, size_t n, const uint8_t* r8x, const uint8_t* r8y, const uint8_t* ax,
               const uint8_t* ay, const uint8_t* hm, size_t plane, uint8_t* ok, U128* table, const CombEntry* comb,
               int mode, unsigned long long* work) {
    k_verify_ec<<<grid, BJJ_BLOCK, 0, st>>>(n, r8x, r8y, ax, ay, hm, plane, ok, table, comb, mode, work);
}
void verify_exact(int grid, cudaStream_t st, const uint8_t* r8x, const uint8_t* r8y, const uint8_t* s, const uint8_t* ax,
                  const uint8_t* ay, const uint8_t* hm, uint8_t* ok, ExactQueue qa, ExactQueue qr, const CombEntry* comb,
                  int mode) {
    k_verify_exact<<<grid & ~1, BJJ_VERIFY_EXACT_BLOCK, 0, st>>>(r8x, r8y, s, ax, ay, hm, ok, qa, qr, comb, mode);
}

}  // namespace bjjk
