// pack.cuh — transformed chunks from their 16-byte aligned device slots into the packed object layout (chunks back to
// back), so that a batch leaves the device in ONE copy instead of one per chunk.
// Replaces the byte-exact concatenation TransformFinisher's SequenceInputStream performs on the host
// (core/M/transform/TransformFinisher.java:134-144).  Pure data movement: 128-bit stores, source words by funnel shift.
#pragma once
#include "ts_common.cuh"
#include "index_scan.cuh"

namespace ts {

constexpr uint32_t PACK_PIECE = 64 * 1024;       // bytes of one chunk a CTA moves
constexpr int PACK_T = 256;

struct PackArgs {
    const uint8_t* base; uint64_t stride; uint32_t head;     // chunk i starts at base + i * stride + head
    const uint32_t* len; uint32_t n;
    uint8_t* out;                                            // packed: chunk i at sum(len[0..i))
};

__global__ void __launch_bounds__(PACK_T) pack_chunks_kernel(const __grid_constant__ PackArgs A) {
    __shared__ uint64_t wsum[PACK_T / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t chunk = blockIdx.y, piece = blockIdx.x;
    const uint32_t n = A.len[chunk];
    if ((uint64_t)piece * PACK_PIECE >= n) return;           // whole CTA
    // where the chunk starts in the packed layout: sum of the sizes before it
    uint64_t part = 0;
    for (uint32_t i = tid; i < chunk; i += PACK_T) part += A.len[i];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(TS_FULL, part, o);
    if (lane == 0) wsum[w] = part;
    __syncthreads();
    uint64_t at = 0;
    for (int k = 0; k < PACK_T / 32; k++) at += wsum[k];
    const uint32_t lo = piece * PACK_PIECE, hi = min(n, lo + PACK_PIECE);
    const uint8_t* src = A.base + (uint64_t)chunk * A.stride + A.head + lo;
    uint8_t* dst = A.out + at + lo;
    const uint32_t cnt = hi - lo;
    const uint32_t headb = min(cnt, (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15));
    if (tid < headb) dst[tid] = src[tid];
    const uint32_t vecs = (cnt - headb) >> 4;
    uint4* d4 = (uint4*)(dst + headb);
    const uint8_t* s = src + headb;
    for (uint32_t i = tid; i < vecs; i += PACK_T) {
        const uint8_t* q = s + 16 * (size_t)i;
        d4[i] = make_uint4(ld_u32_unaligned(q), ld_u32_unaligned(q + 4), ld_u32_unaligned(q + 8), ld_u32_unaligned(q + 12));
    }
    const uint32_t done = headb + 16 * vecs;
    if (tid < cnt - done) dst[done + tid] = src[done + tid];
}

}  // namespace ts
