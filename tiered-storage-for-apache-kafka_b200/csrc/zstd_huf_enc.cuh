// zstd_huf_enc.cuh — literals section of a compressed block (RFC 8878 §3.1.1.3.1).
#pragma once
#include "ts_common.cuh"
#include "zstd_format.h"

namespace ts {

// Raw_Literals_Block with the 3-byte header (Size_Format 11: 20-bit Regenerated_Size).
__device__ __forceinline__ uint32_t ze_raw_literals(const uint8_t* __restrict__ lits, uint32_t n, uint8_t* body, uint32_t lane) {
    if (lane == 0) {
        body[0] = (uint8_t)((3u << 2) | ((n & 0xf) << 4));
        body[1] = (uint8_t)(n >> 4);
        body[2] = (uint8_t)(n >> 12);
    }
    for (uint32_t i = lane; i < n; i += 32) body[3 + i] = lits[i];
    return 3 + n;
}

// Returns the bytes written at `body`.  `work` is ZB bytes of shared memory, `aux` the (free) hash table area.
__device__ __forceinline__ uint32_t ze_encode_literals(const uint8_t* __restrict__ lits, uint32_t n, uint8_t* body,
                                                       uint32_t* work, uint16_t* aux, uint32_t lane) {
    (void)work; (void)aux;
    return ze_raw_literals(lits, n, body, lane);
}

}  // namespace ts
