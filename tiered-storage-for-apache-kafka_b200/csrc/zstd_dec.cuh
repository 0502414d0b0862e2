// zstd_dec.cuh — placeholder until the decompressor lands (next commit).
#pragma once
#include "ts_common.cuh"
#include "rt.h"
#include "launch_prof.h"
namespace ts {
struct ZstdDecScratch { void* p = nullptr; };
inline const char* zstd_dec_scratch_alloc(ZstdDecScratch&, uint32_t, uint32_t) { return nullptr; }
inline void zstd_dec_scratch_free(ZstdDecScratch&) {}
inline int zstd_decompress_batch(ZstdDecScratch&, rt::stream_t, const uint8_t*, const uint64_t*, const uint32_t*, uint32_t, uint32_t,
                                 uint8_t*, uint64_t*, uint32_t*, uint32_t*, bool, LaunchProf&) { return -2; }
}
