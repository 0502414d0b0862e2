// zstd_enc.cuh — batched Zstandard compression for sm_100a (kernel K1 of SURVEY.md §2a).
//
// Replaces ZstdCompressCtx.compress as called per chunk from
//   core/M/transform/CompressionChunkEnumeration.java:49-62 (new ctx, default level, setPledgedSrcSize,
//   setContentSize(true)): one independent RFC 8878 frame WITH Frame_Content_Size per chunk.
// Parity bar: libzstd decodes every frame to the original bytes and
// ZSTD_getFrameContentSize(frame) == original size.  Compressed bytes are not expected to equal libzstd's.
//
// B200-first decomposition: a chunk is cut into ZB = 8 KiB zstd blocks that never reference each other
// (no cross-block matches, no repeat offsets, no repeated tables), so a 1 GiB segment is 131,072 independent
// units instead of 256 sequential ones.
//   zstd_enc_blocks_kernel    one WARP per block: block staged in shared memory; 32 positions hashed and
//                             verified per step against a per-warp hash table, greedy left-to-right selection
//                             by ballot/ffs, match extension per lane then warp-wide; literals (Raw or
//                             Huffman, see zstd_huf_enc.cuh) + sequences (predefined FSE tables; the three
//                             state chains run on lanes 0-2, bit packing on all lanes with shuffle prefix sums)
//   zstd_enc_assemble_kernel  one CTA per chunk: frame header (same form libzstd picks for the size), exclusive
//                             scan of block sizes, byte-granular gather of the blocks into the frame
// Everything is integer/byte work on the ALU and shared-memory pipes; there is no dense contraction to put
// on tensor cores.
#pragma once
#include "ts_common.cuh"
#include "rt.h"
#include "launch_prof.h"
#include "zstd_format.h"
#include "index_scan.cuh"

namespace ts {

constexpr uint32_t ZB = 8192;                  // bytes of original data per zstd block
#ifndef ZE_TUNE_HLOG                           // the ZE_TUNE_* macros exist for parameter studies (-DZE_TUNE_...=v); the defaults are the product
#define ZE_TUNE_HLOG 10
#define ZE_TUNE_MIN_MATCH 5
#define ZE_TUNE_LANE_EXT 12
#endif
constexpr int ZE_HLOG = ZE_TUNE_HLOG;          // per-warp hash table: 2^10 x u16 (position + 1)
constexpr uint32_t ZE_HSIZE = 1u << ZE_HLOG;
constexpr int ZE_WPB = 4;                      // warps (= blocks in flight) per CTA
constexpr uint32_t ZE_MAXSEQ = ZB / 8;         // sequences kept per block; beyond that the rest goes out as literals
constexpr uint32_t ZE_MIN_MATCH = ZE_TUNE_MIN_MATCH;   // 5: a 4-byte match costs more bits than four Huffman-coded literals (libzstd level 3 also uses 5)
constexpr uint32_t ZE_LANE_EXT = ZE_TUNE_LANE_EXT;     // 12: bytes a lane extends its own match beyond the first 4
constexpr uint32_t ZE_BUF_PAD = 160;
constexpr uint32_t ZE_SLOT = ZB + 512;         // per-block output slot: 3-byte header + payload (< ZB once accepted; table descriptions are written before that is known)
constexpr uint32_t ZE_SMEM_WARP = ZB + ZE_BUF_PAD + ZE_HSIZE * 2;   // buf, ht
static_assert(ZB < 65535 && ZE_MAXSEQ <= 1024, "16-bit hash slots / 7-bit FSE tables assume small blocks");
// per-tile FSE scratch (codes, state bits) lives in the tail of `buf`: the sequence bit stream staged there is at most
// ZE_MAXSEQ * 58 bits, which ends below this offset
constexpr uint32_t ZE_SEQ_AUX_OFF = ZB - 512;
static_assert(ZE_MAXSEQ * 58 / 8 + 16 <= ZE_SEQ_AUX_OFF, "sequence bit stream would overlap the FSE tile scratch");
static_assert(ZE_SEQ_AUX_OFF + 3 * 32 + 3 * 32 * 2 + 3 * 32 <= ZB + ZE_BUF_PAD, "FSE tile scratch must fit the block buffer");
constexpr uint32_t ZE_SMEM_WARP_AL = (ZE_SMEM_WARP + 15) & ~15u;

__constant__ zf::SeqTables g_seq_tables = zf::make_seq_tables();
__constant__ zf::PredefinedCTables g_pre_ctables = zf::make_predefined_ctables();

struct ZstdEncScratch {
    uint8_t* blk_out = nullptr;      // n_chunks * blocks_per_chunk * ZE_SLOT
    uint32_t* blk_size = nullptr;    // n_chunks * blocks_per_chunk
    uint2* seqs = nullptr;           // n_chunks * blocks_per_chunk * ZE_MAXSEQ
    uint8_t* lits = nullptr;         // n_chunks * blocks_per_chunk * ZB   (literal staging for Huffman)
    uint32_t* blk_meta = nullptr;    // n_chunks * blocks_per_chunk * 2    (nseq, nlit: only the two-launch variant uses it)
    bool split = false;              // TSGPU_ENC_SPLIT=1: parse and entropy stage as two launches
    uint32_t blocks_per_chunk = 0;
    uint32_t max_batch = 0;
};

struct ZstdEncArgs {
    const uint8_t* in_base; const uint64_t* in_off; const uint32_t* in_len;
    uint8_t* blk_out; uint32_t* blk_size; uint2* seqs; uint8_t* lits; uint32_t* blk_meta;
    uint32_t blocks_per_chunk;
    uint8_t* out_base; const uint64_t* out_off; uint32_t* out_len;
};

// ------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ uint32_t ze_hash(uint32_t v) { return (v * 2654435761u) >> (32 - ZE_HLOG); }

// number of equal leading bytes (0..4) of two words given their XOR
__device__ __forceinline__ uint32_t ze_common_bytes(uint32_t x) { return x ? (uint32_t)(__ffs((int)x) - 1) >> 3 : 4u; }

__device__ __forceinline__ uint32_t ze_ll_code(uint32_t ll) {
    return ll < 64 ? g_seq_tables.ll_code[ll] : (uint32_t)zf::highbit32(ll) + 19;
}
__device__ __forceinline__ uint32_t ze_ml_code(uint32_t mlbase) {
    return mlbase < 128 ? g_seq_tables.ml_code[mlbase] : (uint32_t)zf::highbit32(mlbase) + 36;
}

// OR a bit field (val, nb <= 58 bits) into a zeroed little-endian word buffer at bit offset `o` (shared memory)
__device__ __forceinline__ void ze_put_bits(uint32_t* words, uint32_t o, uint64_t val, uint32_t nb) {
    if (nb == 0) return;
    const uint32_t w = o >> 5, sh = o & 31;
    atomicOr(&words[w], (uint32_t)(val << sh));
    if (sh + nb > 32) atomicOr(&words[w + 1], (uint32_t)(val >> (32 - sh)));
    if (sh + nb > 64) atomicOr(&words[w + 2], (uint32_t)(val >> (64 - sh)));
}

#ifdef TSGPU_SIMT
static inline unsigned __match_any_sync(unsigned, unsigned v) {
    simt::Warp& w = simt::g_blk->warps[simt::g_cur->warp];
    w.buf[simt::g_cur->lane] = v;
    simt::warp_barrier();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (w.buf[i] == v) r |= 1u << i;
    simt::warp_barrier();
    return r;
}
#endif

}  // namespace ts

#include "zstd_fse_enc.cuh"
#include "zstd_huf_enc.cuh"
static_assert(sizeof(ts::ZeCTab) <= ts::ZE_HSIZE * 2 && 2048 <= ts::ZE_HSIZE * 2,
              "phase B aliases the FSE encoding tables and the 2 KiB Huffman histogram/code table into the hash-table area");

namespace ts {

// ------------------------------------------------------------------------------------------ block compressor
struct ZeFseShared {      // per-CTA copy of the predefined encoding tables
    zf::PredefinedCTables t;
};


// Encodes the sequences section (everything after the Number_of_Sequences field): the modes byte and table
// descriptions go straight to `hdr_out` (global), the bit stream is staged in the word buffer `bits` (shared).
// Returns the bit-stream bytes; *desc_bytes = 1 (modes byte) + table descriptions.  Warp-uniform, N >= 1.
__device__ __forceinline__ uint32_t ze_encode_sequences(const uint2* __restrict__ seqs, uint32_t N, uint32_t* bits, ZeCTab* ct,
                                                        uint8_t* codes /*[3][32]*/, uint16_t* stv /*[3][32]*/, uint8_t* stn /*[3][32]*/,
                                                        const ZeFseShared* fs, uint8_t* hdr_out, uint32_t* desc_bytes, uint32_t lane) {
    // ---- pass 1: code histograms (then normalised counts), in the word buffer beyond ze_build_kind's own scratch
    uint32_t* cnt = bits + 256;
    for (uint32_t i = lane; i < ZE_NSYM_LL + ZE_NSYM_ML + ZE_NSYM_OF; i += 32) cnt[i] = 0;
    __syncwarp();
    uint2 pre = lane < N ? seqs[lane] : make_uint2(0, 0);
    for (uint32_t t0 = 0; t0 < N; t0 += 32) {
        const uint32_t j = t0 + lane;
        const uint2 s = pre;
        if (j + 32 < N) pre = seqs[j + 32];                  // next tile in flight while this one is counted
        if (j < N) {
            atomicAdd(&cnt[ze_ll_code(s.x & 0xffff)], 1u);
            atomicAdd(&cnt[ZE_NSYM_LL + ze_ml_code(s.x >> 16)], 1u);
            atomicAdd(&cnt[ZE_NSYM_LL + ZE_NSYM_ML + (uint32_t)zf::highbit32(s.y + 3)], 1u);
        }
    }
    __syncwarp();
    // ---- tables (descriptions in stream order LL, OF, ML); `bits` doubles as scratch until it is cleared
    ZeKind kll, kof, kml;
    uint8_t* desc = hdr_out + 1;
    uint32_t dn = ze_build_kind(cnt, ZE_NSYM_LL, N, zf::LL_MAX_LOG, zf::LL_DEFAULT_LOG, fs->t.ll.state, fs->t.ll.sym,
                                ct->st_ll, ct->sy_ll, (uint8_t*)bits, desc, &kll, true, lane);
    dn += ze_build_kind(cnt + ZE_NSYM_LL + ZE_NSYM_ML, ZE_NSYM_OF, N, zf::OF_MAX_LOG, zf::OF_DEFAULT_LOG, fs->t.of.state, fs->t.of.sym,
                        ct->st_of, ct->sy_of, (uint8_t*)bits, desc + dn, &kof, true, lane);
    dn += ze_build_kind(cnt + ZE_NSYM_LL, ZE_NSYM_ML, N, zf::ML_MAX_LOG, zf::ML_DEFAULT_LOG, fs->t.ml.state, fs->t.ml.sym,
                        ct->st_ml, ct->sy_ml, (uint8_t*)bits, desc + dn, &kml, true, lane);
    if (lane == 0) hdr_out[0] = (uint8_t)((kll.mode << 6) | (kof.mode << 4) | (kml.mode << 2));
    *desc_bytes = 1 + dn;
    __syncwarp();
    for (uint32_t i = lane; i < (ZB + ZE_BUF_PAD) / 4; i += 32) bits[i] = 0;
    __syncwarp();

    // ---- pass 2: encode
    uint32_t bitpos = 0;
    uint32_t state = 0;                              // lanes 0..2: OF, ML, LL chains
    const uint16_t* st_tab = lane == 0 ? ct->st_of : lane == 1 ? ct->st_ml : ct->st_ll;
    const zf::FseCSym* sy = lane == 0 ? ct->sy_of : lane == 1 ? ct->sy_ml : ct->sy_ll;
    const bool rle = (lane == 0 ? kof.mode : lane == 1 ? kml.mode : kll.mode) == 1;
    pre = lane < N ? seqs[N - 1 - lane] : make_uint2(0, 0);
    for (uint32_t t0 = 0; t0 < N; t0 += 32) {
        const uint32_t j = t0 + lane;                // stream order: j = 0 is the LAST sequence
        const bool have = j < N;
        uint32_t ll = 0, mlb = 0, offb = 0, llc = 0, mlc = 0, ofc = 0;
        const uint2 s = pre;
        if (j + 32 < N) pre = seqs[N - 1 - (j + 32)];
        if (have) {
            ll = s.x & 0xffff; mlb = s.x >> 16; offb = s.y + 3;       // mlb = matchLength - 3, offb = offset + 3 (no repcodes)
            llc = ze_ll_code(ll); mlc = ze_ml_code(mlb); ofc = (uint32_t)zf::highbit32(offb);
            codes[lane] = (uint8_t)ofc; codes[32 + lane] = (uint8_t)mlc; codes[64 + lane] = (uint8_t)llc;
        }
        __syncwarp();
        if (lane < 3) {
            const uint32_t cnt = min(32u, N - t0);
            if (rle) {
                for (uint32_t i = 0; i < cnt; i++) { stv[lane * 32 + i] = 0; stn[lane * 32 + i] = 0; }
            } else {
                zf::FseCSym c = sy[codes[lane * 32]];
                for (uint32_t i = 0; i < cnt; i++) {
                    const zf::FseCSym cn = sy[codes[lane * 32 + min(i + 1, cnt - 1)]];    // prefetch: independent of the state chain
                    if (t0 + i == 0) {               // FSE_initCState2
                        const uint32_t nb = (uint32_t)(c.delta_nb_bits + (1 << 15)) >> 16;
                        const uint32_t v = (nb << 16) - (uint32_t)c.delta_nb_bits;
                        state = st_tab[(int32_t)(v >> nb) + c.delta_find_state];
                        stv[lane * 32 + i] = 0; stn[lane * 32 + i] = 0;
                    } else {                         // FSE_encodeSymbol
                        const uint32_t nb = (state + (uint32_t)c.delta_nb_bits) >> 16;
                        stv[lane * 32 + i] = (uint16_t)(state & ((1u << nb) - 1));
                        stn[lane * 32 + i] = (uint8_t)nb;
                        state = st_tab[(int32_t)(state >> nb) + c.delta_find_state];
                    }
                    c = cn;
                }
            }
        }
        __syncwarp();
        uint64_t field = 0; uint32_t nb = 0;
        if (have) {
            const uint32_t llb = g_seq_tables.ll_bits[llc], mlbits = g_seq_tables.ml_bits[mlc];
            // order inside a field (first written = lowest bits): OF state, ML state, LL state, LL extra, ML extra, OF extra
            field = stv[lane]; nb = stn[lane];
            field |= (uint64_t)stv[32 + lane] << nb; nb += stn[32 + lane];
            field |= (uint64_t)stv[64 + lane] << nb; nb += stn[64 + lane];
            field |= (uint64_t)(ll & ((1u << llb) - 1)) << nb; nb += llb;
            field |= (uint64_t)(mlb & ((1u << mlbits) - 1)) << nb; nb += mlbits;
            field |= (uint64_t)(offb & ((1u << ofc) - 1)) << nb; nb += ofc;
        }
        const uint32_t inc = warp_inclusive_scan_u32(nb, lane);
        ze_put_bits(bits, bitpos + inc - nb, field, nb);
        bitpos += __shfl_sync(TS_FULL, inc, 31);
        __syncwarp();
    }
    // FSE_flushCState x3 (ML, OF, LL) then the closing 1 bit
    const uint32_t st_of = __shfl_sync(TS_FULL, state, 0), st_ml = __shfl_sync(TS_FULL, state, 1),
                   st_ll = __shfl_sync(TS_FULL, state, 2);
    if (lane == 0) {
        uint64_t f = st_ml & ((1u << kml.log) - 1);
        uint32_t nb = kml.log;
        f |= (uint64_t)(st_of & ((1u << kof.log) - 1)) << nb; nb += kof.log;
        f |= (uint64_t)(st_ll & ((1u << kll.log) - 1)) << nb; nb += kll.log;
        f |= 1ull << nb; nb += 1;
        ze_put_bits(bits, bitpos, f, nb);
    }
    bitpos += kml.log + kof.log + kll.log + 1;
    __syncwarp();
    return (bitpos + 7) >> 3;
}

// warp copy of n bytes, any alignment on either side: aligned 32-bit stores in the middle, source words by funnel shift
__device__ TS_NOINLINE void ze_warp_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n, uint32_t lane) {
    const uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3));
    if (lane < head) dst[lane] = src[lane];
    const uint32_t words = (n - head) >> 2;
    uint32_t* d32 = (uint32_t*)(dst + head);
    const uint8_t* s = src + head;
    for (uint32_t i = lane; i < words; i += 32) d32[i] = ld_u32_unaligned(s + 4 * i);
    const uint32_t done = head + 4 * words;
    if (lane < n - done) dst[done + lane] = src[done + lane];
}

__global__ void __launch_bounds__(ZE_WPB * 32) zstd_enc_blocks_kernel(const __grid_constant__ ZstdEncArgs A) {
#include "zstd_enc_prologue.inc"
#include "zstd_enc_parse.inc"
#include "zstd_enc_emit.inc"
}

// The same block compressor as two launches (TSGPU_ENC_SPLIT=1; off by default): the parse kernel's hot loop is a few KB
// of SASS that stays in the instruction caches, instead of competing with the entropy stage's code in one 90 KB kernel.
// Sequences and literals already travel through global scratch, so the split adds 8 bytes of per-block state.
__global__ void __launch_bounds__(ZE_WPB * 32) zstd_enc_parse_kernel(const __grid_constant__ ZstdEncArgs A) {
#include "zstd_enc_prologue.inc"
#include "zstd_enc_parse.inc"
    (void)fs; (void)last_block; (void)out; (void)stn;              // the prologue is shared with the entropy stage
    if (lane == 0) { A.blk_meta[2 * gblk] = nseq; A.blk_meta[2 * gblk + 1] = nlit; }
}
__global__ void __launch_bounds__(ZE_WPB * 32) zstd_enc_entropy_kernel(const __grid_constant__ ZstdEncArgs A) {
#include "zstd_enc_prologue.inc"
    const uint32_t nseq = A.blk_meta[2 * gblk], nlit = A.blk_meta[2 * gblk + 1];
#include "zstd_enc_emit.inc"
}

// ------------------------------------------------------------------------------------------ frame assembly
__device__ __forceinline__ uint32_t ze_frame_header(uint8_t* h, uint32_t n) {     // see tshost::zstdFrameHeader
    uint32_t p = 0;
    h[p++] = 0x28; h[p++] = 0xB5; h[p++] = 0x2F; h[p++] = 0xFD;
    if (n <= (1u << 21)) {
        if (n < 256) { h[p++] = 0x20; h[p++] = (uint8_t)n; }
        else if (n < 65536 + 256) { h[p++] = 0x60; h[p++] = (uint8_t)(n - 256); h[p++] = (uint8_t)((n - 256) >> 8); }
        else { h[p++] = 0xA0; h[p++] = (uint8_t)n; h[p++] = (uint8_t)(n >> 8); h[p++] = (uint8_t)(n >> 16); h[p++] = (uint8_t)(n >> 24); }
    } else {
        h[p++] = 0x80; h[p++] = 0x58;
        h[p++] = (uint8_t)n; h[p++] = (uint8_t)(n >> 8); h[p++] = (uint8_t)(n >> 16); h[p++] = (uint8_t)(n >> 24);
    }
    return p;
}

__global__ void __launch_bounds__(256) zstd_enc_assemble_kernel(const __grid_constant__ ZstdEncArgs A) {
    __shared__ uint32_t pos[1024 + 1];
    __shared__ uint32_t hdr_len;
    const uint32_t chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t clen = A.in_len[chunk];
    const uint32_t nblk = (clen + ZB - 1) / ZB;
    uint8_t* frame = A.out_base + A.out_off[chunk];
    const uint32_t* bs = A.blk_size + (size_t)chunk * A.blocks_per_chunk;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nblk; base += 1024) {          // chunks of up to 1024 blocks at a time
        const uint32_t cnt = min(1024u, nblk - base);
        if (warp == 0) {
            uint32_t run = carry;
            for (uint32_t b0 = 0; b0 < cnt; b0 += 32) {
                const uint32_t i = b0 + lane;
                const uint32_t v = i < cnt ? bs[base + i] : 0;
                const uint32_t inc = warp_inclusive_scan_u32(v, lane);
                if (i < cnt) pos[i] = run + inc - v;
                run += __shfl_sync(TS_FULL, inc, 31);
            }
            if (lane == 0) {
                pos[cnt] = run;
                if (base == 0) {
                    uint8_t h[16];
                    const uint32_t hl = ze_frame_header(h, clen);
                    for (uint32_t k = 0; k < hl; k++) frame[k] = h[k];
                    hdr_len = hl;
                }
            }
        }
        __syncthreads();
        const uint32_t hl = hdr_len;
        for (uint32_t b = warp; b < cnt; b += blockDim.x >> 5) {
            const uint8_t* s = A.blk_out + ((size_t)chunk * A.blocks_per_chunk + base + b) * ZE_SLOT;
            ze_warp_copy(frame + hl + pos[b], s, pos[b + 1] - pos[b], lane);
        }
        carry = pos[cnt];
        __syncthreads();
    }
    if (tid == 0) {
        uint32_t total = hdr_len + carry;
        if (nblk == 0) {                                           // empty chunk: header + empty last raw block
            uint8_t h[16];
            const uint32_t hl = ze_frame_header(h, 0);
            for (uint32_t k = 0; k < hl; k++) frame[k] = h[k];
            frame[hl] = 1; frame[hl + 1] = 0; frame[hl + 2] = 0;
            total = hl + 3;
        }
        A.out_len[chunk] = total;
    }
}

// ------------------------------------------------------------------------------------------ host side
static thread_local const char* g_zstd_err = "";
inline const char* zstd_last_error() { return g_zstd_err; }

inline const char* zstd_enc_scratch_alloc(ZstdEncScratch& s, uint32_t chunk_cap, uint32_t max_batch) {
    s.blocks_per_chunk = (chunk_cap + ZB - 1) / ZB;
    s.max_batch = max_batch;
    const size_t nblk = (size_t)s.blocks_per_chunk * max_batch;
    const char* e;
    if ((e = rt::malloc_device((void**)&s.blk_out, nblk * ZE_SLOT + 256))) return e;
    if ((e = rt::malloc_device((void**)&s.blk_size, nblk * 4 + 256))) return e;
    if ((e = rt::malloc_device((void**)&s.seqs, nblk * ZE_MAXSEQ * sizeof(uint2) + 256))) return e;
    if ((e = rt::malloc_device((void**)&s.lits, nblk * ZB + 256))) return e;
    if ((e = rt::malloc_device((void**)&s.blk_meta, nblk * 8 + 256))) return e;
    { const char* v = getenv("TSGPU_ENC_SPLIT"); s.split = v && atoi(v) != 0; }
    return nullptr;
}
inline void zstd_enc_scratch_free(ZstdEncScratch& s) {
    rt::free_device(s.blk_out); rt::free_device(s.blk_size); rt::free_device(s.seqs); rt::free_device(s.lits); rt::free_device(s.blk_meta);
    s = ZstdEncScratch{};
}

constexpr uint32_t ZE_SMEM_BYTES = ZE_WPB * ZE_SMEM_WARP_AL;

inline int zstd_compress_batch(ZstdEncScratch& s, rt::stream_t st, const uint8_t* in_base, const uint64_t* d_in_off,
                               const uint32_t* d_in_len, uint32_t n_chunks, uint32_t chunk_size,
                               uint8_t* out_base, const uint64_t* d_out_off, uint32_t* d_out_len, LaunchProf& prof) {
    if (n_chunks > s.max_batch) { g_zstd_err = "batch larger than the context"; return -1; }
    const uint32_t bpc = (chunk_size + ZB - 1) / ZB;
    if (bpc > s.blocks_per_chunk) { g_zstd_err = "chunk larger than the context"; return -1; }
    ZstdEncArgs A;
    A.in_base = in_base; A.in_off = d_in_off; A.in_len = d_in_len;
    A.blk_out = s.blk_out; A.blk_size = s.blk_size; A.seqs = s.seqs; A.lits = s.lits; A.blk_meta = s.blk_meta;
    A.blocks_per_chunk = s.blocks_per_chunk;
    A.out_base = out_base; A.out_off = d_out_off; A.out_len = d_out_len;
    if (bpc) {
        const dim3 grid((bpc + ZE_WPB - 1) / ZE_WPB, n_chunks), block(ZE_WPB * 32);
        if (!s.split) {
            TS_LAUNCH_P(prof, "zstd_enc_blocks", zstd_enc_blocks_kernel, grid, block, ZE_SMEM_BYTES, st, A);
        } else {
            TS_LAUNCH_P(prof, "zstd_enc_parse", zstd_enc_parse_kernel, grid, block, ZE_SMEM_BYTES, st, A);
            TS_LAUNCH_P(prof, "zstd_enc_entropy", zstd_enc_entropy_kernel, grid, block, ZE_SMEM_BYTES, st, A);
        }
        const char* e = rt::last_error();
        if (e) { g_zstd_err = e; return -7; }
    }
    TS_LAUNCH_P(prof, "zstd_enc_assemble", zstd_enc_assemble_kernel, dim3(n_chunks), dim3(256), 0, st, A);
    const char* e = rt::last_error();
    if (e) { g_zstd_err = e; return -7; }
    return 0;
}

}  // namespace ts
