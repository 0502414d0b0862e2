// zstd_enc_blk.cuh — the SPEED mode of the compressor: one warp per independent 8 KiB zstd block.
//
// Same contract as zstd_enc.cuh (one RFC 8878 frame with Frame_Content_Size per chunk, CompressionChunkEnumeration.java:49-62),
// different trade: every 8 KiB block is self-contained (own Huffman tree, own FSE tables, no match leaves the block), so a
// 1 GiB segment is 131,072 independent warps with no barrier between them, 20 warps per SM.  That is the round-1 kernel with
// the streamlined FSE chains and the backwards growth of taken matches ("catch up"); it compresses the K corpus ~3.1 : 1 at
// ~14.7 ms per GiB, where the region kernel of zstd_enc.cuh reaches ~3.5 : 1 at ~27 ms per GiB (one 64 KiB window and one set
// of tables per 8 blocks).  TSGPU_FLAG_ZSTD picks this kernel, TSGPU_FLAG_ZSTD | TSGPU_FLAG_ZSTD_DENSE the region kernel;
// the decoder reads both through the same region path (zstd_dec.cuh).
//   zstd_enc_blocks_kernel    one WARP per block: block staged in shared memory; 32 positions hashed and verified per step
//                             against a per-warp hash table, greedy left-to-right selection by ballot/ffs, match extension
//                             per lane then warp-wide; literals (Raw / RLE / Huffman) + sequences (per-block FSE tables)
//                             placement: look-back over the previous blocks' sizes, the block copies itself into the frame
// Hash-slot winners inside a step are whichever lane the hardware keeps (every outcome is a valid parse: candidates are
// verified before use), so frames of this mode may differ between runs in bytes, never in what they decode to; the region
// kernel is deterministic.
#pragma once
#include "zstd_enc.cuh"

namespace ts {

constexpr int ZB_HLOG = 10;                    // per-warp hash table: 2^10 x u16 (position + 1)
constexpr uint32_t ZB_HSIZE = 1u << ZB_HLOG;
// ZB_TUNE_* exist for A/B runs on the GPU (scripts/ab_variants.sh builds one library per setting); the defaults are the product
#ifndef ZB_TUNE_WPB
#define ZB_TUNE_WPB 4
#endif
#ifndef ZB_TUNE_LANE_EXT
#define ZB_TUNE_LANE_EXT 12
#endif
#ifndef ZB_TUNE_BACK      // 1: taken matches grow backwards over up to 4 literals ("catch up"): +3.6 % ratio for +0.76 ms per GiB (measured)
#define ZB_TUNE_BACK 1
#endif
#ifndef ZB_TUNE_STRAIGHT  // 1: per-lane match extension as straight-line code over 12 bytes instead of a divergent loop:
#define ZB_TUNE_STRAIGHT 1 //   16.21 -> 14.72 ms per GiB on the B200 (the loop was 16 % of the kernel's stall samples; profiles/r02_ab.md)
#endif
#ifndef ZB_TUNE_EXTLDS    // 1: the warp-wide extension of long matches reads shared memory with LDS word pairs instead of generic loads
#define ZB_TUNE_EXTLDS 0  //    (measured on the B200: 14.92 instead of 14.72 ms per GiB — off; profiles/r02_ab.md)
#endif
constexpr int ZB_WPB = ZB_TUNE_WPB;            // warps (= blocks in flight) per CTA
constexpr uint32_t ZB_LANE_EXT = ZB_TUNE_LANE_EXT;   // bytes a lane extends its own match beyond the first 4
constexpr uint32_t ZB_BUF_PAD = 416;           // zero pad for over-reads; its tail also holds the FSE tile scratch
constexpr uint32_t ZB_SLOT = ZB + 512;         // per-block output slot: 3-byte header + payload
constexpr uint32_t ZB_SMEM_WARP = ZB + ZB_BUF_PAD + ZB_HSIZE * 2;   // buf, ht
constexpr uint32_t ZB_SMEM_WARP_AL = (ZB_SMEM_WARP + 15) & ~15u;
// per-tile FSE scratch lives in the tail of `buf`: the sequence bit stream staged there is at most ZE_MAXSEQ * 58 bits
constexpr uint32_t ZB_SEQ_AUX_OFF = 7456;
static_assert(ZB_SEQ_AUX_OFF + 3 * 32 * 8 + 3 * 32 * 4 <= ZB + ZB_BUF_PAD && ZB_SEQ_AUX_OFF % 8 == 0, "FSE tile scratch must fit the block buffer");
static_assert(ZE_MAXSEQ * 58 / 8 + 16 <= 7456, "sequence bit stream bound");
struct ZbCTab {                                // per-warp FSE encoding tables (logs <= 7 for <= 1024 sequences); aliases the hash-table area
    uint16_t st_ll[128], st_ml[128], st_of[128];
    zf::FseCSym sy_ll[ZE_NSYM_LL], sy_ml[ZE_NSYM_ML], sy_of[ZE_NSYM_OF];
};
static_assert(sizeof(ZbCTab) <= ZB_HSIZE * 2 && 2048 <= ZB_HSIZE * 2, "FSE tables and the 2 KiB Huffman scratch alias the hash-table area");
struct ZbFsePre { zf::PredefinedCTables t; };

__device__ __forceinline__ uint32_t zb_hash(uint32_t v) { return (v * 2654435761u) >> (32 - ZB_HLOG); }
__device__ __forceinline__ uint32_t zb_ll_code(uint32_t ll) { return ll < 64 ? g_seq_tables.ll_code[ll] : (uint32_t)zf::highbit32(ll) + 19; }
__device__ __forceinline__ uint32_t zb_ml_code(uint32_t mlbase) { return mlbase < 128 ? g_seq_tables.ml_code[mlbase] : (uint32_t)zf::highbit32(mlbase) + 36; }

struct ZstdBlkScratch {
    uint8_t* blk_out = nullptr;      // blocks * ZB_SLOT
    uint32_t* blk_size = nullptr;    // blocks
    uint2* seqs = nullptr;           // blocks * ZE_MAXSEQ
    uint8_t* lits = nullptr;         // blocks * ZB
    unsigned long long* blk_state = nullptr;   // blocks: look-back words (ready bit 63 | inclusive frame bytes)
    uint32_t blocks_per_chunk = 0, max_batch = 0;
};
struct ZstdBlkArgs {
    const uint8_t* in_base; const uint64_t* in_off; const uint32_t* in_len;
    uint8_t* blk_out; uint32_t* blk_size; uint2* seqs; uint8_t* lits; unsigned long long* blk_state;
    uint32_t blocks_per_chunk;
    uint8_t* out_base; const uint64_t* out_off; uint32_t* out_len;
};

// Raw_Literals_Block with the 3-byte header (Size_Format 11: 20-bit Regenerated_Size).
__device__ TS_NOINLINE uint32_t zb_raw_literals(const uint8_t* __restrict__ lits, uint32_t n, uint8_t* body, uint32_t lane) {
    if (lane == 0) {
        body[0] = (uint8_t)((3u << 2) | ((n & 0xf) << 4));
        body[1] = (uint8_t)(n >> 4);
        body[2] = (uint8_t)(n >> 12);
    }
    for (uint32_t i = lane; i < n; i += 32) body[3 + i] = lits[i];
    return 3 + n;
}
__device__ __forceinline__ uint32_t zb_rle_literals(uint8_t v, uint32_t n, uint8_t* body, uint32_t lane) {
    if (lane == 0) {
        body[0] = (uint8_t)(1u | (3u << 2) | ((n & 0xf) << 4));
        body[1] = (uint8_t)(n >> 4);
        body[2] = (uint8_t)(n >> 12);
        body[3] = v;
    }
    return 4;
}


// Returns the bytes written at `body`.  `work`: >= ZB + 160 bytes of shared memory (tree scratch, then the
// stream staging area); `aux16`: 2 KiB of shared memory (histogram + code table).
__device__ __forceinline__ uint32_t zb_encode_literals(const uint8_t* __restrict__ lits, uint32_t n, uint8_t* body,
                                                       uint32_t* work, uint16_t* aux16, uint32_t lane) {
    if (n < ZE_HUF_MIN) return zb_raw_literals(lits, n, body, lane);
    uint32_t* hist = (uint32_t*)aux16;         // [256]
    uint32_t* ctab = hist + 256;               // [256] code | len << 16
    uint32_t* keys = work + 1280;              // [256] used symbols: count << 8 | symbol
    uint32_t* sorted = work;                   // [256]
    uint32_t* nodew = work + 256;              // [512]
    uint16_t* parent = (uint16_t*)(work + 768);   // [512]
    uint8_t* depth = (uint8_t*)(work + 1024);  // [512]
    uint32_t* meta = work + 1200;              // small scalars shared by the warp

    for (uint32_t i = lane; i < 256; i += 32) { hist[i] = 0; ctab[i] = 0; }
    __syncwarp();
    _Pragma("unroll 2")
    for (uint32_t i = lane; i < n; i += 32) atomicAdd(&hist[lits[i]], 1u);
    __syncwarp();

    // compact the used symbols (8 per lane, ascending symbol order)
    uint32_t mine = 0;
    for (uint32_t k = 0; k < 8; k++) mine += hist[lane * 8 + k] ? 1u : 0u;
    const uint32_t inc = warp_inclusive_scan_u32(mine, lane);
    const uint32_t m = __shfl_sync(TS_FULL, inc, 31);
    {
        uint32_t at = inc - mine;
        for (uint32_t k = 0; k < 8; k++) { const uint32_t s = lane * 8 + k; if (hist[s]) keys[at++] = (hist[s] << 8) | s; }
    }
    __syncwarp();
    if (m == 1) return zb_rle_literals((uint8_t)(keys[0] & 0xff), n, body, lane);
    {   // Shannon estimate in 1/16 bit units: sum c * (log2(n) - log2(c)), log2 by leading zeros + a linear fraction
        uint32_t cost = 0;
        for (uint32_t k = 0; k < 8; k++) {
            const uint32_t c = hist[lane * 8 + k];
            if (c) {
                const uint32_t hb = (uint32_t)zf::highbit32(c);
                const uint32_t lg16 = (hb << 4) + (((c << (31 - hb)) >> 27) & 15);       // ~ 16 * log2(c)
                cost += c * (((uint32_t)zf::highbit32(n) << 4) + (((n << (31 - zf::highbit32(n))) >> 27) & 15) + 1 - lg16);
            }
        }
        const uint32_t bits16 = __reduce_add_sync(TS_FULL, cost);
        if ((bits16 >> 7) + m / 2 + 16 >= n) return zb_raw_literals(lits, n, body, lane);   // >= n bytes even before rounding losses
    }
    const uint32_t last_sym = keys[m - 1] & 0xff;

    // ---- code lengths: rebuild with halved counts until the tree is at most 11 deep
    uint32_t max_len = 0;
    for (uint32_t round = 0; round < 16; round++) {
        for (uint32_t e = lane; e < m; e += 32) {                         // rank sort (keys are distinct)
            const uint32_t key = keys[e];
            uint32_t r = 0;
            _Pragma("unroll 2")
            for (uint32_t j = 0; j < m; j++) r += keys[j] < key ? 1u : 0u;
            sorted[r] = key;
        }
        __syncwarp();
        if (lane == 0) {
            _Pragma("unroll 1")
            for (uint32_t i = 0; i < m; i++) nodew[i] = sorted[i] >> 8;
            uint32_t li = 0, ii = m, ni = m;
            _Pragma("unroll 1")
            for (uint32_t k = 0; k + 1 < m; k++) {
                uint32_t a, b;
                if (li < m && (ii >= ni || nodew[li] <= nodew[ii])) a = li++; else a = ii++;
                if (li < m && (ii >= ni || nodew[li] <= nodew[ii])) b = li++; else b = ii++;
                nodew[ni] = nodew[a] + nodew[b];
                parent[a] = (uint16_t)ni; parent[b] = (uint16_t)ni;
                ni++;
            }
            const uint32_t root = 2 * m - 2;
            depth[root] = 0;
            uint32_t mx = 0;
            _Pragma("unroll 1")
            for (int32_t i = (int32_t)root - 1; i >= 0; i--) {
                depth[i] = (uint8_t)(depth[parent[i]] + 1);
                if ((uint32_t)i < m && depth[i] > mx) mx = depth[i];
            }
            meta[0] = mx;
        }
        __syncwarp();
        max_len = meta[0];
        if (max_len <= (uint32_t)zf::HUF_MAX_LOG) break;
        for (uint32_t e = lane; e < m; e += 32) {                         // flatten the distribution and retry
            const uint32_t key = keys[e];
            keys[e] = ((((key >> 8) + 1) >> 1) << 8) | (key & 0xff);
        }
        __syncwarp();
    }
    if (max_len > (uint32_t)zf::HUF_MAX_LOG) return zb_raw_literals(lits, n, body, lane);

    // ---- canonical codes: weight w = max_len + 1 - len; cells are dealt weight-ascending, symbol-ascending
    if (lane == 0) {
        uint32_t cnt[16];
        for (int w = 0; w < 16; w++) cnt[w] = 0;
        _Pragma("unroll 1")
        for (uint32_t i = 0; i < m; i++) cnt[max_len + 1 - depth[i]]++;
        uint32_t start[16], pos = 0;
        for (uint32_t w = 1; w <= max_len; w++) { start[w] = pos; pos += cnt[w] << (w - 1); }
        // sorted[] is ordered by count; symbol order inside a weight class comes from walking symbols ascending
        _Pragma("unroll 1")
        for (uint32_t i = 0; i < m; i++) ctab[sorted[i] & 0xff] = (uint32_t)depth[i] << 16;      // park the length
        uint64_t total_bits = 0;
        _Pragma("unroll 1")
        for (uint32_t s = 0; s <= last_sym; s++) {
            const uint32_t len = ctab[s] >> 16;
            if (!len) continue;
            const uint32_t w = max_len + 1 - len;
            ctab[s] = (start[w] >> (w - 1)) | (len << 16);
            start[w] += 1u << (w - 1);
            total_bits += (uint64_t)len * hist[s];
        }
        meta[1] = (uint32_t)total_bits;
    }
    __syncwarp();
    const uint32_t nweights = last_sym;                                   // symbols 0 .. last_sym-1 are listed
    // Tree description: direct 4-bit weights (at most 128 of them) or FSE-compressed weights (RFC 8878 §4.2.1.1);
    // the shorter wins.  The FSE form is what lets alphabets above byte value 128 (binary payloads) be Huffman-coded.
    uint32_t tree_bytes = nweights <= 128 ? 1 + (nweights + 1) / 2 : 0xffffffffu;
    uint8_t* wdesc = (uint8_t*)hist;                                      // the histogram is dead: header byte + FSE description
    {
        uint8_t* wts = (uint8_t*)(work + 1536);                           // weights, then FSE tables, in dead tree scratch
        uint32_t* wcnt = (uint32_t*)(wts + 256);                          // [16]
        uint16_t* wst = (uint16_t*)(wcnt + 16);                           // [64]
        zf::FseCSym* wsy = (zf::FseCSym*)(wst + 64);                      // [16]
        uint8_t* wscratch = (uint8_t*)(wsy + 16);                         // >= 772 bytes
        if (lane < 16) wcnt[lane] = 0;
        __syncwarp();
        for (uint32_t s2 = lane; s2 < nweights; s2 += 32) {
            const uint32_t l = ctab[s2] >> 16;
            const uint32_t w = l ? max_len + 1 - l : 0;
            wts[s2] = (uint8_t)w;
            atomicAdd(&wcnt[w], 1u);
        }
        __syncwarp();
        if (nweights >= 2 && (ZE_FSE_WEIGHTS_ALWAYS || nweights > 128)) {
            ZeKind wk;
            const uint32_t dsz = ze_build_kind(wcnt, 13, nweights, zf::HUFW_MAX_LOG, zf::HUFW_MAX_LOG, nullptr, nullptr, wst, wsy,
                                               wscratch, wdesc + 1, &wk, false, lane);
            if (wk.mode == 2) {
                if (lane == 0) {                                          // two interleaved states, last weight first
                    uint8_t* o = wdesc + 1 + dsz;
                    uint64_t acc = 0; uint32_t nb = 0, ob = 0;
                    uint32_t st[2];
                    for (uint32_t q = 0; q < 2; q++) {                    // FSE_initCState2 for the two last weights
                        const uint32_t i = nweights - 1 - q;
                        const zf::FseCSym c = wsy[wts[i]];
                        const uint32_t nbo = (uint32_t)(c.delta_nb_bits + (1 << 15)) >> 16;
                        const uint32_t v = (nbo << 16) - (uint32_t)c.delta_nb_bits;
                        st[i & 1] = wst[(int32_t)(v >> nbo) + c.delta_find_state];
                    }
                    for (int32_t i = (int32_t)nweights - 3; i >= 0; i--) {
                        const zf::FseCSym c = wsy[wts[i]];
                        const uint32_t sv = st[i & 1];
                        const uint32_t nbo = (sv + (uint32_t)c.delta_nb_bits) >> 16;
                        acc |= (uint64_t)(sv & ((1u << nbo) - 1)) << nb; nb += nbo;
                        st[i & 1] = wst[(int32_t)(sv >> nbo) + c.delta_find_state];
                        while (nb >= 8) { o[ob++] = (uint8_t)acc; acc >>= 8; nb -= 8; }
                    }
                    acc |= (uint64_t)(st[1] & ((1u << wk.log) - 1)) << nb; nb += wk.log;      // flush odd chain, then even chain
                    acc |= (uint64_t)(st[0] & ((1u << wk.log) - 1)) << nb; nb += wk.log;
                    acc |= 1ull << nb; nb += 1;                                                // end mark
                    while (nb > 0) { o[ob++] = (uint8_t)acc; acc >>= 8; nb = nb >= 8 ? nb - 8 : 0; }
                    meta[2] = dsz + ob;
                }
                __syncwarp();
                const uint32_t fsz = meta[2];
                if (fsz < 128 && 1 + fsz < tree_bytes) {
                    if (lane == 0) wdesc[0] = (uint8_t)fsz;
                    tree_bytes = 1 + fsz;
                } else if (lane == 0) wdesc[0] = 0xff;
            } else if (lane == 0) wdesc[0] = 0xff;
        } else if (lane == 0) wdesc[0] = 0xff;
        __syncwarp();
    }
    if (tree_bytes == 0xffffffffu) return zb_raw_literals(lits, n, body, lane);   // neither form can describe this tree
    const bool fse_weights = wdesc[0] != 0xff;
    const uint32_t est = tree_bytes + 6 + (meta[1] >> 3) + 8;
    if (est + 5 >= n) return zb_raw_literals(lits, n, body, lane);        // Huffman would not pay

    // ---- the four streams, staged in `work` (cleared first; tree scratch is dead from here on)
    const uint32_t seg = (n + 3) / 4;
    __syncwarp();
    for (uint32_t i = lane; i < (ZB + 128) / 4; i += 32) work[i] = 0;
    __syncwarp();
    uint32_t byte_pos = 0;
    uint64_t ssz_all = 0;                                                 // four 16-bit stream sizes
    _Pragma("unroll 1")
    for (uint32_t st = 0; st < 4; st++) {
        const uint32_t s0 = st * seg, s1 = st < 3 ? min(n, s0 + seg) : n;
        const uint32_t cnt = s1 > s0 ? s1 - s0 : 0;
        const uint32_t per = (cnt + 31) / 32;
        const uint32_t a = min(cnt, lane * per), b = min(cnt, a + per);   // this lane's run [a, b) of the stream
        uint32_t mybits = 0;
        _Pragma("unroll 2")
        for (uint32_t i = a; i < b; i++) mybits += ctab[lits[s0 + i]] >> 16;
        // symbols are written last-to-first: the offset of a run is the number of bits of all LATER runs
        const uint32_t incb = warp_inclusive_scan_u32(mybits, lane);
        const uint32_t total = __shfl_sync(TS_FULL, incb, 31);
        uint32_t off = byte_pos * 8 + (total - incb);
        for (uint32_t i = b; i > a; i--) {
            const uint32_t c = ctab[lits[s0 + i - 1]];
            const uint32_t len = c >> 16, code = c & 0xffff;
            const uint32_t w = off >> 5, sh = off & 31;
            atomicOr(&work[w], code << sh);
            if (sh + len > 32) atomicOr(&work[w + 1], code >> (32 - sh));
            off += len;
        }
        if (lane == 0) {                                                  // end mark
            const uint32_t o = byte_pos * 8 + total;
            atomicOr(&work[o >> 5], 1u << (o & 31));
        }
        const uint32_t sz = (total + 1 + 7) >> 3;
        ssz_all |= (uint64_t)sz << (16 * st);
        byte_pos += sz;
        __syncwarp();
    }
    const uint32_t comp = tree_bytes + 6 + byte_pos;
    const uint32_t hsz = (n < 1024 && comp < 1024) ? 3u : (n < 16384 && comp < 16384) ? 4u : 5u;
    if (hsz + comp >= 3 + n) return zb_raw_literals(lits, n, body, lane);
    if (lane == 0) {
        const uint32_t sf = hsz - 2;                                      // 1, 2, 3: all with four streams
        uint64_t h;
        if (hsz == 3) h = 2u | (sf << 2) | ((uint64_t)n << 4) | ((uint64_t)comp << 14);
        else if (hsz == 4) h = 2u | (sf << 2) | ((uint64_t)n << 4) | ((uint64_t)comp << 18);
        else h = 2u | (sf << 2) | ((uint64_t)n << 4) | ((uint64_t)comp << 22);
        for (uint32_t k = 0; k < hsz; k++) body[k] = (uint8_t)(h >> (8 * k));
        uint8_t* t = body + hsz;
        if (fse_weights) {
            for (uint32_t i = 0; i < tree_bytes; i++) t[i] = wdesc[i];
        } else {
            t[0] = (uint8_t)(127 + nweights);
            for (uint32_t i = 0; i < nweights; i += 2) {
                const uint32_t l0 = ctab[i] >> 16, l1 = i + 1 < nweights ? ctab[i + 1] >> 16 : 0;
                const uint32_t w0 = l0 ? max_len + 1 - l0 : 0, w1 = l1 ? max_len + 1 - l1 : 0;
                t[1 + i / 2] = (uint8_t)((w0 << 4) | w1);
            }
        }
        uint8_t* j = t + tree_bytes;
        for (uint32_t k = 0; k < 6; k++) j[k] = (uint8_t)(ssz_all >> (8 * k));            // jump table: sizes of streams 1-3
    }
    uint8_t* sp = body + hsz + tree_bytes + 6;
    const uint8_t* wb = (const uint8_t*)work;
    _Pragma("unroll 2")
    for (uint32_t i = lane; i < byte_pos; i += 32) sp[i] = wb[i];
    __syncwarp();
    return hsz + comp;
}


// Encodes the sequences section (everything after the Number_of_Sequences field): the modes byte and table
// descriptions go straight to `hdr_out` (global), the bit stream is staged in the word buffer `bits` (shared).
// Returns the bit-stream bytes; *desc_bytes = 1 (modes byte) + table descriptions.  Warp-uniform, N >= 1.
__device__ __forceinline__ uint32_t zb_encode_sequences(const uint2* __restrict__ seqs, uint32_t N, uint32_t* bits, ZbCTab* ct,
                                                        uint2* pk /*[3][32]*/, uint32_t* sb /*[3][32]*/,
                                                        const ZbFsePre* fs, uint8_t* hdr_out, uint32_t* desc_bytes, uint32_t lane) {
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
            atomicAdd(&cnt[zb_ll_code(s.x & 0xffff)], 1u);
            atomicAdd(&cnt[ZE_NSYM_LL + zb_ml_code(s.x >> 16)], 1u);
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
    for (uint32_t i = lane; i < (ZB + ZB_BUF_PAD) / 4; i += 32) bits[i] = 0;
    __syncwarp();

    // ---- pass 2: encode.  Per tile of 32 sequences every lane looks up the transforms of its sequence's three codes;
    // lanes 0-2 then walk the three state chains (one 8-byte load, the bit count, one store, one table load per symbol).
    uint32_t bitpos = 0;
    uint32_t state = 0;                              // lanes 0..2: OF, ML, LL chains
    const uint16_t* st_tab = lane == 0 ? ct->st_of : lane == 1 ? ct->st_ml : ct->st_ll;
    const bool rle = (lane == 0 ? kof.mode : lane == 1 ? kml.mode : kll.mode) == 1;
    const uint2* mypk = pk + (lane < 3 ? lane : 0) * 32;
    uint32_t* mysb = sb + (lane < 3 ? lane : 0) * 32;
    pre = lane < N ? seqs[N - 1 - lane] : make_uint2(0, 0);
    for (uint32_t t0 = 0; t0 < N; t0 += 32) {
        const uint32_t j = t0 + lane;                // stream order: j = 0 is the LAST sequence
        const bool have = j < N;
        uint32_t ll = 0, mlb = 0, offb = 0, llc = 0, mlc = 0, ofc = 0;
        const uint2 s = pre;
        if (j + 32 < N) pre = seqs[N - 1 - (j + 32)];
        if (have) {
            ll = s.x & 0xffff; mlb = s.x >> 16; offb = s.y + 3;       // mlb = matchLength - 3, offb = offset + 3 (no repcodes)
            llc = zb_ll_code(ll); mlc = zb_ml_code(mlb); ofc = (uint32_t)zf::highbit32(offb);
            const zf::FseCSym a = ct->sy_of[ofc], b = ct->sy_ml[mlc], c = ct->sy_ll[llc];
            pk[lane] = make_uint2((uint32_t)a.delta_nb_bits, (uint32_t)a.delta_find_state);
            pk[32 + lane] = make_uint2((uint32_t)b.delta_nb_bits, (uint32_t)b.delta_find_state);
            pk[64 + lane] = make_uint2((uint32_t)c.delta_nb_bits, (uint32_t)c.delta_find_state);
        }
        __syncwarp();
        if (lane < 3) {
            const uint32_t cnt = min(32u, N - t0);
            if (rle) {
                for (uint32_t i = 0; i < cnt; i++) mysb[i] = 0;
            } else {
                uint32_t i = 0;
                if (t0 == 0) {                       // FSE_initCState2 with the last sequence's symbol
                    const uint2 c = mypk[0];
                    const uint32_t nb = (uint32_t)((int32_t)c.x + (1 << 15)) >> 16;
                    const uint32_t v = (nb << 16) - c.x;
                    state = st_tab[(int32_t)(v >> nb) + (int32_t)c.y];
                    mysb[0] = 0;
                    i = 1;
                }
                _Pragma("unroll 4")
                for (; i < cnt; i++) {               // FSE_encodeSymbol
                    const uint2 c = mypk[i];
                    const uint32_t nb = (state + c.x) >> 16;
                    mysb[i] = (state & ((1u << nb) - 1)) | (nb << 16);
                    state = st_tab[(int32_t)(state >> nb) + (int32_t)c.y];
                }
            }
        }
        __syncwarp();
        uint64_t field = 0; uint32_t nb = 0;
        if (have) {
            const uint32_t llb = g_seq_tables.ll_bits[llc], mlbits = g_seq_tables.ml_bits[mlc];
            // order inside a field (first written = lowest bits): OF state, ML state, LL state, LL extra, ML extra, OF extra
            const uint32_t b0 = sb[lane], b1 = sb[32 + lane], b2 = sb[64 + lane];
            field = b0 & 0xffff; nb = b0 >> 16;
            field |= (uint64_t)(b1 & 0xffff) << nb; nb += b1 >> 16;
            field |= (uint64_t)(b2 & 0xffff) << nb; nb += b2 >> 16;
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


__global__ void __launch_bounds__(ZB_WPB * 32) zstd_enc_blocks_kernel(const __grid_constant__ ZstdBlkArgs A) {
// Common prologue of the block kernels (included inside a __global__ function body; see zstd_enc.cuh): which block this
// warp owns, its scratch slots and its shared-memory window.
    TS_DYN_SMEM(smem);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ZbFsePre& fs = *(const ZbFsePre*)&g_pre_ctables;   // predefined tables stay in constant memory

    const uint32_t chunk = blockIdx.y;
    const uint32_t blk = blockIdx.x * ZB_WPB + warp;
    const uint32_t clen = A.in_len[chunk];
    if (clen == 0 && blk == 0) {                                  // empty chunk: header + empty last raw block
        if (lane == 0) {
            uint8_t* frame = A.out_base + A.out_off[chunk];
            const uint32_t hl = ze_frame_header(frame, 0);
            frame[hl] = 1; frame[hl + 1] = 0; frame[hl + 2] = 0;
            A.out_len[chunk] = hl + 3;
        }
        return;
    }
    if ((uint64_t)blk * ZB >= clen) return;                       // whole warp
    const uint32_t bn = min(ZB, clen - blk * ZB);
    const bool last_block = (uint64_t)(blk + 1) * ZB >= clen;
    const uint8_t* src = A.in_base + A.in_off[chunk] + (size_t)blk * ZB;
    const size_t gblk = (size_t)chunk * A.blocks_per_chunk + blk;
    uint8_t* out = A.blk_out + gblk * ZB_SLOT;
    uint2* seqs = A.seqs + gblk * ZE_MAXSEQ;
    uint8_t* lits = A.lits + gblk * ZB;

    uint8_t* wbase = smem + warp * ZB_SMEM_WARP_AL;
    uint8_t* buf = wbase;
    uint16_t* ht = (uint16_t*)(wbase + ZB + ZB_BUF_PAD);
    uint2* pk = (uint2*)(buf + ZB_SEQ_AUX_OFF);
    uint32_t* sb = (uint32_t*)(pk + 96);
// Stage + LZ parse of one 8 KiB block (included inside a __global__ function body after
// zstd_enc_prologue.inc).  Leaves nseq sequences in `seqs`, nlit literals in `lits`.
    // ---- stage the block in shared memory (128-bit loads when the source is aligned), zero the pad, reset the table
    if ((((uintptr_t)src) & 15) == 0) {
        for (uint32_t i = lane * 16; i < bn; i += 512) {
            if (i + 16 <= bn) *(uint4*)(buf + i) = ldg128_stream((const uint4*)(src + i));
            else for (uint32_t k = i; k < bn; k++) buf[k] = src[k];
        }
    } else {
        _Pragma("unroll 2")
        for (uint32_t i = lane; i < bn; i += 32) buf[i] = src[i];
    }
    _Pragma("unroll 1")
    for (uint32_t i = bn + lane; i < ZB + ZB_BUF_PAD; i += 32) buf[i] = 0;
    for (uint32_t i = lane; i < ZB_HSIZE / 2; i += 32) ((uint32_t*)ht)[i] = 0;
    __syncwarp();

    // ---- phase A: greedy LZ parse, 32 positions per step
    // Selection (which of the 32 candidate matches survive, left to right) is the only serial part and costs a
    // handful of instructions per taken match; sequences are then written by their own lanes in parallel and
    // the step's literals (the positions no taken match covers) leave in the same step.
    uint32_t anchor = 0, cur = 0, nseq = 0, nlit = 0;
    while (cur + 4 <= bn && nseq + 8 <= ZE_MAXSEQ) {                // a step adds at most 8 sequences (min match 4)
        const uint32_t p = cur + lane;
        const bool valid = p + 4 <= bn;
        // unaligned 4-byte reads as a rolling pair of aligned words per stream: one new LDS per stream and step
        const uint32_t* wp = (const uint32_t*)(buf + (p & ~3u));
        const uint32_t shp = (p & 3) * 8;
        uint32_t a0 = wp[0], a1 = wp[1];
        const uint32_t v = __funnelshift_r(a0, a1, shp);
        const uint32_t h = zb_hash(v);
        const uint32_t slot = valid ? ht[h] : 0u;                  // position + 1, 0 = empty
        __syncwarp();
        // Lanes of this step that share a slot all store; the CUDA model lets any one of them win.  Every outcome is a
        // valid parse (candidates are verified before use), so frames may differ in bytes, never in what they decode to.
        if (valid) ht[h] = (uint16_t)(p + 1);
        __syncwarp();
        const uint32_t cand = slot ? slot - 1 : 0u;
        const uint32_t* wc = (const uint32_t*)(buf + (cand & ~3u));
        const uint32_t shc = (cand & 3) * 8;
        uint32_t c0 = wc[0], c1 = wc[1];
        const bool ok = slot != 0 && __funnelshift_r(c0, c1, shc) == v;
#if ZB_TUNE_BACK
        // how many of the 4 bytes before the position also match (straight-line, all lanes: two loads, two funnel shifts):
        // a taken match grows backwards over the literals before it, libzstd's "catch up".  (A first version did this in a
        // branch of the lanes with a verified candidate, through ld_u32_unaligned: 1.55 ms per GiB instead of 0.76.)
        const uint32_t am = p >= 4 ? wp[-1] : 0u, cm = cand >= 4 ? wc[-1] : 0u;
        const uint32_t bkr = min((uint32_t)__clz((int)(__funnelshift_r(am, a0, shp) ^ __funnelshift_r(cm, c0, shc))) >> 3, min(cand, 4u));
#else
        const uint32_t bkr = 0;
#endif
        uint32_t len = 0;
#if ZB_TUNE_STRAIGHT
        {   // the 12 bytes after the verified 4, straight-line on all lanes (three more words per stream): no divergent loop
            static_assert(ZB_TUNE_LANE_EXT == 12, "the straight-line extension measures exactly 12 bytes");
            const uint32_t a2 = wp[2], a3 = wp[3], a4 = wp[4], c2 = wc[2], c3 = wc[3], c4 = wc[4];
            const uint32_t x1 = __funnelshift_r(a1, a2, shp) ^ __funnelshift_r(c1, c2, shc);
            const uint32_t x2 = __funnelshift_r(a2, a3, shp) ^ __funnelshift_r(c2, c3, shc);
            const uint32_t x3 = __funnelshift_r(a3, a4, shp) ^ __funnelshift_r(c3, c4, shc);
            const uint32_t e1 = ze_common_bytes(x1), e2 = ze_common_bytes(x2), e3 = ze_common_bytes(x3);   // 0..4 each
            const uint32_t e = e1 + (e1 >> 2) * (e2 + (e2 >> 2) * e3);         // branch-free: a word counts only if the one before it matched whole
            len = ok ? min(4 + e, min(bn - p, 4 + ZB_LANE_EXT)) : 0u;
        }
#else
        if (ok) {
            len = 4;
            const uint32_t lim = min(bn - p, 4 + ZB_LANE_EXT);
            for (uint32_t k = 2; len < lim; k++) {
                a0 = a1; a1 = wp[k]; c0 = c1; c1 = wc[k];
                const uint32_t c = ze_common_bytes(__funnelshift_r(a0, a1, shp) ^ __funnelshift_r(c0, c1, shc));
                len += c;
                if (c < 4) break;
            }
            len = min(len, lim);
        }
#endif
        const uint32_t mask = __ballot_sync(TS_FULL, ok && len >= ZE_MIN_MATCH);
        // Greedy selection, left to right.  Every lane precomputes where its match would end and which candidate would
        // come next, so one shuffle per taken match walks the chain (the only serial part of the parse).
        const uint32_t e = lane + len;
        const uint32_t mnext = e < 32 ? mask & (0xffffffffu << e) : 0u;
        const uint32_t nxt = mnext ? (uint32_t)__ffs((int)mnext) - 1 : 32u;
        const bool capped = ok && len == 4 + ZB_LANE_EXT && p + len < bn;
        const uint32_t packed = e | (nxt << 8) | (capped ? 1u << 16 : 0u);
        uint32_t taken = 0, pos = 0;
        uint32_t f = mask ? (uint32_t)__ffs((int)mask) - 1 : 32u;
        while (f < 32) {
            const uint32_t info = __shfl_sync(TS_FULL, packed, f);
            uint32_t end = info & 0xffu, nf = (info >> 8) & 0xffu;
            if (info >> 16) {                                      // warp-wide extension of a long match
                uint32_t L = 4 + ZB_LANE_EXT;
                const uint32_t off = __shfl_sync(TS_FULL, p - cand, f);
                const uint32_t mpos = cur + f;
                while (true) {
                    const uint32_t q = mpos + L + 4 * lane;
                    uint32_t c = 0;
                    if (q < bn) {
#if ZB_TUNE_EXTLDS      // shared-memory word pairs + funnel shifts (ld_u32_unaligned goes through a generic pointer: LD.E, not LDS)
                        const uint32_t r = q - off;
                        const uint32_t* wq = (const uint32_t*)(buf + (q & ~3u));
                        const uint32_t* wr = (const uint32_t*)(buf + (r & ~3u));
                        c = ze_common_bytes(__funnelshift_r(wq[0], wq[1], (q & 3) * 8) ^ __funnelshift_r(wr[0], wr[1], (r & 3) * 8));
#else
                        c = ze_common_bytes(ld_u32_unaligned(buf + q) ^ ld_u32_unaligned(buf + q - off));
#endif
                        c = min(c, bn - q);
                    }
                    const uint32_t stop = __ballot_sync(TS_FULL, c < 4);
                    if (stop) {
                        const uint32_t fl = (uint32_t)__ffs((int)stop) - 1;
                        L += 4 * fl + __shfl_sync(TS_FULL, c, fl);
                        break;
                    }
                    L += 128;
                }
                if (lane == f) len = L;
                end = f + L;
                const uint32_t m2 = end < 32 ? mask & (0xffffffffu << end) : 0u;
                nf = m2 ? (uint32_t)__ffs((int)m2) - 1 : 32u;
            }
            taken |= 1u << f;
            pos = end;
            f = nf;
        }
        uint32_t cov = 0;                                          // positions of this step covered by a taken match
        if (taken) {
            const bool mine_taken = (taken >> lane) & 1;
            const uint32_t my_end = p + len;                       // meaningful on taken lanes
            const uint32_t lower = taken & ((1u << lane) - 1);
            const uint32_t prev_lane = lower ? (uint32_t)(31 - __clz((int)lower)) : 0u;
            uint32_t prev_end = __shfl_sync(TS_FULL, my_end, prev_lane);
            if (!lower) prev_end = anchor;
            const uint32_t bk = mine_taken ? min(bkr, p - max(prev_end, cur)) : 0u;   // "catch up" over this step's literals, never into the previous match
            if (mine_taken)
                seqs[nseq + (uint32_t)__popc(lower)] = make_uint2((p - bk - prev_end) | ((len + bk - 3) << 16), p - cand);
            nseq += (uint32_t)__popc(taken);
            anchor = cur + pos;
            const uint32_t sl = lane - bk, tl = len + bk;
            cov = __reduce_or_sync(TS_FULL, mine_taken ? (tl >= 32 - sl ? 0xffffffffu : (1u << tl) - 1) << sl : 0u);
        }
        // the step's literals, in order: one byte per position of the block no taken match covers
        {
            const uint32_t inside = bn - cur >= 32 ? 0xffffffffu : (1u << (bn - cur)) - 1;
            const uint32_t lm = ~cov & inside;
            if ((lm >> lane) & 1) lits[nlit + (uint32_t)__popc(lm & ((1u << lane) - 1))] = (uint8_t)v;
            nlit += (uint32_t)__popc(lm);
        }
        cur = max(cur + 32, anchor);
    }
    if (cur < bn) {                                                // the rest of the block is literals
        const uint32_t ll = bn - cur;
        for (uint32_t k = lane; k < ll; k += 32) lits[nlit + k] = buf[cur + k];
        nlit += ll;
    }
    __syncwarp();
    __threadfence_block();
// Entropy stage + emission of one block (included inside a __global__ function body after zstd_enc_prologue.inc; needs
// nseq / nlit).  `buf` and `ht` are scratch here: the block's bytes are no longer needed in shared memory.
    // ---- phase B: entropy stage into the (now free) shared block buffer, then emit
    uint32_t payload = 0xffffffffu;                                // "not compressible"
    {
        // literals section first (into `out` directly), then the sequences bit stream staged in `buf`
        uint8_t* body = out + 3;
        const uint32_t lit_bytes = zb_encode_literals(lits, nlit, body, (uint32_t*)buf, ht, lane);   // header + payload
        __syncwarp();
        if (nseq == 0) {
            // no match at all: the block may still be worth a compressed block with entropy-coded literals and zero sequences
            if (lit_bytes + 1 < bn) {
                if (lane == 0) body[lit_bytes] = 0;                // Number_of_Sequences = 0: the sequences section ends here
                payload = lit_bytes + 1;
            }
        } else {
            const uint32_t shdr = nseq < 128 ? 1u : (nseq < 0x7f00 ? 2u : 3u);
            uint8_t* sp = body + lit_bytes;
            uint32_t desc_bytes = 0;
            const uint32_t sbytes = zb_encode_sequences(seqs, nseq, (uint32_t*)buf, (ZbCTab*)ht, pk, sb, &fs, sp + shdr, &desc_bytes, lane);
            const uint32_t total = lit_bytes + shdr + desc_bytes + sbytes;
            if (total < bn) {
                if (lane == 0) {
                    if (shdr == 1) sp[0] = (uint8_t)nseq;
                    else if (shdr == 2) { sp[0] = (uint8_t)((nseq >> 8) + 0x80); sp[1] = (uint8_t)nseq; }
                    else { sp[0] = 0xff; sp[1] = (uint8_t)(nseq - 0x7f00); sp[2] = (uint8_t)((nseq - 0x7f00) >> 8); }
                }
                _Pragma("unroll 2")
                for (uint32_t i = lane; i < sbytes; i += 32) sp[shdr + desc_bytes + i] = buf[i];
                payload = total;
            }
        }
    }
    if (payload == 0xffffffffu) {                                  // Raw_Block
        ze_warp_copy(out + 3, src, bn, lane);
    }
    if (lane == 0) {
        const uint32_t type = payload == 0xffffffffu ? 0u : 2u;
        const uint32_t bsize = payload == 0xffffffffu ? bn : payload;
        const uint32_t hdr = (last_block ? 1u : 0u) | (type << 1) | (bsize << 3);
        out[0] = (uint8_t)hdr; out[1] = (uint8_t)(hdr >> 8); out[2] = (uint8_t)(hdr >> 16);
        A.blk_size[gblk] = 3 + bsize;
    }
    // ---- placement: the block claims its place in the frame by a decoupled look-back over the sizes of the blocks before
    // it (ze_lookback; they were launched earlier — CTAs start in grid order — so a wait is for work in flight, never for
    // work not yet scheduled) and copies itself there while it is still hot in L2: no assemble launch, no second pass.
    {
        const uint32_t bsize_all = 3 + (payload == 0xffffffffu ? bn : payload);
        unsigned long long* st = A.blk_state + (size_t)chunk * A.blocks_per_chunk;
        uint8_t* frame = A.out_base + A.out_off[chunk];
        if (blk == 0 && lane == 0) ze_frame_header(frame, clen);
        __threadfence();                                           // (the slot's bytes are this warp's own: ordered by the warp sync below)
        const uint32_t base = ze_lookback(st, blk, bsize_all, ze_frame_header_size(clen), last_block, lane);
        if (last_block && lane == 0) A.out_len[chunk] = base + bsize_all;
        __syncwarp();
        ze_warp_copy(frame + base, out, bsize_all, lane);
    }
}

// ------------------------------------------------------------------------------------------ host side
inline const char* zstd_blk_scratch_alloc(ZstdBlkScratch& s, uint32_t chunk_cap, uint32_t max_batch) {
    s.blocks_per_chunk = (chunk_cap + ZB - 1) / ZB;
    s.max_batch = max_batch;
    const size_t nblk = (size_t)s.blocks_per_chunk * max_batch;
    const char* e;
    if ((e = rt::malloc_device((void**)&s.blk_out, nblk * ZB_SLOT + 256))) return e;
    if ((e = rt::malloc_device((void**)&s.blk_size, nblk * 4 + 256))) return e;
    if ((e = rt::malloc_device((void**)&s.seqs, nblk * ZE_MAXSEQ * sizeof(uint2) + 256))) return e;
    if ((e = rt::malloc_device((void**)&s.lits, nblk * ZB + 256))) return e;
    if ((e = rt::malloc_device((void**)&s.blk_state, nblk * 8 + 256))) return e;
    return nullptr;
}
inline void zstd_blk_scratch_free(ZstdBlkScratch& s) {
    rt::free_device(s.blk_out); rt::free_device(s.blk_size); rt::free_device(s.seqs); rt::free_device(s.lits); rt::free_device(s.blk_state);
    s = ZstdBlkScratch{};
}
constexpr uint32_t ZB_SMEM_BYTES = ZB_WPB * ZB_SMEM_WARP_AL;

inline int zstd_compress_batch_blocks(ZstdBlkScratch& s, rt::stream_t st, const uint8_t* in_base, const uint64_t* d_in_off,
                                      const uint32_t* d_in_len, uint32_t n_chunks, uint32_t chunk_size,
                                      uint8_t* out_base, const uint64_t* d_out_off, uint32_t* d_out_len, LaunchProf& prof) {
    if (n_chunks > s.max_batch) { g_zstd_err = "batch larger than the context"; return -1; }
    const uint32_t bpc = (chunk_size + ZB - 1) / ZB;
    if (bpc > s.blocks_per_chunk) { g_zstd_err = "chunk larger than the context"; return -1; }
    ZstdBlkArgs A;
    A.in_base = in_base; A.in_off = d_in_off; A.in_len = d_in_len;
    A.blk_out = s.blk_out; A.blk_size = s.blk_size; A.seqs = s.seqs; A.lits = s.lits; A.blk_state = s.blk_state;
    A.blocks_per_chunk = s.blocks_per_chunk;
    A.out_base = out_base; A.out_off = d_out_off; A.out_len = d_out_len;
    const uint32_t gx = bpc ? (bpc + ZB_WPB - 1) / ZB_WPB : 1;   // (empty chunks still get their frame from block 0)
    const char* e = rt::memset_async(s.blk_state, 0, (size_t)s.blocks_per_chunk * n_chunks * 8, st);
    if (e) { g_zstd_err = e; return -7; }
    TS_LAUNCH_P(prof, "zstd_enc_blocks", zstd_enc_blocks_kernel, dim3(gx, n_chunks), dim3(ZB_WPB * 32), ZB_SMEM_BYTES, st, A);
    e = rt::last_error();
    if (e) { g_zstd_err = e; return -7; }
    return 0;
}

}  // namespace ts
