// zstd_huf_enc.cuh — literals section of a compressed block (RFC 8878 §3.1.1.3.1, §4.2): Raw, RLE, or Huffman-compressed with
// ONE tree per 64 KiB region: the first block of the region that uses it carries the tree description
// (Compressed_Literals_Block), the later ones are Treeless_Literals_Blocks.
//   ze_huf_build          one warp, once per region: region histogram -> rank sort of the used symbols -> two-queue Huffman
//                         merge on lane 0 (depth limited to 11 by halving counts and rebuilding) -> canonical codes in the
//                         order the decoder's table is filled (weight ascending, symbol ascending) -> tree description
//                         (direct 4-bit weights, or FSE-compressed weights when byte values above 128 occur)
//   ze_huf_encode_block   one warp per block: exact size first (sum of code lengths per stream), then the four streams
//                         packed by all lanes with a shuffle suffix-scan of code lengths and atomicOr into a zeroed staging
//                         buffer in shared memory
#pragma once
#include "ts_common.cuh"
#include "zstd_format.h"
#include "index_scan.cuh"
#include "zstd_fse_enc.cuh"

namespace ts {

constexpr uint32_t ZE_HUF_MIN = 64;            // below this many literals in a block Huffman streams cannot pay for their 6-byte jump table
constexpr uint32_t ZE_HUF_REGION_MIN = 256;    // below this many literals in a region a tree cannot pay for itself
// FSE-compressed weights shave ~1.5 % off text-like data at 8 KiB granularity; with one tree per 64 KiB region the tree is
// < 0.5 % of the output, so they are used only where direct weights cannot describe the tree (symbols > 128).
constexpr bool ZE_FSE_WEIGHTS_ALWAYS = false;

struct ZeHuf {                                 // region-wide result of ze_huf_build (shared memory)
    uint32_t ok;                               // 0: no usable tree (blocks write Raw / RLE literals)
    uint32_t tree_bytes;                       // size of the description in `desc`
    uint32_t max_len;
    uint8_t desc[132];                         // header byte + weights (direct: <= 1 + 64; FSE-compressed: <= 1 + 127)
};

// Region-wide tree.  hist[256] (counts, destroyed), ctab[256] out (code | len << 16, 0 = unused symbol), work: >= 7.5 KiB of
// shared-memory scratch, n = literals in the region.  One warp.
__device__ TS_NOINLINE void ze_huf_build(uint32_t* hist, uint32_t* ctab, uint32_t* work, uint32_t n, ZeHuf* hf, uint32_t lane) {
    uint32_t* keys = work + 1280;              // [256] used symbols: count << 8 | symbol
    uint32_t* sorted = work;                   // [256]
    uint32_t* nodew = work + 256;              // [512]
    uint16_t* parent = (uint16_t*)(work + 768);   // [512]
    uint8_t* depth = (uint8_t*)(work + 1024);  // [512]
    uint32_t* meta = work + 1200;              // small scalars shared by the warp
    if (lane == 0) { hf->ok = 0; hf->tree_bytes = 0; hf->max_len = 0; }
    for (uint32_t i = lane; i < 256; i += 32) ctab[i] = 0;
    __syncwarp();
    if (n < ZE_HUF_REGION_MIN) return;

    // compact the used symbols (8 per lane, ascending symbol order); counts are capped so that count << 8 cannot overflow
    uint32_t mine = 0;
    for (uint32_t k = 0; k < 8; k++) mine += hist[lane * 8 + k] ? 1u : 0u;
    const uint32_t inc = warp_inclusive_scan_u32(mine, lane);
    const uint32_t m = __shfl_sync(TS_FULL, inc, 31);
    {
        uint32_t at = inc - mine;
        for (uint32_t k = 0; k < 8; k++) { const uint32_t s = lane * 8 + k; if (hist[s]) keys[at++] = (hist[s] << 8) | s; }
    }
    __syncwarp();
    if (m < 2) return;                         // a single symbol: every block of the region is an RLE literals block
    {   // Shannon estimate in 1/16 bit units: sum c * (log2(n) - log2(c)), log2 by leading zeros + a linear fraction
        uint32_t cost = 0;
        const uint32_t hn = (uint32_t)zf::highbit32(n);
        const uint32_t lgn = (hn << 4) + (((n << (31 - hn)) >> 27) & 15) + 1;
        for (uint32_t k = 0; k < 8; k++) {
            const uint32_t c = hist[lane * 8 + k];
            if (c) {
                const uint32_t hb = (uint32_t)zf::highbit32(c);
                const uint32_t lg16 = (hb << 4) + (((c << (31 - hb)) >> 27) & 15);       // ~ 16 * log2(c)
                cost += c * (lgn - lg16);
            }
        }
        const uint32_t bits16 = __reduce_add_sync(TS_FULL, cost);
        if ((bits16 >> 7) + m / 2 + 16 >= n) return;                      // >= n bytes even before rounding losses
    }
    const uint32_t last_sym = keys[m - 1] & 0xff;

    // ---- code lengths: rebuild with halved counts until the tree is at most 11 deep
    uint32_t max_len = 0;
    for (uint32_t round = 0; round < 20; round++) {
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
    if (max_len > (uint32_t)zf::HUF_MAX_LOG) return;

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
        _Pragma("unroll 1")
        for (uint32_t s = 0; s <= last_sym; s++) {
            const uint32_t len = ctab[s] >> 16;
            if (!len) continue;
            const uint32_t w = max_len + 1 - len;
            ctab[s] = (start[w] >> (w - 1)) | (len << 16);
            start[w] += 1u << (w - 1);
        }
    }
    __syncwarp();
    const uint32_t nweights = last_sym;                                   // symbols 0 .. last_sym-1 are listed
    // Tree description: direct 4-bit weights (at most 128 of them) or FSE-compressed weights (RFC 8878 §4.2.1.1);
    // the shorter wins.  The FSE form is what lets alphabets above byte value 128 (binary payloads) be Huffman-coded.
    uint32_t tree_bytes = nweights <= 128 ? 1 + (nweights + 1) / 2 : 0xffffffffu;
    uint8_t* wdesc = hf->desc;
    bool fse_weights = false;
    {
        uint8_t* wts = (uint8_t*)(work + 1536);                           // weights, then FSE tables, in dead tree scratch
        uint32_t* wcnt = (uint32_t*)(wts + 256);                          // [16]
        uint16_t* wst = (uint16_t*)(wcnt + 16);                           // [64]
        zf::FseCSym* wsy = (zf::FseCSym*)(wst + 64);                      // [16]
        uint8_t* wscratch = (uint8_t*)(wsy + 16);                         // >= 784 bytes
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
                        while (nb >= 8 && ob < 120) { o[ob++] = (uint8_t)acc; acc >>= 8; nb -= 8; }
                    }
                    acc |= (uint64_t)(st[1] & ((1u << wk.log) - 1)) << nb; nb += wk.log;      // flush odd chain, then even chain
                    acc |= (uint64_t)(st[0] & ((1u << wk.log) - 1)) << nb; nb += wk.log;
                    acc |= 1ull << nb; nb += 1;                                                // end mark
                    while (nb > 0 && ob < 126) { o[ob++] = (uint8_t)acc; acc >>= 8; nb = nb >= 8 ? nb - 8 : 0; }
                    meta[2] = nb ? 0xffffu : dsz + ob;                    // did not fit 127 bytes: unusable
                }
                __syncwarp();
                const uint32_t fsz = meta[2];
                if (fsz < 128 && 1 + fsz < tree_bytes) {
                    if (lane == 0) wdesc[0] = (uint8_t)fsz;
                    tree_bytes = 1 + fsz;
                    fse_weights = true;
                }
            }
        }
        __syncwarp();
    }
    if (tree_bytes == 0xffffffffu) return;                                // neither form can describe this tree
    if (!fse_weights && lane == 0) {
        wdesc[0] = (uint8_t)(127 + nweights);
        for (uint32_t i = 0; i < nweights; i += 2) {
            const uint32_t l0 = ctab[i] >> 16, l1 = i + 1 < nweights ? ctab[i + 1] >> 16 : 0;
            const uint32_t w0 = l0 ? max_len + 1 - l0 : 0, w1 = l1 ? max_len + 1 - l1 : 0;
            wdesc[1 + i / 2] = (uint8_t)((w0 << 4) | w1);
        }
    }
    if (lane == 0) { hf->ok = 1; hf->tree_bytes = tree_bytes; hf->max_len = max_len; }
    __syncwarp();
}

struct ZeLitBlock {                  // what a block's literals section will consist of (decided before anything is written)
    uint32_t kind;                   // 0 Raw, 1 RLE, 2 Huffman (tree or treeless decided later)
    uint32_t n;                      // regenerated size
    uint32_t stream_bytes;           // Huffman: bytes of the four streams
    uint32_t ssz[3];                 // Huffman: sizes of streams 1-3 (jump table)
    uint32_t rle_byte;
};

// Exact sizes of the four Huffman streams of this block (no output).  Returns false when a literal has no code.
__device__ __forceinline__ void ze_huf_plan_block(const uint8_t* __restrict__ lits, uint32_t n, const uint32_t* ctab, const ZeHuf* hf,
                                                  ZeLitBlock* lb, uint32_t lane) {
    uint32_t kind = 0, stream_bytes = 0, s1 = 0, s2 = 0, s3 = 0, rle = 0;
    if (n >= 1) {                                                         // RLE: all literals equal
        const uint8_t first = lits[0];
        bool same = true;
        for (uint32_t i = lane; i < n; i += 32) same = same && lits[i] == first;
        if (__all_sync(TS_FULL, same)) { kind = 1; rle = first; }
    }
    if (kind == 0 && hf->ok && n >= ZE_HUF_MIN) {
        const uint32_t seg = (n + 3) / 4;
        uint32_t tot[4];
        bool coded = true;
        for (uint32_t st = 0; st < 4; st++) {
            const uint32_t a = st * seg, b = st < 3 ? min(n, a + seg) : n;
            uint32_t bits = 0;
            for (uint32_t i = a + lane; i < b; i += 32) { const uint32_t l = ctab[lits[i]] >> 16; coded = coded && l != 0; bits += l; }
            tot[st] = __reduce_add_sync(TS_FULL, bits);
        }
        if (__all_sync(TS_FULL, coded)) {
            s1 = (tot[0] + 8) >> 3; s2 = (tot[1] + 8) >> 3; s3 = (tot[2] + 8) >> 3;         // + end mark, rounded up
            const uint32_t s4 = (tot[3] + 8) >> 3;
            stream_bytes = s1 + s2 + s3 + s4;
            if (stream_bytes + 6 + 2 < n && s1 < 65536 && s2 < 65536 && s3 < 65536) kind = 2;   // must beat Raw (3-byte header vs up to 5)
        }
    }
    if (lane == 0) { lb->kind = kind; lb->n = n; lb->stream_bytes = stream_bytes; lb->ssz[0] = s1; lb->ssz[1] = s2; lb->ssz[2] = s3; lb->rle_byte = rle; }
    __syncwarp();
}

// The four streams of this block into `work` (shared-memory words, >= stream_bytes + 8 bytes).  Layout: streams back to back.
__device__ __forceinline__ void ze_huf_encode_block(const uint8_t* __restrict__ lits, uint32_t n, const uint32_t* ctab, uint32_t stream_bytes,
                                                    uint32_t* work, uint32_t lane) {
    for (uint32_t i = lane; i < (stream_bytes + 11) / 4; i += 32) work[i] = 0;
    __syncwarp();
    const uint32_t seg = (n + 3) / 4;
    uint32_t byte_pos = 0;
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
        byte_pos += (total + 1 + 7) >> 3;
        __syncwarp();
    }
}

// Literals_Section_Header for a Raw / RLE literals block (always the 3-byte form) — one lane.
__device__ __forceinline__ uint32_t ze_lit_header_raw(uint8_t* body, uint32_t type, uint32_t n) {
    body[0] = (uint8_t)(type | (3u << 2) | ((n & 0xf) << 4));
    body[1] = (uint8_t)(n >> 4);
    body[2] = (uint8_t)(n >> 12);
    return 3;
}
// ... and for a Compressed (type 2) / Treeless (type 3) block with four streams; comp = tree + jump table + streams.
__device__ __forceinline__ uint32_t ze_lit_header_huf(uint8_t* body, uint32_t type, uint32_t n, uint32_t comp) {
    const uint32_t hsz = (n < 1024 && comp < 1024) ? 3u : (n < 16384 && comp < 16384) ? 4u : 5u;
    const uint32_t sf = hsz - 2;                                          // 1, 2, 3: all with four streams
    uint64_t h;
    if (hsz == 3) h = type | (sf << 2) | ((uint64_t)n << 4) | ((uint64_t)comp << 14);
    else if (hsz == 4) h = type | (sf << 2) | ((uint64_t)n << 4) | ((uint64_t)comp << 18);
    else h = type | (sf << 2) | ((uint64_t)n << 4) | ((uint64_t)comp << 22);
    for (uint32_t k = 0; k < hsz; k++) body[k] = (uint8_t)(h >> (8 * k));
    return hsz;
}
__device__ __forceinline__ uint32_t ze_lit_header_huf_size(uint32_t n, uint32_t comp) {
    return (n < 1024 && comp < 1024) ? 3u : (n < 16384 && comp < 16384) ? 4u : 5u;
}

}  // namespace ts
