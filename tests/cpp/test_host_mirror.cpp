// Mirrors the reference's unit tests for the operator surface, against the C++ host mirror
// (tiered-storage-for-apache-kafka_b200/host/chunk_transform.hpp):
//   core/T/manifest/index/ChunkIndexBuilderCommonTest.java:37-127, FixedSizeChunkIndexBuilderTest.java:35-88,
//   VariableSizeChunkIndexBuilderTest.java:45-81, core/T/transform/BaseTransformChunkEnumerationTest.java:65-94,
//   core/T/transform/TransformFinisherTest.java:97-125, core/T/fetch/FetchChunkEnumerationTest.java:105-145,
//   core/T/transform/TransformsEndToEndTest.java:44-116 (round trip through the batched GPU chain, checked
//   against the oracle's libzstd/OpenSSL reader),
//   core/T/SegmentCompressionCheckerTest.java:43-130, core/T/transform/RateLimitedInputStreamTest.java:36-100 (with an
//   injected clock instead of wall-clock waits), RemoteStorageManager.requiresCompression (RemoteStorageManager.java:381-398).
// Links libtsgpu.so on a GPU box, or the test-only SIMT build on the CPU box (same sources, emulated kernels).
#include <cstdio>
#include <cstring>
#include <random>
#include <sstream>
#include "../../tiered-storage-for-apache-kafka_b200/host/chunk_transform.hpp"
#include "../../tiered-storage-for-apache-kafka_b200/host/segment_upload.hpp"
#include "../../oracle/tsoracle.h"

using namespace tieredstorage;
static int checks = 0, failures = 0;
#define CHECK(c) do { checks++; if (!(c)) { failures++; printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); } } while (0)
template <class E, class F> static bool throwsWith(F f, const char* msg) {
    try { f(); } catch (const E& e) { if (std::string(e.what()).find(msg) != std::string::npos) return true; printf("  got: %s\n", e.what()); return false; }
    catch (...) { return false; }
    return false;
}

static void builderTests() {
    for (int fixed = 0; fixed < 2; fixed++) {
        auto mk = [&](int ocs, int ofs) -> std::unique_ptr<AbstractChunkIndexBuilder> {
            if (fixed) return std::make_unique<FixedSizeChunkIndexBuilder>(ocs, ofs, 110);
            return std::make_unique<VariableSizeChunkIndexBuilder>(ocs, ofs);
        };
        CHECK(throwsWith<IllegalArgumentException>([&] { mk(-1, 250); }, "Original chunk size must be non-negative, -1 given"));
        CHECK(throwsWith<IllegalArgumentException>([&] { mk(100, -1); }, "Original file size must be non-negative, -1 given"));
        auto b = mk(100, 250);
        CHECK(throwsWith<IllegalArgumentException>([&] { b->addChunk(-1); }, "Transformed chunk size must be non-negative, -1 given"));
        b->addChunk(110); b->addChunk(110);
        CHECK(throwsWith<IllegalStateException>([&] { b->addChunk(110); }, "This must be final chunk. Call `finish` instead."));
        auto idx = b->finish(30);
        CHECK(throwsWith<IllegalStateException>([&] { b->addChunk(110); }, "Cannot add chunk to already finished index"));
        CHECK(throwsWith<IllegalStateException>([&] { b->finish(30); }, "Cannot finish already finished index"));
        CHECK(idx->chunks().size() == 3);
        CHECK((idx->chunks()[2] == Chunk{2, 200, 50, 220, 30}));
        for (int off = 0; off < 250; off++) CHECK(*idx->findChunkForOriginalOffset(off) == idx->chunks()[off / 100]);
        CHECK(!idx->findChunkForOriginalOffset(250).has_value());
        CHECK(throwsWith<IllegalArgumentException>([&] { idx->findChunkForOriginalOffset(-1); }, "Offset must be non-negative, -1 given"));
        auto b2 = mk(100, 250); b2->addChunk(110);
        CHECK(throwsWith<IllegalStateException>([&] { b2->finish(30); }, "This cannot be final chunk: not enough chunks to cover original file"));
        auto e = mk(100, 0)->finish(0);
        CHECK(e->chunks().size() == 1 && (e->chunks()[0] == Chunk{0, 0, 0, 0, 0}));
        CHECK(!e->findChunkForOriginalOffset(0).has_value());
    }
    FixedSizeChunkIndexBuilder fb(100, 250, 110);
    CHECK(throwsWith<IllegalArgumentException>([&] { fb.addChunk(109); }, "Non-final chunk must be of size 110, but 109 given"));
    // golden JSON (ChunkIndexSerializationTest.java:63-74)
    CHECK(FixedSizeChunkIndex(100, 250, 110, 30).toJson() ==
          "{\"type\":\"fixed\",\"originalChunkSize\":100,\"originalFileSize\":250,\"transformedChunkSize\":110,\"finalTransformedChunkSize\":30}");
    CHECK(VariableSizeChunkIndex(100, 250, {10, 20, 30}).toJson() ==
          "{\"type\":\"variable\",\"originalChunkSize\":100,\"originalFileSize\":250,\"transformedChunks\":\"KLUv/SAPeQAAAAAAAwAAAAoBAAoAAAAe\"}");
    // fetch plan (FetchChunkEnumerationTest.java:105-145)
    FixedSizeChunkIndex f(10, 30, 10, 10);
    auto p = fetchPlan(f, 2, 4);     CHECK(p.size() == 1 && p[0].chunkId == 0 && p[0].skip == 2 && p[0].take == 3);
    p = fetchPlan(f, 5, 24);         CHECK(p.size() == 3 && p[0].skip == 5 && p[0].take == 5 && p[1].take == 10 && p[2].take == 5);
    p = fetchPlan(f, 25, 1000);      CHECK(p.size() == 1 && p[0].chunkId == 2 && p[0].skip == 5 && p[0].take == 5);
    CHECK(throwsWith<IllegalArgumentException>([&] { fetchPlan(f, 30, 31); }, "Invalid start position 30"));
}

static void manifestTests() {
    // golden strings: core/T/manifest/SegmentManifestV1SerdeTest.java:82-133
    const std::string META = "{\"remoteLogSegmentId\":{\"topicIdPartition\":{\"topicId\":\"lZ6vvmajTWKDBUTV6SQAtQ\",\"topicPartition\":"
        "{\"topic\":\"topic1\",\"partition\":42}},\"id\":\"adh9f8BMS4anaUnD8KWfWg\"},\"startOffset\":0,\"endOffset\":1000,"
        "\"maxTimestampMs\":1000000000,\"brokerId\":2,\"eventTimestampMs\":2000000000,\"segmentLeaderEpochs\":{\"0\":100,\"1\":200,\"2\":300}}";
    const std::string HEAD = "{\"version\":\"1\",\"chunkIndex\":{\"type\":\"fixed\",\"originalChunkSize\":100,\"originalFileSize\":1000,"
        "\"transformedChunkSize\":110,\"finalTransformedChunkSize\":110},\"segmentIndexes\":{\"offset\":{\"position\":0,\"size\":1},"
        "\"timestamp\":{\"position\":1,\"size\":1},\"producerSnapshot\":{\"position\":2,\"size\":1},\"leaderEpoch\":{\"position\":3,\"size\":1},";
    RemoteLogSegmentMetadataJson m{"lZ6vvmajTWKDBUTV6SQAtQ", "topic1", 42, "adh9f8BMS4anaUnD8KWfWg", 0, 1000, 1000000000LL, 2, 2000000000LL,
                                   {{0, 100}, {1, 200}, {2, 300}}};
    FixedSizeChunkIndex idx(100, 1000, 110, 110);
    SegmentIndexesV1 with{{0, 1}, {1, 1}, {2, 1}, {3, 1}, SegmentIndexV1{4, 1}}, without{{0, 1}, {1, 1}, {2, 1}, {3, 1}, std::nullopt};
    Bytes aad{10, 11, 12, 13};
    CHECK(segmentManifestV1Json(idx, with, false, std::nullopt, nullptr, m) ==
          HEAD + "\"transaction\":{\"position\":4,\"size\":1}},\"compression\":false,\"remoteLogSegmentMetadata\":" + META + "}");
    CHECK(segmentManifestV1Json(idx, without, false, std::nullopt, nullptr, m) ==
          HEAD + "\"transaction\":null},\"compression\":false,\"remoteLogSegmentMetadata\":" + META + "}");
    CHECK(segmentManifestV1Json(idx, with, false, std::nullopt, &aad, m) ==
          HEAD + "\"transaction\":{\"position\":4,\"size\":1}},\"compression\":false,\"encryption\":{\"aad\":\"CgsMDQ==\"},\"remoteLogSegmentMetadata\":" + META + "}");
    CHECK(segmentManifestV1Json(idx, with, true, std::string("key1:AAEC"), &aad, m).find("\"compression\":true,\"encryption\":{\"dataKey\":\"key1:AAEC\",\"aad\":\"CgsMDQ==\"}") != std::string::npos);
    // reader: the golden strings parse back to what they were written from, and re-serialise to themselves
    for (int v = 0; v < 4; v++) {
        const std::string text = v == 0 ? segmentManifestV1Json(idx, with, false, std::nullopt, nullptr, m)
                               : v == 1 ? segmentManifestV1Json(idx, without, false, std::nullopt, nullptr, m)
                               : v == 2 ? segmentManifestV1Json(idx, with, false, std::nullopt, &aad, m)
                                        : segmentManifestV1Json(idx, with, true, std::string("key1:AAEC"), &aad, m);
        SegmentManifestV1 r = parseSegmentManifestV1(text, nullptr);
        CHECK(r.chunkIndex->toJson() == idx.toJson() && r.compression == (v == 3));
        CHECK(r.segmentIndexes.transaction.has_value() == (v != 1) && r.segmentIndexes.leaderEpoch.position == 3);
        CHECK(r.aad.has_value() == (v >= 2) && (!r.aad || *r.aad == aad));
        CHECK(r.wrappedDataKey.has_value() == (v == 3) && (!r.wrappedDataKey || *r.wrappedDataKey == "key1:AAEC"));
        CHECK(r.remoteLogSegmentMetadata.topic == "topic1" && r.remoteLogSegmentMetadata.partition == 42 && r.remoteLogSegmentMetadata.segmentLeaderEpochs.size() == 3);
        CHECK(segmentManifestV1Json(*r.chunkIndex, r.segmentIndexes, r.compression, r.wrappedDataKey, r.aad ? &*r.aad : nullptr, r.remoteLogSegmentMetadata) == text);
    }
    CHECK(throwsWith<IllegalArgumentException>([&] { parseSegmentManifestV1("{\"version\":\"2\"}", nullptr); }, "unsupported version"));
    CHECK(throwsWith<IllegalArgumentException>([&] { parseSegmentManifestV1("{\"version\":\"1\",\"extra\":1}", nullptr); }, "unknown property 'extra'"));
    CHECK(throwsWith<IllegalArgumentException>([&] { parseSegmentManifestV1("{\"version\":\"1\"", nullptr); }, "manifest JSON"));
    CHECK(base64StdDecode("CgsMDQ==") == aad);
    // core/IT/RemoteStorageManagerTest.java:117-120 style key
    CHECK(objectKey("test/", m, Suffix::LOG) == "test/topic1-lZ6vvmajTWKDBUTV6SQAtQ/42/00000000000000000000-adh9f8BMS4anaUnD8KWfWg.log");
    CHECK(objectKey("", m, Suffix::MANIFEST) == "topic1-lZ6vvmajTWKDBUTV6SQAtQ/42/00000000000000000000-adh9f8BMS4anaUnD8KWfWg.rsm-manifest");
}

static void baseAndFinisherTests() {
    std::istringstream in(std::string("0123456789"));
    BaseTransformChunkEnumeration base(&in, 3);                       // BaseTransformChunkEnumerationTest: "012" "345" "678" "9"
    CHECK(base.originalChunkSize() == 3 && *base.transformedChunkSize() == 3);
    std::vector<std::string> got;
    while (base.hasMoreElements()) { Bytes c = base.nextElement(); got.emplace_back(c.begin(), c.end()); }
    CHECK((got == std::vector<std::string>{"012", "345", "678", "9"}));
    CHECK(throwsWith<NoSuchElementException>([&] { base.nextElement(); }, "NoSuchElement"));
    CHECK(throwsWith<IllegalArgumentException>([&] { BaseTransformChunkEnumeration(&in, -1); }, "originalChunkSize must be non-negative, -1 given"));
    CHECK(throwsWith<NullPointerException>([&] { BaseTransformChunkEnumeration(nullptr, 1); }, "inputStream cannot be null"));
    // TransformFinisherTest.java:97-125: 7 bytes, chunk 3 -> Chunk(0,0,3,0,3),(1,3,3,3,3),(2,6,1,6,1); fixed index for the base transform
    std::istringstream in2(std::string("\0\1\2\3\4\5\6", 7));
    BaseTransformChunkEnumeration b2(&in2, 3);
    TransformFinisher fin(&b2, 7);
    Bytes all = fin.readAll();
    CHECK(all.size() == 7);
    auto idx = fin.chunkIndex();
    CHECK(dynamic_cast<FixedSizeChunkIndex*>(idx.get()) != nullptr);
    CHECK(idx->chunks().size() == 3 && (idx->chunks()[1] == Chunk{1, 3, 3, 3, 3}) && (idx->chunks()[2] == Chunk{2, 6, 1, 6, 1}));
    // not consumed + base transform => index computed from the file size (TransformFinisher.java:124-132)
    std::istringstream in3(std::string(250, 'x'));
    BaseTransformChunkEnumeration b3(&in3, 100);
    CHECK(TransformFinisher(&b3, 250).chunkIndex()->chunks().size() == 3);
}

static void gpuChainTests(tsgpu_ctx* ctx) {
    std::mt19937 rng(7);
    const int n = 181200;
    Bytes src(n);
    for (int i = 0; i < n; i++) src[i] = (i < n / 2) ? (uint8_t)("the quick brown fox jumps over the lazy dog "[(i * 7 + (i >> 9)) % 44]) : (uint8_t)rng();
    DataKeyAndAAD km{Bytes(32), Bytes(32)};
    for (auto& b : km.dataKey) b = (uint8_t)rng();
    for (auto& b : km.aad) b = (uint8_t)rng();
    uint32_t ivc = 0;
    IvSupplier ivs = [&](uint8_t* iv) { memset(iv, 0, 12); memcpy(iv, &ivc, 4); ivc++; };
    for (int cs : {0, 13, 1024, 5123, n - 1, 2 * n}) {
        for (int mode = 1; mode < 4; mode++) {                         // 1 = zstd, 2 = aes, 3 = both (TransformsEndToEndTest grid)
            const int use_n = cs == 13 ? 1300 : n;
            std::istringstream in(std::string((const char*)src.data(), use_n));
            BaseTransformChunkEnumeration base(&in, cs);
            std::unique_ptr<CompressionChunkEnumeration> comp;
            std::unique_ptr<EncryptionChunkEnumeration> enc;
            TransformChunkEnumeration* top = &base;
            if (mode & 1) { comp = std::make_unique<CompressionChunkEnumeration>(ctx, top, 8); top = comp.get(); }
            if (mode & 2) { enc = std::make_unique<EncryptionChunkEnumeration>(ctx, top, km, ivs, 8); top = enc.get(); }
            // transformedChunkSize(): fixed only without compression (EncryptionChunkEnumeration.java:35-47)
            if (mode == 2 && cs) CHECK(*top->transformedChunkSize() == cs + 28); else if (mode & 1) CHECK(!top->transformedChunkSize().has_value());
            TransformFinisher fin(top, use_n, cs != 0);
            Bytes obj = fin.readAll();
            auto idx = fin.chunkIndex();
            CHECK((mode & 1) ? dynamic_cast<VariableSizeChunkIndex*>(idx.get()) != nullptr : dynamic_cast<FixedSizeChunkIndex*>(idx.get()) != nullptr);
            const auto& chunks = idx->chunks();
            CHECK((size_t)(chunks.back().transformedPosition + chunks.back().transformedSize) == obj.size());
            // the reference-side reader (oracle: libzstd + OpenSSL) recovers the bytes chunk by chunk
            std::vector<uint32_t> ts; for (auto& c : chunks) ts.push_back((uint32_t)c.transformedSize);
            Bytes back(use_n + 64); std::vector<uint32_t> osz(ts.size());
            int rc = ora_detransform_chunks(mode, obj.data(), ts.data(), (uint32_t)ts.size(), km.dataKey.data(), km.aad.data(), 32, back.data(), back.size(), osz.data());
            CHECK(rc == 0);
            CHECK(memcmp(back.data(), src.data(), use_n) == 0);
            // and our batched detransform enumeration does too
            std::istringstream oin(std::string((const char*)obj.data(), obj.size()));
            DetransformChunkEnumeration de(ctx, &oin, chunks, mode & 1, (mode & 2) ? &km : nullptr, cs ? (uint32_t)std::min(cs, use_n) : use_n, 8);
            Bytes round;
            while (de.hasMoreElements()) { Bytes c = de.nextElement(); round.insert(round.end(), c.begin(), c.end()); }
            CHECK(round.size() == (size_t)use_n && memcmp(round.data(), src.data(), use_n) == 0);
            {   // the manifest a segment like this one would get: written, read back (variable indexes decode their size list on the device)
                RemoteLogSegmentMetadataJson md{"lZ6vvmajTWKDBUTV6SQAtQ", "topic1", 7, "adh9f8BMS4anaUnD8KWfWg", 0, 99, 1, 2, 3, {{0, 0}}};
                SegmentIndexesV1 six{{0, 10}, {10, 10}, {20, 10}, {30, 10}, std::nullopt};
                const std::string text = segmentManifestV1Json(*idx, six, mode & 1, (mode & 2) ? std::optional<std::string>("k:AAEC") : std::nullopt,
                                                               (mode & 2) ? &km.aad : nullptr, md, ctx);   // transformedChunks compressed on the device
                CHECK(text.size() <= segmentManifestV1Json(*idx, six, mode & 1, (mode & 2) ? std::optional<std::string>("k:AAEC") : std::nullopt,
                                                           (mode & 2) ? &km.aad : nullptr, md).size());
                SegmentManifestV1 back = parseSegmentManifestV1(text, ctx);
                CHECK(back.chunkIndex->chunks().size() == idx->chunks().size());
                bool same = true;
                for (size_t k = 0; k < idx->chunks().size(); k++)
                    same = same && back.chunkIndex->chunks()[k].transformedPosition == idx->chunks()[k].transformedPosition &&
                           back.chunkIndex->chunks()[k].transformedSize == idx->chunks()[k].transformedSize;
                CHECK(same && back.compression == ((mode & 1) != 0) && back.aad.has_value() == ((mode & 2) != 0));
            }
            {   // FetchChunkEnumerationTest.java:105-145 on real transformed data: one ranged GET, batched detransform
                int gets = 0;
                auto fetcher = [&](const BytesRange& r) { gets++; return Bytes(obj.begin() + r.from, obj.begin() + r.to + 1); };
                const int last = use_n - 1;
                const int ranges[][2] = {{0, last}, {last, last}, {0, 0}, {use_n / 3, use_n / 3 + 2}, {use_n / 2, last + 1000}, {1, std::max(1, last - 1)}};
                for (auto& rg : ranges) {
                    if (rg[0] > last) continue;
                    gets = 0;
                    FetchChunkEnumeration fe(ctx, *idx, BytesRange::of(rg[0], rg[1]), fetcher, mode & 1, (mode & 2) ? &km : nullptr, 4);
                    Bytes got = fe.readAll();
                    const int to = std::min(rg[1], last);
                    CHECK(gets == 1);
                    CHECK(got.size() == (size_t)(to - rg[0] + 1) && memcmp(got.data(), src.data() + rg[0], got.size()) == 0);
                }
                CHECK(throwsWith<IllegalArgumentException>([&] { FetchChunkEnumeration(ctx, *idx, BytesRange::of(use_n, use_n + 5), fetcher, mode & 1, (mode & 2) ? &km : nullptr); },
                                                           "Invalid start position"));
                CHECK(throwsWith<IllegalArgumentException>([&] { FetchChunkEnumeration(ctx, *idx, BytesRange::empty(0), fetcher, mode & 1, (mode & 2) ? &km : nullptr); },
                                                           "range cannot be empty"));
            }
            {   // DetransformFinisher: the same bytes as one stream
                std::istringstream fin2(std::string((const char*)obj.data(), obj.size()));
                DetransformChunkEnumeration de2(ctx, &fin2, chunks, mode & 1, (mode & 2) ? &km : nullptr, cs ? (uint32_t)std::min(cs, use_n) : use_n, 3);
                DetransformFinisher df(&de2);
                Bytes all = df.readAll();
                CHECK(all.size() == (size_t)use_n && memcmp(all.data(), src.data(), use_n) == 0);
                CHECK(!df.hasMoreElements());
                CHECK(throwsWith<NoSuchElementException>([&] { df.nextElement(); }, "NoSuchElementException"));
            }
            // truncated stream: "Stream has fewer bytes than expected" (BaseDetransformChunkEnumerationTest.java:100-117)
            std::istringstream tin(std::string((const char*)obj.data(), obj.size() - 1));
            DetransformChunkEnumeration dt(ctx, &tin, chunks, mode & 1, (mode & 2) ? &km : nullptr, cs ? (uint32_t)std::min(cs, use_n) : use_n, 1024);
            CHECK(throwsWith<std::runtime_error>([&] { while (dt.hasMoreElements()) dt.nextElement(); }, "Stream has fewer bytes than expected"));
        }
    }
}


// ---- record batches as kafka-clients 3.6.0 writes them (MemoryRecordsBuilder), built by hand from the published format
static void put32(Bytes& b, size_t at, uint32_t v) { b[at] = v >> 24; b[at + 1] = v >> 16; b[at + 2] = v >> 8; b[at + 3] = v; }
static Bytes batchV2(int codec, const std::string& payload, int64_t baseOffset = 0) {
    Bytes b(61 + payload.size(), 0);
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(baseOffset >> (56 - 8 * i));
    put32(b, 8, (uint32_t)(b.size() - 12));                  // batchLength
    b[16] = 2;                                               // magic
    b[21] = 0; b[22] = (uint8_t)codec;                       // attributes
    put32(b, 57, 1);                                         // recordsCount
    memcpy(b.data() + 61, payload.data(), payload.size());
    put32(b, 17, crc32c(b.data() + 21, b.size() - 21));
    return b;
}
static Bytes recordV1(int codec, const std::string& payload) {
    Bytes b(12 + 22 + payload.size(), 0);                    // offset, size | crc, magic, attributes, timestamp, keyLen(-1), valueLen
    put32(b, 8, (uint32_t)(b.size() - 12));
    b[16] = 1; b[17] = (uint8_t)codec;
    put32(b, 26, 0xffffffffu);                               // null key
    put32(b, 30, (uint32_t)payload.size());
    memcpy(b.data() + 34, payload.data(), payload.size());
    put32(b, 12, crc32(b.data() + 16, b.size() - 16));
    return b;
}

static void bytesRangeTests() {                              // storage/core/T/.../BytesRangeTest.java
    CHECK(throwsWith<IllegalArgumentException>([] { BytesRange::of(-1, 1); }, "from cannot be negative, -1 given"));
    CHECK(throwsWith<IllegalArgumentException>([] { BytesRange::of(2, 1); }, "to cannot be less than from, from=2, to=1 given"));
    CHECK(BytesRange::of(1, 1).size() == 1 && BytesRange::of(0, 9).size() == 10 && BytesRange::of(3, 9).lastPosition() == 9);
    CHECK(BytesRange::empty(5).isEmpty() && BytesRange::empty(5).size() == 0 && BytesRange::ofFromPositionAndSize(5, 0) == BytesRange::empty(5));
    CHECK(BytesRange::ofFromPositionAndSize(1, 2) == BytesRange::of(1, 2));
    CHECK(throwsWith<IllegalStateException>([] { BytesRange::empty(0).lastPosition(); }, "No last position, range is empty"));
    CHECK(BytesRange::of(10, 19).toString() == "BytesRange{position=10, size=10}");
    auto idx = VariableSizeChunkIndex(100, 250, {10, 20, 30});
    auto r = transformedRange(idx.chunksForRange(BytesRange::of(100, 249).from, BytesRange::of(100, 249).to));
    CHECK(r == BytesRange::of(10, 59));                      // chunks 1 and 2: transformed bytes [10, 60)
    // the reference steps by the chunk's size from the RANGE start (AbstractChunkIndex.java:113-123), so an unaligned
    // start can stop before the last chunk the range touches; the mirror keeps that behaviour
    CHECK(idx.chunksForRange(150, 249).size() == 1 && idx.chunksForRange(150, 250).size() == 1 && idx.chunksForRange(100, 200).size() == 2);
}

static void uploadSideTests() {
    // checksum known answers (the "check" values of the CRC catalogue): pins both polynomials and the reflection
    CHECK(crc32c((const uint8_t*)"123456789", 9) == 0xE3069283u);
    CHECK(crc32((const uint8_t*)"123456789", 9) == 0xCBF43926u);
    // SegmentCompressionCheckerTest.shouldFailWhenReadingEmptyFile
    CHECK(throwsWith<InvalidRecordBatchException>([] { SegmentCompressionChecker::check(nullptr, 0); }, "Record batch is null"));
    // shouldReturnCompressedWhenEnabled: NONE -> false, ZSTD -> true (and the other codecs)
    for (int codec = 0; codec <= 4; codec++) {
        Bytes b = batchV2(codec, "key-0value-0");
        CHECK(SegmentCompressionChecker::check(b.data(), b.size()) == (codec != 0));
        Bytes l = recordV1(codec & 3, "value-0");
        CHECK(SegmentCompressionChecker::check(l.data(), l.size()) == ((codec & 3) != 0));
    }
    // shouldReturnCompressedWhenCompressionChanges: only the FIRST batch decides
    for (int first = 0; first < 2; first++) {
        Bytes a = batchV2(first ? 4 : 0, "key-0value-0"), b2 = batchV2(first ? 0 : 4, "key-1value-1", 1);
        a.insert(a.end(), b2.begin(), b2.end());
        CHECK(SegmentCompressionChecker::check(a.data(), a.size()) == (first == 1));
    }
    // shouldFailWhenReadingInvalidFile: a damaged record body fails the checksum
    {
        Bytes b = batchV2(0, "key-0value-0");
        for (auto& c : b) if (c == '0') c = '1';
        CHECK(throwsWith<InvalidRecordBatchException>([&] { SegmentCompressionChecker::check(b.data(), b.size()); }, "Failed to read and validate first batch"));
        Bytes t = batchV2(4, "key-0value-0"); t.pop_back();  // a partial batch is "no batch" (FileLogInputStream.nextBatch)
        CHECK(throwsWith<InvalidRecordBatchException>([&] { SegmentCompressionChecker::check(t.data(), t.size()); }, "Record batch is null"));
        Bytes m = batchV2(0, "x"); m[16] = 9;
        CHECK(throwsWith<InvalidRecordBatchException>([&] { SegmentCompressionChecker::check(m.data(), m.size()); }, "Failed to read and validate first batch"));
    }
    // RemoteStorageManager.requiresCompression
    {
        Bytes plain = batchV2(0, "v"), z = batchV2(4, "v"), bad = batchV2(0, "v"); bad[30] ^= 1;
        CHECK(!requiresCompression(false, false, plain.data(), plain.size()));
        CHECK(requiresCompression(true, false, z.data(), z.size()));            // heuristic off: always compress
        CHECK(requiresCompression(true, true, plain.data(), plain.size()));
        CHECK(!requiresCompression(true, true, z.data(), z.size()));
        CHECK(!requiresCompression(true, true, bad.data(), bad.size()));        // unreadable first batch: upload uncompressed
    }
    // RateLimitedInputStreamTest with an injected clock: sleeping advances it
    {
        int64_t now = 0, slept = 0;
        auto clock = [&] { return now; };
        auto sleep = [&](int64_t ns) { slept += ns; now += ns; };
        RateLimitBucket bucket(1, clock, sleep);
        CHECK(bucket.capacity() == RateLimitBucket::MIN_RATE);                  // rateLimitBucket: max(uploadRate, MIN_RATE)
        CHECK(bucket.consume(RateLimitBucket::MIN_RATE - 1) == 0);              // testDoesNotBlockRead
        CHECK(bucket.consume(0) == 0);
        const int64_t w = bucket.consume(RateLimitBucket::MIN_RATE - 1);        // testBlocksOnSeparateStreams: ~1 s for the refill
        CHECK(w > 990000000 && w <= 1000000000);
        RateLimitBucket b2(1, clock, sleep);
        const int64_t w2 = b2.consume(RateLimitBucket::MIN_RATE + 1);           // testBlocksRead: one token short of the capacity
        CHECK(w2 > 0 && w2 < 1000000);
        b2.forceAddTokens(100);                                                 // tokens handed back for bytes not read
        CHECK(b2.availableTokens() > 98.0 && b2.availableTokens() < 101.0);
        now += 5000000000ll;                                                    // refill stops at the capacity
        CHECK(b2.availableTokens() == (double)RateLimitBucket::MIN_RATE);
        b2.forceAddTokens(7);                                                   // ... but forced tokens may exceed it
        CHECK(b2.availableTokens() == (double)RateLimitBucket::MIN_RATE + 7);
    }
    // S3MultiPartOutputStream part plan
    {
        auto parts = multipartPlan(12, 5);
        CHECK(parts.size() == 3 && parts[0].partNumber == 1 && parts[2].offset == 10 && parts[2].size == 2);
        CHECK(multipartPlan(0, 5).empty() && multipartPlan(10, 5).size() == 2);
        CHECK(throwsWith<IllegalArgumentException>([] { multipartPlan(1, 0); }, "partSize must be positive"));
    }
}

// RemoteStorageManager.uploadSegmentLog through the GPU chain: heuristic decides the flags, parts arrive in order, the
// object equals what the reference-side reader expects, the chunk index is the reference's
// RemoteStorageManager.uploadIndexes / fetchIndexBytes: five index files, one ragged AES batch, positions as the reference records them
static void indexFileTests(tsgpu_ctx* ctx) {
    // SegmentIndexesV1BuilderTest-style state machine
    CHECK(throwsWith<IllegalStateException>([] { SegmentIndexesV1Builder().add(IndexType::OFFSET, 1).add(IndexType::OFFSET, 1); }, "Index OFFSET is already added"));
    CHECK(throwsWith<IllegalStateException>([] { SegmentIndexesV1Builder().add(IndexType::OFFSET, 1).add(IndexType::TIMESTAMP, 1).build(); },
                                            "Not enough indexes have been added; at least 4 required. Indexes included: [OFFSET, TIMESTAMP]"));
    CHECK(throwsWith<IllegalStateException>([] { SegmentIndexesV1Builder().add(IndexType::OFFSET, 1).add(IndexType::TIMESTAMP, 1).add(IndexType::PRODUCER_SNAPSHOT, 1)
                                                     .add(IndexType::TRANSACTION, 1).build(); }, "OFFSET, TIMESTAMP, PRODUCER_SNAPSHOT, and LEADER_EPOCH indexes are required"));
    SegmentIndexesV1 plain = SegmentIndexesV1Builder().add(IndexType::OFFSET, 10).add(IndexType::TIMESTAMP, 20).add(IndexType::PRODUCER_SNAPSHOT, 0)
                                 .add(IndexType::LEADER_EPOCH, 5).build();
    CHECK(plain.timestamp.position == 10 && plain.producerSnapshot.position == 30 && plain.leaderEpoch.position == 30 && !plain.transaction);
    std::mt19937 rng(11);
    auto blob = [&](size_t n) { Bytes b(n); for (auto& x : b) x = (uint8_t)rng(); return b; };
    DataKeyAndAAD km; km.dataKey = blob(32); km.aad = blob(32);
    for (int enc = 0; enc < 2; enc++) for (int txn = 0; txn < 2; txn++) {
        std::vector<std::pair<IndexType, Bytes>> blobs = {{IndexType::OFFSET, blob(10485)}, {IndexType::TIMESTAMP, blob(9000)}, {IndexType::PRODUCER_SNAPSHOT, blob(0)},
                                                          {IndexType::LEADER_EPOCH, blob(126)}};
        if (txn) blobs.push_back({IndexType::TRANSACTION, blob(2048)});
        Bytes ivs = blob(12 * 4);
        SegmentIndexesUpload up = uploadIndexes(ctx, blobs, enc != 0, enc ? &km : nullptr, enc ? ivs.data() : nullptr);
        const int ov = enc ? 28 : 0;
        CHECK(up.segmentIndexes.offset.position == 0 && up.segmentIndexes.offset.size == 10485 + ov);
        CHECK(up.segmentIndexes.timestamp.position == 10485 + ov && up.segmentIndexes.producerSnapshot.size == 0);
        CHECK(up.segmentIndexes.leaderEpoch.position == 10485 + 9000 + 2 * ov && up.segmentIndexes.transaction.has_value() == (txn != 0));
        CHECK(up.object.size() == (size_t)(10485 + 9000 + 126 + (txn ? 2048 : 0) + ov * (3 + txn)));
        if (enc) {                                           // bit-exact with the reference-side cipher, blob by blob
            Bytes want(10485 + 28);
            CHECK(ora_aesgcm_encrypt_chunk(km.dataKey.data(), ivs.data(), km.aad.data(), 32, blobs[0].second.data(), 10485, want.data()) == 0);
            CHECK(memcmp(want.data(), up.object.data(), want.size()) == 0);
        }
        const SegmentIndexV1 all[] = {up.segmentIndexes.offset, up.segmentIndexes.timestamp, up.segmentIndexes.producerSnapshot, up.segmentIndexes.leaderEpoch};
        for (int k = 0; k < 4; k++) CHECK(fetchIndexBytes(ctx, up.object, all[k], enc ? &km : nullptr) == blobs[k].second);
        if (txn) CHECK(fetchIndexBytes(ctx, up.object, *up.segmentIndexes.transaction, enc ? &km : nullptr) == blobs[4].second);
    }
}

// A manifest as the REFERENCE writes it: the size list compressed by libzstd (oracle), read back through the device decoder
static void referenceManifestTests(tsgpu_ctx* ctx) {
    std::vector<int32_t> sizes;
    for (int i = 0; i < 600; i++) sizes.push_back(1000000 + (i * 7919) % 5000);
    std::vector<char> b64(64 * 1024);
    const int64_t n = ora_transformed_chunks_serialize(sizes.data(), (int32_t)sizes.size(), b64.data(), b64.size());
    CHECK(n > 0);
    const std::string text = std::string("{\"version\":\"1\",\"chunkIndex\":{\"type\":\"variable\",\"originalChunkSize\":1048576,\"originalFileSize\":")
        + std::to_string(599LL * 1048576 + 1) + ",\"transformedChunks\":\"" + std::string(b64.data(), (size_t)n) + "\"},\"segmentIndexes\":{\"offset\":{\"position\":0,\"size\":1},"
        "\"timestamp\":{\"position\":1,\"size\":1},\"producerSnapshot\":{\"position\":2,\"size\":1},\"leaderEpoch\":{\"position\":3,\"size\":1},\"transaction\":null},"
        "\"compression\":true,\"encryption\":{\"dataKey\":\"k1:AAEC\",\"aad\":\"CgsMDQ==\"}}";
    SegmentManifestV1 m = parseSegmentManifestV1(text, ctx);
    auto* v = dynamic_cast<VariableSizeChunkIndex*>(m.chunkIndex.get());
    CHECK(v != nullptr && v->transformedChunks == sizes && m.compression && m.wrappedDataKey && *m.wrappedDataKey == "k1:AAEC");
    CHECK(m.chunkIndex->chunks().size() == 600 && m.chunkIndex->chunks()[599].originalSize == 1);
    // the golden variable index of ChunkIndexSerializationTest.java:63-74
    SegmentManifestV1 g = parseSegmentManifestV1("{\"version\":\"1\",\"chunkIndex\":{\"type\":\"variable\",\"originalChunkSize\":100,\"originalFileSize\":250,"
        "\"transformedChunks\":\"KLUv/SAPeQAAAAAAAwAAAAoBAAoAAAAe\"},\"segmentIndexes\":{\"offset\":{\"position\":0,\"size\":1},\"timestamp\":{\"position\":1,\"size\":1},"
        "\"producerSnapshot\":{\"position\":2,\"size\":1},\"leaderEpoch\":{\"position\":3,\"size\":1},\"transaction\":null},\"compression\":false}", ctx);
    CHECK(g.chunkIndex->chunks().size() == 3 && g.chunkIndex->chunks()[2].transformedPosition == 30 && g.chunkIndex->chunks()[2].transformedSize == 30);
}

static void segmentUploadTests(tsgpu_ctx* ctx) {
    std::mt19937 rng(5);
    for (int codec : {0, 4}) for (int enc = 0; enc < 2; enc++) {
        std::string payload;
        for (int i = 0; i < 40000; i++) payload += codec ? (char)rng() : "topic-partition-offset "[i % 23];
        Bytes seg = batchV2(codec, payload);
        Bytes more = batchV2(codec, payload, 1); seg.insert(seg.end(), more.begin(), more.end());
        const int cs = 30000;
        const int n = (int)seg.size(), nch = (n + cs - 1) / cs;
        DataKeyAndAAD km; km.dataKey.resize(32); km.aad.resize(32);
        for (auto& b : km.dataKey) b = (uint8_t)rng();
        for (auto& b : km.aad) b = (uint8_t)rng();
        Bytes ivs(12 * nch); for (auto& b : ivs) b = (uint8_t)rng();
        int64_t now = 0, slept = 0;
        RateLimitBucket bucket(1 << 20, [&] { return now; }, [&](int64_t ns) { slept += ns; now += ns; });
        SegmentLogUploader up(ctx, cs, /*compressionEnabled=*/true, /*heuristic=*/true, enc != 0, /*partSize=*/16384, &bucket);
        Bytes object; int lastPart = 0; bool ordered = true;
        auto res = up.upload(seg.data(), (uint64_t)n, enc ? &km : nullptr, enc ? ivs.data() : nullptr, [&](const UploadPart& p, const uint8_t* d) {
            ordered = ordered && p.partNumber == lastPart + 1 && p.offset == object.size(); lastPart = p.partNumber;
            object.insert(object.end(), d, d + p.size);
        });
        CHECK(ordered && object.size() == res.objectBytes);
        CHECK(res.compressed == (codec == 0));                                  // already-compressed batches skip zstd
        CHECK((dynamic_cast<VariableSizeChunkIndex*>(res.chunkIndex.get()) != nullptr) == res.compressed);
        const uint32_t flags = (res.compressed ? 1u : 0u) | (enc ? 2u : 0u);
        std::vector<uint32_t> ts; for (auto& c : res.chunkIndex->chunks()) ts.push_back((uint32_t)c.transformedSize);
        CHECK((int)ts.size() == nch);
        Bytes back(n + 64); std::vector<uint32_t> osz(ts.size());
        if (flags) {
            int rc = ora_detransform_chunks(flags, object.data(), ts.data(), (uint32_t)ts.size(), km.dataKey.data(), km.aad.data(), 32, back.data(), back.size(), osz.data());
            CHECK(rc == 0);
            CHECK(memcmp(back.data(), seg.data(), n) == 0);
        } else CHECK(object == seg);
        if (object.size() > (size_t)bucket.capacity()) CHECK(slept > 0);       // the limiter was charged for every part
    }
}

int main(int argc, char** argv) {
    builderTests();
    bytesRangeTests();
    uploadSideTests();
    manifestTests();
    baseAndFinisherTests();
    if (argc > 1 && !strcmp(argv[1], "--device")) {
        tsgpu_ctx* ctx = nullptr;
        int rc = tsgpu_create(nullptr, 0, 400000, 16, &ctx);
        if (rc) { printf("tsgpu_create failed: %s\n", tsgpu_last_error()); return 2; }
        gpuChainTests(ctx);
        segmentUploadTests(ctx);
        referenceManifestTests(ctx);
        indexFileTests(ctx);
        tsgpu_destroy(ctx);
    }
    printf("%s: %d checks, %d failures\n", failures ? "FAILED" : "OK", checks, failures);
    return failures ? 1 : 0;
}
