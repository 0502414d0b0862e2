package io.aiven.kafka.tieredstorage.fetch.gpu;

import java.io.ByteArrayInputStream;
import java.io.IOException;
import java.io.InputStream;
import java.util.ArrayList;
import java.util.List;
import java.util.Optional;

import io.aiven.kafka.tieredstorage.Chunk;
import io.aiven.kafka.tieredstorage.fetch.ChunkManager;
import io.aiven.kafka.tieredstorage.manifest.SegmentEncryptionMetadata;
import io.aiven.kafka.tieredstorage.manifest.SegmentManifest;
import io.aiven.kafka.tieredstorage.storage.BytesRange;
import io.aiven.kafka.tieredstorage.storage.ObjectFetcher;
import io.aiven.kafka.tieredstorage.storage.ObjectKey;
import io.aiven.kafka.tieredstorage.storage.StorageBackendException;
import io.aiven.kafka.tieredstorage.transform.DetransformFinisher;
import io.aiven.kafka.tieredstorage.transform.gpu.GpuDetransformChunkEnumeration;
import io.aiven.kafka.tieredstorage.transform.gpu.PinnedPool;

/**
 * Batched replacement for DefaultChunkManager (DefaultChunkManager.java:50-70).  getChunk(key, manifest, id) keeps the
 * ChunkManager contract — FetchChunkEnumeration (FetchChunkEnumeration.java:100-138) and ChunkCache work unchanged — and
 * getChunks(key, manifest, first, count) is what makes ranged fetches fast: ONE ranged GET for the transformed bytes of the
 * chunks and ONE detransform call for all of them, instead of a GET and a cipher/zstd context per chunk.  ChunkCache's
 * prefetch (ChunkCache.java:159-184) calls it with the prefetch window and stores the chunks individually.
 */
public class GpuChunkManager implements ChunkManager {
    private final ObjectFetcher fetcher;
    private final long ctx;
    private final PinnedPool pool;

    public GpuChunkManager(final ObjectFetcher fetcher, final long ctx, final PinnedPool pool) {
        this.fetcher = fetcher;
        this.ctx = ctx;
        this.pool = pool;
    }

    @Override
    public InputStream getChunk(final ObjectKey objectKey, final SegmentManifest manifest,
                                final int chunkId) throws StorageBackendException, IOException {
        return getChunks(objectKey, manifest, chunkId, 1).get(0);
    }

    /** Plain-text streams of chunks [firstChunkId, firstChunkId + count), fetched and detransformed as one batch. */
    public List<InputStream> getChunks(final ObjectKey objectKey, final SegmentManifest manifest,
                                       final int firstChunkId, final int count) throws StorageBackendException, IOException {
        final List<Chunk> all = manifest.chunkIndex().chunks();
        final List<Chunk> chunks = all.subList(firstChunkId, Math.min(all.size(), firstChunkId + count));
        final Chunk first = chunks.get(0);
        final Chunk last = chunks.get(chunks.size() - 1);
        final BytesRange range = BytesRange.of(first.transformedPosition, last.transformedPosition + last.transformedSize - 1);
        final Optional<SegmentEncryptionMetadata> encryption = manifest.encryption();
        final List<InputStream> result = new ArrayList<>(chunks.size());
        try (GpuDetransformChunkEnumeration detransform = new GpuDetransformChunkEnumeration(
            ctx, pool, fetcher.fetch(objectKey, range), chunks, manifest.compression(),
            encryption.map(e -> e.dataKey().getEncoded()).orElse(null), encryption.map(SegmentEncryptionMetadata::aad).orElse(null))) {
            // DetransformFinisher would concatenate; the cache wants the chunks one by one
            while (detransform.hasMoreElements()) {
                result.add(new ByteArrayInputStream(detransform.nextElement()));
            }
        }
        return result;
    }

    /** The whole range as one stream, for callers that do not cache (DetransformFinisher.toInputStream semantics). */
    public InputStream getRange(final ObjectKey objectKey, final SegmentManifest manifest,
                                final int firstChunkId, final int count) throws StorageBackendException, IOException {
        final List<Chunk> all = manifest.chunkIndex().chunks();
        final List<Chunk> chunks = all.subList(firstChunkId, Math.min(all.size(), firstChunkId + count));
        final Chunk first = chunks.get(0);
        final Chunk last = chunks.get(chunks.size() - 1);
        final BytesRange range = BytesRange.of(first.transformedPosition, last.transformedPosition + last.transformedSize - 1);
        final Optional<SegmentEncryptionMetadata> encryption = manifest.encryption();
        return new DetransformFinisher(new GpuDetransformChunkEnumeration(
            ctx, pool, fetcher.fetch(objectKey, range), chunks, manifest.compression(),
            encryption.map(e -> e.dataKey().getEncoded()).orElse(null), encryption.map(SegmentEncryptionMetadata::aad).orElse(null)))
            .toInputStream();
    }
}
