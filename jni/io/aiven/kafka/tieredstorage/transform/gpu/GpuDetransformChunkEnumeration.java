package io.aiven.kafka.tieredstorage.transform.gpu;

import java.io.IOException;
import java.io.InputStream;
import java.nio.ByteBuffer;
import java.nio.channels.Channels;
import java.nio.channels.ReadableByteChannel;
import java.util.ArrayDeque;
import java.util.List;
import java.util.NoSuchElementException;
import java.util.Objects;

import io.aiven.kafka.tieredstorage.Chunk;
import io.aiven.kafka.tieredstorage.transform.DetransformChunkEnumeration;

/**
 * Drop-in for the chain BaseDetransform -> [Decryption] -> [Decompression] built by DefaultChunkManager.getChunk
 * (DefaultChunkManager.java:50-70), for a LIST of consecutive chunks: the transformed bytes of all of them are read from the
 * (ranged) object stream into pinned memory, detransformed by one native call, and served one byte[] per nextElement() —
 * the DetransformChunkEnumeration contract (DetransformChunkEnumeration.java:28-29) DetransformFinisher consumes.
 *
 * <p>Errors keep the reference's shape: a stream shorter than the chunks announce is
 * "Stream has fewer bytes than expected" (BaseDetransformChunkEnumeration.java:106-108), a failed tag or a malformed
 * frame is a RuntimeException from the native call (DecryptionChunkEnumeration.java:59-61,
 * DecompressionChunkEnumeration.java:42-44).  The stream is closed once the last chunk has been read
 * (BaseDetransformChunkEnumeration.java:83-95).
 */
public class GpuDetransformChunkEnumeration implements DetransformChunkEnumeration, AutoCloseable {
    private final long ctx;
    private final PinnedPool pool;
    private final InputStream inputStream;
    private final List<Chunk> chunks;
    private final int flags;
    private final byte[] key;
    private final byte[] aad;
    private final ArrayDeque<byte[]> ready = new ArrayDeque<>();
    private boolean done = false;

    public GpuDetransformChunkEnumeration(final long ctx, final PinnedPool pool, final InputStream inputStream,
                                          final List<Chunk> chunks, final boolean compression,
                                          final byte[] dataKey, final byte[] aad) {
        this.ctx = ctx;
        this.pool = pool;
        this.inputStream = Objects.requireNonNull(inputStream, "inputStream cannot be null");
        this.chunks = Objects.requireNonNull(chunks, "chunks cannot be null");
        this.flags = (compression ? TsGpu.FLAG_ZSTD : 0) | (dataKey != null ? TsGpu.FLAG_AES : 0);
        this.key = dataKey;
        this.aad = aad;
    }

    @Override
    public boolean hasMoreElements() {
        fill();
        return !ready.isEmpty();
    }

    @Override
    public byte[] nextElement() {
        fill();
        if (ready.isEmpty()) {
            throw new NoSuchElementException();
        }
        return ready.poll();
    }

    private void fill() {
        if (done) {
            return;
        }
        done = true;
        if (chunks.isEmpty()) {                       // no chunking was applied: everything at once
            try (InputStream in = inputStream) {
                ready.add(in.readAllBytes());
            } catch (final IOException e) {
                throw new RuntimeException(e);
            }
            return;
        }
        long transformed = 0;
        long original = 0;
        final int[] tsizes = new int[chunks.size()];
        for (int i = 0; i < tsizes.length; i++) {
            tsizes[i] = chunks.get(i).transformedSize;
            transformed += tsizes[i];
            original += chunks.get(i).originalSize;
        }
        try (PinnedPool.Lease src = pool.lease(transformed);
             PinnedPool.Lease dst = pool.lease(Math.max(original, 1));
             InputStream in = inputStream) {
            final ByteBuffer sb = src.buffer();
            sb.clear();
            sb.limit((int) transformed);
            final ReadableByteChannel ch = Channels.newChannel(in);
            while (sb.hasRemaining()) {               // the ranged GET lands directly in pinned memory
                if (ch.read(sb) < 0) {
                    throw new RuntimeException("Stream has fewer bytes than expected");   // the reference's type and message
                }
            }
            final int[] osizes = new int[tsizes.length];
            final ByteBuffer db = dst.buffer();
            TsGpu.detransform(ctx, flags, sb, transformed, tsizes, key, aad, db, osizes);
            db.clear();
            for (final int n : osizes) {
                final byte[] chunk = new byte[n];
                db.get(chunk);
                ready.add(chunk);
            }
        } catch (final IOException e) {
            throw new RuntimeException(e);
        }
    }

    @Override
    public void close() {
        done = true;
        ready.clear();
        try {
            inputStream.close();
        } catch (final IOException e) {
            throw new RuntimeException(e);
        }
    }
}
