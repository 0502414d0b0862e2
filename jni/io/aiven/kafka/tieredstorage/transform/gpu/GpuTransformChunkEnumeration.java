package io.aiven.kafka.tieredstorage.transform.gpu;

import java.io.IOException;
import java.io.InputStream;
import java.nio.ByteBuffer;
import java.security.SecureRandom;
import java.util.ArrayDeque;
import java.util.NoSuchElementException;

import io.aiven.kafka.tieredstorage.security.DataKeyAndAAD;
import io.aiven.kafka.tieredstorage.transform.TransformChunkEnumeration;

/**
 * Drop-in for the chain Base -> [Compression] -> [Encryption] built by RemoteStorageManager.transformation
 * (RemoteStorageManager.java:434-453): reads a batch of original chunks from the segment stream, transforms the
 * whole batch on the GPU with one native call and serves one byte[] per nextElement(), so TransformFinisher,
 * the ChunkIndex builders and the storage backends are unchanged.
 */
public class GpuTransformChunkEnumeration implements TransformChunkEnumeration {
    private final long ctx;
    private final InputStream inputStream;
    private final int originalChunkSize;
    private final int flags;
    private final DataKeyAndAAD keyAndAad;   // null when encryption is off
    private final int batchChunks;
    private final SecureRandom random = new SecureRandom();
    private final ArrayDeque<byte[]> ready = new ArrayDeque<>();
    private final ByteBuffer src;
    private final ByteBuffer dst;
    private boolean eof = false;

    public GpuTransformChunkEnumeration(final long ctx, final InputStream inputStream, final int originalChunkSize,
                                        final boolean compression, final DataKeyAndAAD keyAndAad, final int batchChunks) {
        this.ctx = ctx;
        this.inputStream = inputStream;
        this.originalChunkSize = originalChunkSize;
        this.flags = (compression ? TsGpu.FLAG_ZSTD : 0) | (keyAndAad != null ? TsGpu.FLAG_AES : 0);
        this.keyAndAad = keyAndAad;
        this.batchChunks = batchChunks;
        this.src = TsGpu.allocPinned((long) batchChunks * originalChunkSize);
        this.dst = TsGpu.allocPinned(TsGpu.transformBound(flags, (long) batchChunks * originalChunkSize, originalChunkSize));
    }

    @Override
    public int originalChunkSize() {
        return originalChunkSize;
    }

    @Override
    public Integer transformedChunkSize() {
        // Same arithmetic as the decorators: variable with compression, +28 with encryption only.
        if ((flags & TsGpu.FLAG_ZSTD) != 0) {
            return null;
        }
        return (flags & TsGpu.FLAG_AES) != 0 ? originalChunkSize + 28 : originalChunkSize;
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
        if (!ready.isEmpty() || eof) {
            return;
        }
        try {
            src.clear();
            final byte[] buf = new byte[originalChunkSize];
            int chunks = 0;
            long total = 0;
            while (chunks < batchChunks) {
                final int n = inputStream.readNBytes(buf, 0, originalChunkSize);
                if (n == 0) {
                    eof = true;
                    break;
                }
                src.put(buf, 0, n);
                total += n;
                chunks++;
                if (n < originalChunkSize) {
                    eof = true;
                    break;
                }
            }
            if (chunks == 0) {
                return;
            }
            final byte[] ivs = new byte[chunks * TsGpu.IV_SIZE];
            random.nextBytes(ivs);
            final int[] sizes = new int[chunks];
            final int n = TsGpu.transform(ctx, flags, src, total, originalChunkSize,
                keyAndAad == null ? null : keyAndAad.dataKey.getEncoded(), keyAndAad == null ? null : keyAndAad.aad,
                ivs, dst, sizes);
            dst.clear();
            for (int i = 0; i < n; i++) {
                final byte[] chunk = new byte[sizes[i]];
                dst.get(chunk);
                ready.add(chunk);
            }
        } catch (final IOException e) {
            throw new RuntimeException(e);
        }
    }
}
