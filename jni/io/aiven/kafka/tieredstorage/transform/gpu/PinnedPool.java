package io.aiven.kafka.tieredstorage.transform.gpu;

import java.lang.ref.Cleaner;
import java.nio.ByteBuffer;
import java.util.ArrayDeque;
import java.util.concurrent.atomic.AtomicBoolean;

/**
 * Page-locked host buffers shared by the enumerations of one context.  cudaHostAlloc costs milliseconds and its memory is
 * invisible to the garbage collector (a direct ByteBuffer over it frees nothing when collected), so the buffers are leased,
 * returned and reused; whatever the pool still holds is released with {@link #close()} (RemoteStorageManager.close()).
 * A lease that is never closed is returned by a {@link Cleaner} when its owner becomes unreachable.
 */
public final class PinnedPool implements AutoCloseable {
    private static final Cleaner CLEANER = Cleaner.create();
    private final ArrayDeque<ByteBuffer> free = new ArrayDeque<>();
    private final int maxIdle;
    private boolean closed = false;

    public PinnedPool(final int maxIdleBuffers) {
        this.maxIdle = maxIdleBuffers;
    }

    /** A buffer of at least {@code bytes} bytes (the smallest idle one that fits, else a new allocation). */
    public synchronized Lease lease(final long bytes) {
        if (closed) {
            throw new IllegalStateException("pinned pool is closed");
        }
        ByteBuffer best = null;
        for (final ByteBuffer b : free) {
            if (b.capacity() >= bytes && (best == null || b.capacity() < best.capacity())) {
                best = b;
            }
        }
        if (best != null) {
            free.remove(best);
        } else {
            best = TsGpu.allocPinned(Math.max(bytes, 1));
            if (best == null) {
                throw new OutOfMemoryError("cudaHostAlloc of " + bytes + " bytes failed");
            }
        }
        return new Lease(this, best);
    }

    private synchronized void giveBack(final ByteBuffer b) {
        if (closed || free.size() >= maxIdle) {
            TsGpu.freePinned(b);
        } else {
            free.add(b);
        }
    }

    @Override
    public synchronized void close() {
        closed = true;
        for (final ByteBuffer b : free) {
            TsGpu.freePinned(b);
        }
        free.clear();
    }

    /** One leased buffer.  close() returns it; unreachable leases are returned by the cleaner. */
    public static final class Lease implements AutoCloseable {
        private final ByteBuffer buffer;
        private final AtomicBoolean returned = new AtomicBoolean(false);
        private final Cleaner.Cleanable cleanable;

        Lease(final PinnedPool pool, final ByteBuffer buffer) {
            this.buffer = buffer;
            final AtomicBoolean flag = returned;
            this.cleanable = CLEANER.register(this, () -> {
                if (flag.compareAndSet(false, true)) {
                    pool.giveBack(buffer);
                }
            });
        }

        public ByteBuffer buffer() {
            if (returned.get()) {
                throw new IllegalStateException("pinned buffer already returned");
            }
            return buffer;
        }

        @Override
        public void close() {
            cleanable.clean();
        }
    }
}
