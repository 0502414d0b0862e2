package io.aiven.kafka.tieredstorage.transform.gpu;

import java.nio.ByteBuffer;

/** Native entry points of libtsgpu (see include/tsgpu.h and jni/tsgpu_jni.c). */
public final class TsGpu {
    public static final int FLAG_ZSTD = 1;
    public static final int FLAG_AES = 2;
    public static final int IV_SIZE = 12;

    static {
        System.loadLibrary("tsgpu_jni");
    }

    private TsGpu() {
    }

    public static native long create(int[] devices, int maxChunkBytes, int maxBatch);

    public static native void destroy(long ctx);

    public static native ByteBuffer allocPinned(long bytes);

    public static native void freePinned(ByteBuffer buffer);

    public static native long transformBound(int flags, long srcLen, int chunkSize);

    public static native int transform(long ctx, int flags, ByteBuffer src, long srcLen, int chunkSize,
                                       byte[] key, byte[] aad, byte[] ivs, ByteBuffer dst, int[] transformedSizes);

    public static native void detransform(long ctx, int flags, ByteBuffer src, long srcLen, int[] transformedSizes,
                                          byte[] key, byte[] aad, ByteBuffer dst, int[] originalSizes);
}
