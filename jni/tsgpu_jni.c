/*
 * tsgpu_jni.c — JNI glue between the reference's Java operator surface and the C-ABI of include/tsgpu.h.
 * There is no JDK in this repository's build image: the tests compile this file against a test-only stand-in for
 * <jni.h> and drive it through a fake JNIEnv (tests/cpp/test_jni_shim.c).  Build on the broker host with
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude jni/tsgpu_jni.c -L<pkg> -ltsgpu -o libtsgpu_jni.so
 * Java side: jni/io/aiven/kafka/tieredstorage/transform/gpu/{TsGpu,PinnedPool,GpuTransformChunkEnumeration,GpuDetransformChunkEnumeration}.java
 * (implement TransformChunkEnumeration / DetransformChunkEnumeration of the reference, core/M/transform/TransformChunkEnumeration.java:28-42,
 * DetransformChunkEnumeration.java:28-29) and jni/io/aiven/kafka/tieredstorage/fetch/gpu/GpuChunkManager.java (ChunkManager, fetch/ChunkManager.java).
 *
 * Buffers are direct ByteBuffers (ideally over tsgpu_host_alloc'ed pinned memory, see allocPinned), so there is no
 * JNI array copy on the data path.  Errors become RuntimeException(tsgpu_last_error()), which is what the
 * reference's operators throw and RemoteStorageManager converts (RemoteStorageManager.java:258-267, :566-575).
 */
#include <jni.h>
#include <stdint.h>
#include "tsgpu.h"

#define CLS "io/aiven/kafka/tieredstorage/transform/gpu/TsGpu"

static void throw_rt(JNIEnv* env, int rc) {
    jclass ex = (*env)->FindClass(env, rc == TSGPU_E_ARG ? "java/lang/IllegalArgumentException" : "java/lang/RuntimeException");
    (*env)->ThrowNew(env, ex, tsgpu_last_error());
}
static void throw_arg(JNIEnv* env, const char* msg) {
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalArgumentException"), msg);
}
/* A direct ByteBuffer's address and capacity, or NULL with an IllegalArgumentException pending: a heap buffer has no
 * address (GetDirectBufferAddress gives NULL, the capacity -1) and `need` bytes must fit. */
static uint8_t* direct(JNIEnv* env, jobject buf, jlong need, jlong* cap, const char* what) {
    if (!buf) { throw_arg(env, what); return NULL; }
    uint8_t* p = (uint8_t*)(*env)->GetDirectBufferAddress(env, buf);
    const jlong c = (*env)->GetDirectBufferCapacity(env, buf);
    if (!p || c < 0) { throw_arg(env, what); return NULL; }
    if (need < 0 || need > c) { throw_arg(env, "length exceeds the direct buffer's capacity"); return NULL; }
    if (cap) *cap = c;
    return p;
}

JNIEXPORT jlong JNICALL Java_io_aiven_kafka_tieredstorage_transform_gpu_TsGpu_create(JNIEnv* env, jclass c, jintArray devices,
                                                                                     jint maxChunkBytes, jint maxBatch) {
    tsgpu_ctx* ctx = NULL;
    jint n = devices ? (*env)->GetArrayLength(env, devices) : 0;
    jint* ids = n ? (*env)->GetIntArrayElements(env, devices, NULL) : NULL;
    int rc = tsgpu_create((const int*)ids, n, (uint32_t)maxChunkBytes, (uint32_t)maxBatch, &ctx);
    if (ids) (*env)->ReleaseIntArrayElements(env, devices, ids, JNI_ABORT);
    if (rc) { throw_rt(env, rc); return 0; }
    return (jlong)(intptr_t)ctx;
}
JNIEXPORT void JNICALL Java_io_aiven_kafka_tieredstorage_transform_gpu_TsGpu_destroy(JNIEnv* env, jclass c, jlong h) {
    tsgpu_destroy((tsgpu_ctx*)(intptr_t)h);
}
JNIEXPORT jobject JNICALL Java_io_aiven_kafka_tieredstorage_transform_gpu_TsGpu_allocPinned(JNIEnv* env, jclass c, jlong bytes) {
    void* p = tsgpu_host_alloc((size_t)bytes);
    return p ? (*env)->NewDirectByteBuffer(env, p, bytes) : NULL;
}
JNIEXPORT void JNICALL Java_io_aiven_kafka_tieredstorage_transform_gpu_TsGpu_freePinned(JNIEnv* env, jclass c, jobject buf) {
    if (buf) tsgpu_host_free((*env)->GetDirectBufferAddress(env, buf));
}
JNIEXPORT jlong JNICALL Java_io_aiven_kafka_tieredstorage_transform_gpu_TsGpu_transformBound(JNIEnv* env, jclass c, jint flags,
                                                                                             jlong srcLen, jint chunkSize) {
    return (jlong)tsgpu_transform_bound((uint32_t)flags, (uint64_t)srcLen, (uint32_t)chunkSize);
}
/* returns the number of chunks; transformedSizes (int[]) receives one entry per chunk */
JNIEXPORT jint JNICALL Java_io_aiven_kafka_tieredstorage_transform_gpu_TsGpu_transform(
    JNIEnv* env, jclass c, jlong h, jint flags, jobject src, jlong srcLen, jint chunkSize, jbyteArray key, jbyteArray aad,
    jbyteArray ivs, jobject dst, jintArray transformedSizes) {
    jlong dcap = 0;
    uint8_t* s = direct(env, src, srcLen, NULL, "src must be a direct ByteBuffer");
    if (!s) return -1;
    uint8_t* d = direct(env, dst, 0, &dcap, "dst must be a direct ByteBuffer");
    if (!d) return -1;
    if (!transformedSizes) { throw_arg(env, "transformedSizes cannot be null"); return -1; }
    if ((flags & TSGPU_FLAG_AES) && key && (*env)->GetArrayLength(env, key) != 32) { throw_arg(env, "data key must be 32 bytes"); return -1; }
    jbyte* k = key ? (*env)->GetByteArrayElements(env, key, NULL) : NULL;
    jbyte* a = aad ? (*env)->GetByteArrayElements(env, aad, NULL) : NULL;
    jbyte* iv = ivs ? (*env)->GetByteArrayElements(env, ivs, NULL) : NULL;
    jint alen = aad ? (*env)->GetArrayLength(env, aad) : 0;
    uint32_t n = (uint32_t)(*env)->GetArrayLength(env, transformedSizes);
    if ((flags & TSGPU_FLAG_AES) && iv && chunkSize > 0 &&
        (jlong)(*env)->GetArrayLength(env, ivs) < TSGPU_IV_SIZE * ((srcLen + chunkSize - 1) / chunkSize)) {
        if (k) (*env)->ReleaseByteArrayElements(env, key, k, JNI_ABORT);
        if (a) (*env)->ReleaseByteArrayElements(env, aad, a, JNI_ABORT);
        (*env)->ReleaseByteArrayElements(env, ivs, iv, JNI_ABORT);
        throw_arg(env, "ivs holds fewer than 12 bytes per chunk");
        return -1;
    }
    jint* sizes = (*env)->GetIntArrayElements(env, transformedSizes, NULL);
    int rc = tsgpu_transform((tsgpu_ctx*)(intptr_t)h, (uint32_t)flags, s, (uint64_t)srcLen, (uint32_t)chunkSize, (const uint8_t*)k,
                             (const uint8_t*)a, (uint32_t)alen, (const uint8_t*)iv, d, (uint64_t)dcap, (uint32_t*)sizes, &n);
    (*env)->ReleaseIntArrayElements(env, transformedSizes, sizes, 0);
    if (k) (*env)->ReleaseByteArrayElements(env, key, k, JNI_ABORT);
    if (a) (*env)->ReleaseByteArrayElements(env, aad, a, JNI_ABORT);
    if (iv) (*env)->ReleaseByteArrayElements(env, ivs, iv, JNI_ABORT);
    if (rc) { throw_rt(env, rc); return -1; }
    return (jint)n;
}
JNIEXPORT void JNICALL Java_io_aiven_kafka_tieredstorage_transform_gpu_TsGpu_detransform(
    JNIEnv* env, jclass c, jlong h, jint flags, jobject src, jlong srcLen, jintArray transformedSizes, jbyteArray key, jbyteArray aad,
    jobject dst, jintArray originalSizes) {
    jlong dcap = 0;
    uint8_t* s = direct(env, src, srcLen, NULL, "src must be a direct ByteBuffer");
    if (!s) return;
    uint8_t* d = direct(env, dst, 0, &dcap, "dst must be a direct ByteBuffer");
    if (!d) return;
    if (!transformedSizes || !originalSizes) { throw_arg(env, "chunks cannot be null"); return; }
    if ((*env)->GetArrayLength(env, originalSizes) < (*env)->GetArrayLength(env, transformedSizes)) { throw_arg(env, "originalSizes is shorter than transformedSizes"); return; }
    if ((flags & TSGPU_FLAG_AES) && key && (*env)->GetArrayLength(env, key) != 32) { throw_arg(env, "data key must be 32 bytes"); return; }
    jbyte* k = key ? (*env)->GetByteArrayElements(env, key, NULL) : NULL;
    jbyte* a = aad ? (*env)->GetByteArrayElements(env, aad, NULL) : NULL;
    jint alen = aad ? (*env)->GetArrayLength(env, aad) : 0;
    uint32_t n = (uint32_t)(*env)->GetArrayLength(env, transformedSizes);
    jint* ts = (*env)->GetIntArrayElements(env, transformedSizes, NULL);
    jint* os = (*env)->GetIntArrayElements(env, originalSizes, NULL);
    int rc = tsgpu_detransform((tsgpu_ctx*)(intptr_t)h, (uint32_t)flags, s, (uint64_t)srcLen, (const uint32_t*)ts, n, (const uint8_t*)k,
                               (const uint8_t*)a, (uint32_t)alen, d, (uint64_t)dcap, (uint32_t*)os);
    (*env)->ReleaseIntArrayElements(env, transformedSizes, ts, JNI_ABORT);
    (*env)->ReleaseIntArrayElements(env, originalSizes, os, 0);
    if (k) (*env)->ReleaseByteArrayElements(env, key, k, JNI_ABORT);
    if (a) (*env)->ReleaseByteArrayElements(env, aad, a, JNI_ABORT);
    if (rc) throw_rt(env, rc);
}
