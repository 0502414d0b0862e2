/*
 * Drives jni/tsgpu_jni.c through a fake JNIEnv (tests/jni_stub/jni.h): the glue is compiled as it stands and its native
 * methods are called the way the JVM would call them — Java arrays are COPIED on Get*ArrayElements and only written back
 * on Release with mode 0 / JNI_COMMIT (JNI_ABORT discards), which is exactly what the glue's release modes must get right.
 * Checks the round trip TsGpu.transform -> TsGpu.detransform, the exception mapping, and the oracle as the reference-side
 * reader.  Links the test-only SIMT build here (no GPU in the build box).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <jni.h>
#include "../../jni/tsgpu_jni.c"
#include "../../oracle/tsoracle.h"

struct _jobject { int kind; void* data; jlong len; };      /* kind: 1 int[], 2 byte[], 3 direct buffer, 4 class */
static int checks = 0, failures = 0;
#define CHECK(c) do { checks++; if (!(c)) { failures++; printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); } } while (0)
static char thrown_class[128], thrown_msg[512];
static int live_copies = 0;

static jclass f_FindClass(JNIEnv* e, const char* name) { struct _jobject* o = calloc(1, sizeof *o); o->kind = 4; o->data = strdup(name); return o; }
static jint f_ThrowNew(JNIEnv* e, jclass c, const char* msg) { snprintf(thrown_class, sizeof thrown_class, "%s", (char*)c->data); snprintf(thrown_msg, sizeof thrown_msg, "%s", msg); return 0; }
static jsize f_GetArrayLength(JNIEnv* e, jarray a) { return (jsize)a->len; }
static void* copy_out(jarray a, size_t elem) { void* p = malloc(a->len * elem + 1); memcpy(p, a->data, a->len * elem); live_copies++; return p; }
static void copy_back(jarray a, void* p, size_t elem, jint mode) {
    if (mode != JNI_ABORT) memcpy(a->data, p, a->len * elem);
    if (mode != JNI_COMMIT) { free(p); live_copies--; }
}
static jbyte* f_GetByteArrayElements(JNIEnv* e, jbyteArray a, jboolean* c) { if (c) *c = 1; return copy_out(a, 1); }
static void f_ReleaseByteArrayElements(JNIEnv* e, jbyteArray a, jbyte* p, jint m) { copy_back(a, p, 1, m); }
static jint* f_GetIntArrayElements(JNIEnv* e, jintArray a, jboolean* c) { if (c) *c = 1; return copy_out(a, 4); }
static void f_ReleaseIntArrayElements(JNIEnv* e, jintArray a, jint* p, jint m) { copy_back(a, p, 4, m); }
static jobject f_NewDirectByteBuffer(JNIEnv* e, void* p, jlong cap) { struct _jobject* o = calloc(1, sizeof *o); o->kind = 3; o->data = p; o->len = cap; return o; }
static void* f_GetDirectBufferAddress(JNIEnv* e, jobject b) { return b->data; }
static jlong f_GetDirectBufferCapacity(JNIEnv* e, jobject b) { return b->len; }
static const struct JNINativeInterface_ table = { f_FindClass, f_ThrowNew, f_GetArrayLength, f_GetByteArrayElements, f_ReleaseByteArrayElements,
    f_GetIntArrayElements, f_ReleaseIntArrayElements, f_NewDirectByteBuffer, f_GetDirectBufferAddress, f_GetDirectBufferCapacity };

static jarray mk_array(int kind, const void* src, jlong n, size_t elem) {
    struct _jobject* o = calloc(1, sizeof *o); o->kind = kind; o->len = n; o->data = calloc((size_t)n + 1, elem);
    if (src) memcpy(o->data, src, (size_t)n * elem);
    return o;
}
#define NS(f) Java_io_aiven_kafka_tieredstorage_transform_gpu_TsGpu_##f

int main(void) {
    JNIEnv envp = &table; JNIEnv* env = &envp;
    const jint cs = 20000, n = 70001, nch = 4;
    jlong h = NS(create)(env, NULL, NULL, cs, 3);
    CHECK(h != 0 && thrown_msg[0] == 0);
    jint dev0 = 0; jarray devs = mk_array(1, &dev0, 1, 4);
    jlong h2 = NS(create)(env, NULL, devs, cs, 2); CHECK(h2 != 0); NS(destroy)(env, NULL, h2);
    CHECK(NS(create)(env, NULL, NULL, 0, 3) == 0);                                  /* IllegalArgumentException path */
    CHECK(strcmp(thrown_class, "java/lang/IllegalArgumentException") == 0 && strstr(thrown_msg, "max_chunk_bytes") != NULL);
    thrown_msg[0] = 0;

    const jlong bound = NS(transformBound)(env, NULL, 3, n, cs);
    CHECK(bound >= n + 28 * nch);
    jobject src = NS(allocPinned)(env, NULL, n), dst = NS(allocPinned)(env, NULL, bound), back = NS(allocPinned)(env, NULL, n);
    CHECK(src && dst && back && src->len == n);
    uint8_t* s = src->data;
    for (int i = 0; i < n; i++) s[i] = (uint8_t)("offset=42 key=user-17 value={\"a\":1}\n"[i % 37] + (i / 9000));
    uint8_t key[32], aad[32], ivs[12 * 4];
    for (int i = 0; i < 32; i++) { key[i] = (uint8_t)(i * 7 + 1); aad[i] = (uint8_t)(200 - i); }
    for (int i = 0; i < 48; i++) ivs[i] = (uint8_t)(i * 3 + 5);
    jarray jkey = mk_array(2, key, 32, 1), jaad = mk_array(2, aad, 32, 1), jivs = mk_array(2, ivs, 48, 1);
    jarray sizes = mk_array(1, NULL, nch, 4), osz = mk_array(1, NULL, nch, 4);
    jint got = NS(transform)(env, NULL, h, 3, src, n, cs, jkey, jaad, jivs, dst, sizes);
    CHECK(got == nch && thrown_msg[0] == 0);
    uint64_t total = 0; for (int i = 0; i < nch; i++) { CHECK(((jint*)sizes->data)[i] > 28); total += ((jint*)sizes->data)[i]; }   /* written back (mode 0) */
    CHECK(memcmp(jkey->data, key, 32) == 0);                                       /* inputs released with JNI_ABORT: untouched */
    /* the reference-side reader (oracle: libzstd + OpenSSL) recovers the segment */
    uint8_t* plain = malloc(n + 64); uint32_t o2[4];
    CHECK(ora_detransform_chunks(3, dst->data, (const uint32_t*)sizes->data, nch, key, aad, 32, plain, n + 64, o2) == 0);
    CHECK(memcmp(plain, s, n) == 0);
    /* and TsGpu.detransform does */
    NS(detransform)(env, NULL, h, 3, dst, (jlong)total, sizes, jkey, jaad, back, osz);
    CHECK(thrown_msg[0] == 0 && memcmp(back->data, s, n) == 0);
    CHECK(((jint*)osz->data)[0] == cs && ((jint*)osz->data)[3] == n - 3 * cs);
    /* a flipped ciphertext bit: RuntimeException("Tag mismatch ...") like AEADBadTagException wrapped by the reference */
    ((uint8_t*)dst->data)[40] ^= 1;
    NS(detransform)(env, NULL, h, 3, dst, (jlong)total, sizes, jkey, jaad, back, osz);
    CHECK(strcmp(thrown_class, "java/lang/RuntimeException") == 0 && strstr(thrown_msg, "Tag mismatch") != NULL);
    thrown_msg[0] = 0;
    /* too small an int[] for the sizes: error, nothing thrown away silently */
    jarray small = mk_array(1, NULL, 2, 4);
    CHECK(NS(transform)(env, NULL, h, 3, src, n, cs, jkey, jaad, jivs, dst, small) == -1 && thrown_msg[0] != 0);
    thrown_msg[0] = 0;
    /* a heap (non-direct) buffer has no address; a length beyond the capacity; null arrays -> IllegalArgumentException, no crash */
    struct _jobject heap = { 3, NULL, -1 };
    CHECK(NS(transform)(env, NULL, h, 3, &heap, n, cs, jkey, jaad, jivs, dst, sizes) == -1);
    CHECK(strcmp(thrown_class, "java/lang/IllegalArgumentException") == 0 && strstr(thrown_msg, "direct ByteBuffer") != NULL); thrown_msg[0] = 0;
    CHECK(NS(transform)(env, NULL, h, 3, src, (jlong)n + 1, cs, jkey, jaad, jivs, dst, sizes) == -1 && strstr(thrown_msg, "capacity") != NULL); thrown_msg[0] = 0;
    CHECK(NS(transform)(env, NULL, h, 3, src, n, cs, jkey, jaad, jivs, dst, NULL) == -1 && strstr(thrown_msg, "transformedSizes") != NULL); thrown_msg[0] = 0;
    jarray shortiv = mk_array(2, ivs, 24, 1);
    CHECK(NS(transform)(env, NULL, h, 3, src, n, cs, jkey, jaad, shortiv, dst, sizes) == -1 && strstr(thrown_msg, "12 bytes per chunk") != NULL); thrown_msg[0] = 0;
    jarray badkey = mk_array(2, key, 16, 1);
    CHECK(NS(transform)(env, NULL, h, 3, src, n, cs, badkey, jaad, jivs, dst, sizes) == -1 && strstr(thrown_msg, "32 bytes") != NULL); thrown_msg[0] = 0;
    NS(detransform)(env, NULL, h, 3, dst, (jlong)total, sizes, jkey, jaad, back, NULL);
    CHECK(strstr(thrown_msg, "chunks cannot be null") != NULL); thrown_msg[0] = 0;
    NS(detransform)(env, NULL, h, 3, dst, (jlong)total, sizes, jkey, jaad, &heap, osz);
    CHECK(strstr(thrown_msg, "direct ByteBuffer") != NULL); thrown_msg[0] = 0;
    NS(freePinned)(env, NULL, NULL);                                               /* tolerated */

    /* the flow of GpuDetransformChunkEnumeration / GpuChunkManager.getChunks: a window of consecutive chunks of the object
     * (ranged GET = bytes [pos(first), end(last)) read into a pinned buffer), one detransform call, chunks served one by one */
    ((uint8_t*)dst->data)[40] ^= 1;                                                /* undo the flipped bit */
    {
        const jint first = 1, cnt = 2;
        jlong off = 0, win = 0, orig = 0;
        for (int i = 0; i < first; i++) off += ((jint*)sizes->data)[i];
        for (int i = first; i < first + cnt; i++) { win += ((jint*)sizes->data)[i]; orig += (i == nch - 1) ? n - 3 * cs : cs; }
        jobject wsrc = NS(allocPinned)(env, NULL, win), wdst = NS(allocPinned)(env, NULL, orig);
        memcpy(wsrc->data, (uint8_t*)dst->data + off, (size_t)win);
        jarray wts = mk_array(1, (jint*)sizes->data + first, cnt, 4), wos = mk_array(1, NULL, cnt, 4);
        NS(detransform)(env, NULL, h, 3, wsrc, win, wts, jkey, jaad, wdst, wos);
        CHECK(thrown_msg[0] == 0 && memcmp(wdst->data, s + (size_t)first * cs, (size_t)orig) == 0);
        CHECK(((jint*)wos->data)[0] == cs && ((jint*)wos->data)[1] == cs);
        /* a ranged GET that came back short: "Stream has fewer bytes than expected" (BaseDetransformChunkEnumeration.java:106-108) */
        NS(detransform)(env, NULL, h, 3, wsrc, win - 1, wts, jkey, jaad, wdst, wos);
        CHECK(strstr(thrown_msg, "fewer bytes than expected") != NULL); thrown_msg[0] = 0;
        NS(freePinned)(env, NULL, wsrc); NS(freePinned)(env, NULL, wdst);
    }
    /* the flow of PinnedPool: buffers are reused across enumerations, and a second segment through the same buffers works */
    {
        for (int i = 0; i < n; i++) s[i] = (uint8_t)(i * 31 >> 3);
        jint got2 = NS(transform)(env, NULL, h, 2, src, n, cs, jkey, jaad, jivs, dst, sizes);
        CHECK(got2 == nch && thrown_msg[0] == 0 && ((jint*)sizes->data)[0] == cs + 28);
        uint64_t t2 = 0; for (int i = 0; i < nch; i++) t2 += ((jint*)sizes->data)[i];
        NS(detransform)(env, NULL, h, 2, dst, (jlong)t2, sizes, jkey, jaad, back, osz);
        CHECK(thrown_msg[0] == 0 && memcmp(back->data, s, n) == 0);
    }
    CHECK(live_copies == 0);                                                       /* every Get*Elements was released */
    NS(freePinned)(env, NULL, src); NS(freePinned)(env, NULL, dst); NS(freePinned)(env, NULL, back);
    NS(destroy)(env, NULL, h);
    printf("%s: %d checks, %d failures\n", failures ? "FAILED" : "OK", checks, failures);
    return failures ? 1 : 0;
}
