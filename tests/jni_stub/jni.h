/*
 * TEST-ONLY stand-in for the JDK's <jni.h>: just the types, constants and JNIEnv entries jni/tsgpu_jni.c uses, with the
 * signatures the JNI specification gives them.  There is no JDK in the build image; this header lets the glue be compiled
 * and driven by tests/cpp/test_jni_shim.c against a fake JNIEnv.  A real build uses $JAVA_HOME/include/jni.h.
 */
#ifndef TSGPU_TEST_JNI_STUB_H
#define TSGPU_TEST_JNI_STUB_H
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_COMMIT 1
#define JNI_ABORT 2

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef jint jsize;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jbyteArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv* env, const char* name);
    jint (*ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
    jsize (*GetArrayLength)(JNIEnv* env, jarray array);
    jbyte* (*GetByteArrayElements)(JNIEnv* env, jbyteArray array, jboolean* isCopy);
    void (*ReleaseByteArrayElements)(JNIEnv* env, jbyteArray array, jbyte* elems, jint mode);
    jint* (*GetIntArrayElements)(JNIEnv* env, jintArray array, jboolean* isCopy);
    void (*ReleaseIntArrayElements)(JNIEnv* env, jintArray array, jint* elems, jint mode);
    jobject (*NewDirectByteBuffer)(JNIEnv* env, void* address, jlong capacity);
    void* (*GetDirectBufferAddress)(JNIEnv* env, jobject buf);
    jlong (*GetDirectBufferCapacity)(JNIEnv* env, jobject buf);
};
#endif
