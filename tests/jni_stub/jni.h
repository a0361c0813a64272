// Minimal stand-in for the JDK's <jni.h>, ONLY to type-check
// caffeonspark_b200/csrc/jni_shim.cpp in a container without a JDK
// (tests/test_host_abi.py::test_jni_shim_type_checks).  Not an ABI-accurate
// JNI header: never link against it.
#ifndef COS_TEST_JNI_STUB_H_
#define COS_TEST_JNI_STUB_H_
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef unsigned char jboolean;
typedef jint jsize;
class _jobject {};
typedef _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jobject jobjectArray;
struct _jfieldID;
typedef _jfieldID* jfieldID;
struct _jmethodID;
typedef _jmethodID* jmethodID;
#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
struct JNIEnv {
  jclass FindClass(const char*);
  jint ThrowNew(jclass, const char*);
  jclass GetObjectClass(jobject);
  jfieldID GetFieldID(jclass, const char*, const char*);
  jlong GetLongField(jobject, jfieldID);
  jmethodID GetMethodID(jclass, const char*, const char*);
  void CallVoidMethod(jobject, jmethodID, ...);
  jint CallIntMethod(jobject, jmethodID, ...);
  jobject CallObjectMethod(jobject, jmethodID, ...);
  jboolean ExceptionCheck();
  const char* GetStringUTFChars(jstring, jboolean*);
  void ReleaseStringUTFChars(jstring, const char*);
  jobjectArray NewObjectArray(jsize, jclass, jobject);
  jstring NewStringUTF(const char*);
  void SetObjectArrayElement(jobjectArray, jsize, jobject);
  void DeleteLocalRef(jobject);
  jsize GetArrayLength(jarray);
  jobject GetObjectArrayElement(jobjectArray, jsize);
};
#endif
