/* Host-only exercise of the ORT shim, written the way the reference's model.c / postprocessor.c
 * use the API (model.c:39-71, 122-207; postprocessor.c:32-84).  No GPU: CreateSession must fail
 * with a message instead of crashing or falling back. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "onnxruntime_c_api.h"

const OrtApi* g_ort = NULL;
#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
  g_ort = OrtGetApiBase()->GetApi(ORT_API_VERSION);
  CHECK(g_ort != NULL);
  CHECK(OrtGetApiBase()->GetApi(ORT_API_VERSION + 1) == NULL);
  OrtEnv* env = NULL;
  CHECK(g_ort->CreateEnv(ORT_LOGGING_LEVEL_WARNING, "GLiClass", &env) == NULL && env);

  /* create_tensor (model.c:39-71) */
  OrtMemoryInfo* mi = NULL;
  CHECK(g_ort->CreateCpuMemoryInfo(OrtArenaAllocator, OrtMemTypeDefault, &mi) == NULL);
  int64_t data[6] = {1, 5, 1025, 7, 1026, 2};
  int64_t dims[2] = {2, 3};
  OrtValue* t = NULL;
  CHECK(g_ort->CreateTensorWithDataAsOrtValue(mi, data, sizeof(data), dims, 2, ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64, &t) == NULL);
  g_ort->ReleaseMemoryInfo(mi);

  /* postprocessor.c:39-84 shape plumbing */
  OrtTensorTypeAndShapeInfo* info = NULL;
  CHECK(g_ort->GetTensorTypeAndShape(t, &info) == NULL);
  size_t nd = 0;
  CHECK(g_ort->GetDimensionsCount(info, &nd) == NULL && nd == 2);
  int64_t got[2] = {0, 0};
  CHECK(g_ort->GetDimensions(info, got, nd) == NULL && got[0] == 2 && got[1] == 3);
  void* p = NULL;
  CHECK(g_ort->GetTensorMutableData(t, &p) == NULL && p == (void*)data);   /* zero copy, not owned */
  g_ort->ReleaseTensorTypeAndShapeInfo(info);

  /* too-small buffer is rejected with a message */
  OrtValue* bad = NULL;
  OrtMemoryInfo* mi2 = NULL;
  g_ort->CreateCpuMemoryInfo(OrtArenaAllocator, OrtMemTypeDefault, &mi2);
  OrtStatus* st = g_ort->CreateTensorWithDataAsOrtValue(mi2, data, 8, dims, 2, ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64, &bad);
  CHECK(st != NULL && strlen(g_ort->GetErrorMessage(st)) > 0 && bad == NULL);
  g_ort->ReleaseStatus(st);
  g_ort->ReleaseMemoryInfo(mi2);

  /* allocator + names (model.c:139-186) */
  OrtAllocator* alloc = NULL;
  CHECK(g_ort->GetAllocatorWithDefaultOptions(&alloc) == NULL && alloc && alloc->Free);

  /* create_ort_session (model.c:217-281) */
  OrtSessionOptions* so = NULL;
  CHECK(g_ort->CreateSessionOptions(&so) == NULL);
  CHECK(g_ort->SetIntraOpNumThreads(so, 8) == NULL && g_ort->SetInterOpNumThreads(so, 8) == NULL);
  CHECK(g_ort->SetSessionGraphOptimizationLevel(so, ORT_ENABLE_ALL) == NULL);
  CHECK(OrtSessionOptionsAppendExecutionProvider_CUDA(so, 0) == NULL);
  OrtSession* sess = NULL;
  st = g_ort->CreateSession(env, "/nonexistent/model.onnx", so, &sess);
  CHECK(st != NULL && sess == NULL);
  printf("missing-file: %s\n", g_ort->GetErrorMessage(st));
  g_ort->ReleaseStatus(st);
  if (argc > 1) {
    st = g_ort->CreateSession(env, argv[1], so, &sess);
    if (st) {   /* expected on a box without a B200 */
      printf("no-gpu: %s\n", g_ort->GetErrorMessage(st));
      CHECK(strstr(g_ort->GetErrorMessage(st), "no CPU fallback") != NULL || strstr(g_ort->GetErrorMessage(st), "CUDA") != NULL);
      g_ort->ReleaseStatus(st);
    } else {
      size_t n = 0;
      CHECK(g_ort->SessionGetOutputCount(sess, &n) == NULL && n == 1);
      char* name = NULL;
      CHECK(g_ort->SessionGetOutputName(sess, 0, alloc, &name) == NULL && !strcmp(name, "logits"));
      alloc->Free(alloc, name);
      g_ort->ReleaseSession(sess);
      printf("session-ok\n");
    }
  }
  g_ort->ReleaseSessionOptions(so);
  g_ort->ReleaseValue(t);
  g_ort->ReleaseEnv(env);
  printf("OK\n");
  return 0;
}
