/* shim_e2e — times the hot path through the ORT-named entry points exactly the way the reference's src/model.c drives
 * them ("end-to-end number through the ORT-shim Run incl. H2D/D2H"):
 *     flatten_int_array  (model.c:17-29)   malloc'd int64 [B*S] copies of the token ids / mask      -> PAGEABLE host memory
 *     create_tensor      (model.c:39-71)   CreateCpuMemoryInfo + CreateTensorWithDataAsOrtValue
 *     run_inference      (model.c:122-207) CreateRunOptions, SessionGetOutputName, Run, ReleaseRunOptions
 *     postprocessor.c:32-84                GetTensorTypeAndShape / GetDimensions / GetTensorMutableData
 * Built as lib/libglc_shim_e2e.so next to the engine library; bench.py and the tests call it through ctypes.  It links
 * ONLY against the public shim API (include/onnxruntime_c_api.h) — the same symbols the unchanged reference binary uses. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "onnxruntime_c_api.h"

#define SHIM_EXPORT __attribute__((visibility("default")))

typedef struct {
  const OrtApi* ort;
  OrtEnv* env;
  OrtSessionOptions* opts;
  OrtSession* session;
  char err[512];
} shim_ctx;

static int fail(shim_ctx* c, OrtStatus* st, const char* what) {
  snprintf(c->err, sizeof(c->err), "%s: %s", what, st ? c->ort->GetErrorMessage(st) : "?");
  if (st) c->ort->ReleaseStatus(st);
  return -1;
}

SHIM_EXPORT const char* shim_e2e_error(void* h) { return h ? ((shim_ctx*)h)->err : "null handle"; }

/* initialize_ort_api / initialize_ort_environment / create_ort_session (model.c:217-310) */
SHIM_EXPORT void* shim_e2e_open(const char* model_path, int num_threads) {
  shim_ctx* c = (shim_ctx*)calloc(1, sizeof(shim_ctx));
  if (!c) return NULL;
  c->ort = OrtGetApiBase()->GetApi(ORT_API_VERSION);
  OrtStatus* st = NULL;
  if (!c->ort) { snprintf(c->err, sizeof(c->err), "GetApi failed"); return c; }
  if ((st = c->ort->CreateEnv(ORT_LOGGING_LEVEL_WARNING, "GLiClass", &c->env))) { fail(c, st, "CreateEnv"); return c; }
  if ((st = c->ort->CreateSessionOptions(&c->opts))) { fail(c, st, "CreateSessionOptions"); return c; }
  c->ort->SetIntraOpNumThreads(c->opts, num_threads);
  c->ort->SetInterOpNumThreads(c->opts, num_threads);
  c->ort->SetSessionGraphOptimizationLevel(c->opts, ORT_ENABLE_ALL);
  if ((st = c->ort->CreateSession(c->env, model_path, c->opts, &c->session))) { fail(c, st, "CreateSession"); return c; }
  return c;
}

SHIM_EXPORT int shim_e2e_ok(void* h) { return h && ((shim_ctx*)h)->session != NULL; }

SHIM_EXPORT void shim_e2e_close(void* h) {
  shim_ctx* c = (shim_ctx*)h;
  if (!c) return;
  if (c->session) c->ort->ReleaseSession(c->session);
  if (c->opts) c->ort->ReleaseSessionOptions(c->opts);
  if (c->env) c->ort->ReleaseEnv(c->env);
  free(c);
}

static OrtValue* make_tensor(shim_ctx* c, int64_t* data, int64_t B, int64_t S) {   /* model.c:39-71 */
  OrtMemoryInfo* mi = NULL;
  OrtStatus* st = c->ort->CreateCpuMemoryInfo(OrtArenaAllocator, OrtMemTypeDefault, &mi);
  if (st) { fail(c, st, "CreateCpuMemoryInfo"); return NULL; }
  int64_t dims[2] = {B, S};
  OrtValue* v = NULL;
  st = c->ort->CreateTensorWithDataAsOrtValue(mi, data, (size_t)(B * S) * sizeof(int64_t), dims, 2,
                                              ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64, &v);
  c->ort->ReleaseMemoryInfo(mi);
  if (st) { fail(c, st, "CreateTensorWithDataAsOrtValue"); return NULL; }
  return v;
}

/* one run_inference call as the reference makes it; logits_out fp32 [B, *C_out] (capacity in floats) */
SHIM_EXPORT int shim_e2e_run(void* h, const int64_t* ids, const int64_t* mask, int B, int S, float* logits_out, size_t capacity,
                             int* C_out) {
  shim_ctx* c = (shim_ctx*)h;
  if (!c || !c->session) return -1;
  const size_t n = (size_t)B * (size_t)S;
  int64_t* fi = (int64_t*)malloc(n * sizeof(int64_t));   /* flatten_int_array: pageable malloc'd copies */
  int64_t* fm = (int64_t*)malloc(n * sizeof(int64_t));
  if (!fi || !fm) { free(fi); free(fm); snprintf(c->err, sizeof(c->err), "malloc"); return -1; }
  memcpy(fi, ids, n * sizeof(int64_t));
  memcpy(fm, mask, n * sizeof(int64_t));
  int rc = -1;
  OrtValue* in[2] = {make_tensor(c, fi, B, S), NULL};
  OrtValue* out = NULL;
  OrtRunOptions* ro = NULL;
  char* out_name = NULL;
  OrtAllocator* alloc = NULL;
  OrtStatus* st = NULL;
  if (!in[0]) goto done;
  in[1] = make_tensor(c, fm, B, S);
  if (!in[1]) goto done;
  if ((st = c->ort->CreateRunOptions(&ro))) { fail(c, st, "CreateRunOptions"); goto done; }
  if ((st = c->ort->GetAllocatorWithDefaultOptions(&alloc))) { fail(c, st, "GetAllocatorWithDefaultOptions"); goto done; }
  if ((st = c->ort->SessionGetOutputName(c->session, 0, alloc, &out_name))) { fail(c, st, "SessionGetOutputName"); goto done; }
  {
    const char* in_names[2] = {"input_ids", "attention_mask"};
    const char* out_names[1] = {out_name};
    if ((st = c->ort->Run(c->session, ro, in_names, (const OrtValue* const*)in, 2, out_names, 1, &out))) {
      fail(c, st, "Run");
      goto done;
    }
  }
  {
    OrtTensorTypeAndShapeInfo* info = NULL;
    size_t nd = 0;
    int64_t dims[4] = {0, 0, 0, 0};
    void* data = NULL;
    if ((st = c->ort->GetTensorTypeAndShape(out, &info))) { fail(c, st, "GetTensorTypeAndShape"); goto done; }
    c->ort->GetDimensionsCount(info, &nd);
    if (nd != 2) { c->ort->ReleaseTensorTypeAndShapeInfo(info); snprintf(c->err, sizeof(c->err), "logits rank %zu", nd); goto done; }
    c->ort->GetDimensions(info, dims, nd);
    c->ort->ReleaseTensorTypeAndShapeInfo(info);
    if ((st = c->ort->GetTensorMutableData(out, &data))) { fail(c, st, "GetTensorMutableData"); goto done; }
    if (C_out) *C_out = (int)dims[1];
    if ((size_t)(dims[0] * dims[1]) > capacity) { snprintf(c->err, sizeof(c->err), "logits buffer too small"); goto done; }
    if (logits_out && dims[0] * dims[1] > 0) memcpy(logits_out, data, (size_t)(dims[0] * dims[1]) * sizeof(float));
    rc = 0;
  }
done:
  if (out_name && alloc) alloc->Free(alloc, out_name);
  if (ro) c->ort->ReleaseRunOptions(ro);
  if (out) c->ort->ReleaseValue(out);
  if (in[0]) c->ort->ReleaseValue(in[0]);
  if (in[1]) c->ort->ReleaseValue(in[1]);
  free(fi);
  free(fm);
  return rc;
}

/* `steps` back-to-back run_inference calls; wall seconds through *seconds (CLOCK_MONOTONIC around the whole loop) */
SHIM_EXPORT int shim_e2e_time(void* h, const int64_t* ids, const int64_t* mask, int B, int S, float* logits_out, size_t capacity,
                              int steps, double* seconds) {
  struct timespec t0, t1;
  int C = 0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int k = 0; k < steps; ++k)
    if (shim_e2e_run(h, ids, mask, B, S, logits_out, capacity, &C) != 0) return -1;
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  return 0;
}
