// ORT C-API shim (include/onnxruntime_c_api.h) over the native engine.  Lets the unchanged
// reference sources link against libgliclass_b200.so instead of onnxruntime:
//   CreateSession  (model.c:269)      -> glc_load
//   Run            (model.c:173-182)  -> glc_run
//   ReleaseSession (main.c:186)       -> glc_free
// Values are plain host tensors: inputs are non-owning views of caller memory (model.c:52-60),
// the output is an owned fp32 [B,C] buffer released by ReleaseValue (parallel_processor.c:88).
// Error convention: NULL OrtStatus* == success; otherwise a heap message.
// Threading: Run may be called concurrently on one session (main.c:141-149, CPU build); the
// engine serialises per device and round-robins concurrent small batches across devices.
#include <cstdlib>
#include <cstring>
#include <string>

#include "gliclass_b200.h"
#include "onnxruntime_c_api.h"

struct OrtStatus {
  OrtErrorCode code;
  std::string msg;
};
struct OrtEnv {
  int level;
  std::string logid;
};
struct OrtSessionOptions {
  int intra = 0, inter = 0, opt_level = 0, cuda_device = -1;
};
struct OrtSession {
  glc_model* model;
};
struct OrtRunOptions {
  int unused;
};
struct OrtMemoryInfo {
  int type, mem_type;
};
struct OrtValue {
  void* data;
  bool owned;
  ONNXTensorElementDataType type;
  int64_t dims[8];
  size_t ndim;
};
struct OrtTensorTypeAndShapeInfo {
  ONNXTensorElementDataType type;
  int64_t dims[8];
  size_t ndim;
};

namespace {

OrtStatus* mk(OrtErrorCode c, const std::string& m) { return new OrtStatus{c, m}; }

size_t elem_size(ONNXTensorElementDataType t) {
  switch (t) {
    case ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT: case ONNX_TENSOR_ELEMENT_DATA_TYPE_INT32: return 4;
    case ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64: return 8;
    case ONNX_TENSOR_ELEMENT_DATA_TYPE_UINT8: case ONNX_TENSOR_ELEMENT_DATA_TYPE_INT8: return 1;
    case ONNX_TENSOR_ELEMENT_DATA_TYPE_UINT16: case ONNX_TENSOR_ELEMENT_DATA_TYPE_INT16: return 2;
    default: return 0;
  }
}

OrtStatus* CreateStatus(OrtErrorCode code, const char* msg) { return mk(code, msg ? msg : ""); }
OrtErrorCode GetErrorCode(const OrtStatus* s) { return s ? s->code : ORT_OK; }
const char* GetErrorMessage(const OrtStatus* s) { return s ? s->msg.c_str() : ""; }
void ReleaseStatus(OrtStatus* s) { delete s; }

OrtStatus* CreateEnv(OrtLoggingLevel level, const char* logid, OrtEnv** out) {
  if (!out) return mk(ORT_INVALID_ARGUMENT, "CreateEnv: out is null");
  *out = new OrtEnv{(int)level, logid ? logid : ""};
  return nullptr;
}
void ReleaseEnv(OrtEnv* e) { delete e; }

OrtStatus* CreateSessionOptions(OrtSessionOptions** out) {
  if (!out) return mk(ORT_INVALID_ARGUMENT, "CreateSessionOptions: out is null");
  *out = new OrtSessionOptions();
  return nullptr;
}
void ReleaseSessionOptions(OrtSessionOptions* o) { delete o; }
OrtStatus* SetIntraOpNumThreads(OrtSessionOptions* o, int n) {
  if (!o) return mk(ORT_INVALID_ARGUMENT, "SetIntraOpNumThreads: null options");
  o->intra = n;   // advisory: the GPU engine has no intra-op thread pool
  return nullptr;
}
OrtStatus* SetInterOpNumThreads(OrtSessionOptions* o, int n) {
  if (!o) return mk(ORT_INVALID_ARGUMENT, "SetInterOpNumThreads: null options");
  o->inter = n;
  return nullptr;
}
OrtStatus* SetSessionGraphOptimizationLevel(OrtSessionOptions* o, GraphOptimizationLevel l) {
  if (!o) return mk(ORT_INVALID_ARGUMENT, "SetSessionGraphOptimizationLevel: null options");
  o->opt_level = (int)l;
  return nullptr;
}

OrtStatus* CreateSession(const OrtEnv* env, const ORTCHAR_T* path, const OrtSessionOptions* opts, OrtSession** out) {
  if (!env || !path || !out) return mk(ORT_INVALID_ARGUMENT, "CreateSession: null argument");
  glc_opts go;
  memset(&go, 0, sizeof(go));
  go.struct_size = sizeof(go);
  if (opts && opts->cuda_device >= 0 && !getenv("GLC_DEVICES")) {
    go.num_devices = 1;
    go.device_ids[0] = opts->cuda_device;
  }
  glc_model* m = glc_load(path, &go);
  if (!m) {
    std::string e = glc_last_error();
    OrtErrorCode c = e.find("cannot open") != std::string::npos ? ORT_NO_SUCHFILE
                     : e.find("onnx:") != std::string::npos     ? ORT_INVALID_PROTOBUF
                                                                 : ORT_FAIL;
    return mk(c, e);
  }
  *out = new OrtSession{m};
  return nullptr;
}
void ReleaseSession(OrtSession* s) {
  if (!s) return;
  glc_free(s->model);
  delete s;
}

void* AllocImpl(OrtAllocator*, size_t n) { return malloc(n); }
void FreeImpl(OrtAllocator*, void* p) { free(p); }
const OrtMemoryInfo* InfoImpl(const OrtAllocator*) {
  static OrtMemoryInfo cpu{(int)OrtDeviceAllocator, (int)OrtMemTypeDefault};
  return &cpu;
}
OrtAllocator g_alloc{ORT_API_VERSION, AllocImpl, FreeImpl, InfoImpl};

OrtStatus* GetAllocatorWithDefaultOptions(OrtAllocator** out) {
  if (!out) return mk(ORT_INVALID_ARGUMENT, "GetAllocatorWithDefaultOptions: out is null");
  *out = &g_alloc;
  return nullptr;
}
OrtStatus* AllocatorFree(OrtAllocator* a, void* p) {
  if (!a) return mk(ORT_INVALID_ARGUMENT, "AllocatorFree: null allocator");
  a->Free(a, p);
  return nullptr;
}

OrtStatus* dup_name(OrtAllocator* a, const char* s, char** value) {
  if (!a || !value) return mk(ORT_INVALID_ARGUMENT, "null allocator or value");
  size_t n = strlen(s) + 1;
  char* p = (char*)a->Alloc(a, n);
  if (!p) return mk(ORT_FAIL, "allocation failed");
  memcpy(p, s, n);
  *value = p;
  return nullptr;
}
OrtStatus* SessionGetInputCount(const OrtSession* s, size_t* out) {
  if (!s || !out) return mk(ORT_INVALID_ARGUMENT, "SessionGetInputCount: null argument");
  *out = 2;
  return nullptr;
}
OrtStatus* SessionGetOutputCount(const OrtSession* s, size_t* out) {
  if (!s || !out) return mk(ORT_INVALID_ARGUMENT, "SessionGetOutputCount: null argument");
  *out = 1;
  return nullptr;
}
OrtStatus* SessionGetInputName(const OrtSession* s, size_t i, OrtAllocator* a, char** v) {
  if (!s || i > 1) return mk(ORT_INVALID_ARGUMENT, "SessionGetInputName: index out of range");
  return dup_name(a, i == 0 ? "input_ids" : "attention_mask", v);
}
OrtStatus* SessionGetOutputName(const OrtSession* s, size_t i, OrtAllocator* a, char** v) {
  if (!s || i != 0) return mk(ORT_INVALID_ARGUMENT, "SessionGetOutputName: index out of range");
  return dup_name(a, "logits", v);
}

OrtStatus* CreateRunOptions(OrtRunOptions** out) {
  if (!out) return mk(ORT_INVALID_ARGUMENT, "CreateRunOptions: out is null");
  *out = new OrtRunOptions{0};
  return nullptr;
}
void ReleaseRunOptions(OrtRunOptions* o) { delete o; }

OrtStatus* CreateCpuMemoryInfo(OrtAllocatorType t, OrtMemType mt, OrtMemoryInfo** out) {
  if (!out) return mk(ORT_INVALID_ARGUMENT, "CreateCpuMemoryInfo: out is null");
  *out = new OrtMemoryInfo{(int)t, (int)mt};
  return nullptr;
}
void ReleaseMemoryInfo(OrtMemoryInfo* i) { delete i; }

OrtStatus* CreateTensorWithDataAsOrtValue(const OrtMemoryInfo* info, void* p, size_t len, const int64_t* shape, size_t nd,
                                          ONNXTensorElementDataType type, OrtValue** out) {
  if (!info || !out || (!shape && nd)) return mk(ORT_INVALID_ARGUMENT, "CreateTensorWithDataAsOrtValue: null argument");
  if (nd > 8) return mk(ORT_INVALID_ARGUMENT, "CreateTensorWithDataAsOrtValue: rank > 8");
  const size_t es = elem_size(type);
  if (!es) return mk(ORT_NOT_IMPLEMENTED, "CreateTensorWithDataAsOrtValue: unsupported element type");
  size_t n = 1;
  for (size_t i = 0; i < nd; ++i) {
    if (shape[i] < 0) return mk(ORT_INVALID_ARGUMENT, "CreateTensorWithDataAsOrtValue: negative dimension");
    n *= (size_t)shape[i];
  }
  if (n * es > len) return mk(ORT_INVALID_ARGUMENT, "CreateTensorWithDataAsOrtValue: buffer smaller than shape");
  if (!p && n) return mk(ORT_INVALID_ARGUMENT, "CreateTensorWithDataAsOrtValue: null data");
  OrtValue* v = new OrtValue();
  v->data = p;
  v->owned = false;   // ORT does not take ownership either (the reference leaks these buffers)
  v->type = type;
  v->ndim = nd;
  for (size_t i = 0; i < nd; ++i) v->dims[i] = shape[i];
  *out = v;
  return nullptr;
}
OrtStatus* CreateTensorAsOrtValue(OrtAllocator* a, const int64_t* shape, size_t nd, ONNXTensorElementDataType type,
                                  OrtValue** out) {
  if (!a || !out || (!shape && nd) || nd > 8) return mk(ORT_INVALID_ARGUMENT, "CreateTensorAsOrtValue: bad argument");
  const size_t es = elem_size(type);
  if (!es) return mk(ORT_NOT_IMPLEMENTED, "CreateTensorAsOrtValue: unsupported element type");
  size_t n = 1;
  for (size_t i = 0; i < nd; ++i) n *= (size_t)(shape[i] < 0 ? 0 : shape[i]);
  OrtValue* v = new OrtValue();
  v->data = calloc(n ? n : 1, es);
  v->owned = true;
  v->type = type;
  v->ndim = nd;
  for (size_t i = 0; i < nd; ++i) v->dims[i] = shape[i];
  *out = v;
  return nullptr;
}
void ReleaseValue(OrtValue* v) {
  if (!v) return;
  if (v->owned) free(v->data);
  delete v;
}
OrtStatus* IsTensor(const OrtValue* v, int* out) {
  if (!v || !out) return mk(ORT_INVALID_ARGUMENT, "IsTensor: null argument");
  *out = 1;
  return nullptr;
}
OrtStatus* GetTensorMutableData(OrtValue* v, void** out) {
  if (!v || !out) return mk(ORT_INVALID_ARGUMENT, "GetTensorMutableData: null argument");
  *out = v->data;
  return nullptr;
}
OrtStatus* GetTensorTypeAndShape(const OrtValue* v, OrtTensorTypeAndShapeInfo** out) {
  if (!v || !out) return mk(ORT_INVALID_ARGUMENT, "GetTensorTypeAndShape: null argument");
  OrtTensorTypeAndShapeInfo* i = new OrtTensorTypeAndShapeInfo();
  i->type = v->type;
  i->ndim = v->ndim;
  memcpy(i->dims, v->dims, sizeof(v->dims));
  *out = i;
  return nullptr;
}
void ReleaseTensorTypeAndShapeInfo(OrtTensorTypeAndShapeInfo* i) { delete i; }
OrtStatus* GetTensorElementType(const OrtTensorTypeAndShapeInfo* i, ONNXTensorElementDataType* out) {
  if (!i || !out) return mk(ORT_INVALID_ARGUMENT, "GetTensorElementType: null argument");
  *out = i->type;
  return nullptr;
}
OrtStatus* GetDimensionsCount(const OrtTensorTypeAndShapeInfo* i, size_t* out) {
  if (!i || !out) return mk(ORT_INVALID_ARGUMENT, "GetDimensionsCount: null argument");
  *out = i->ndim;
  return nullptr;
}
OrtStatus* GetDimensions(const OrtTensorTypeAndShapeInfo* i, int64_t* d, size_t n) {
  if (!i || (!d && n)) return mk(ORT_INVALID_ARGUMENT, "GetDimensions: null argument");
  for (size_t k = 0; k < n && k < i->ndim; ++k) d[k] = i->dims[k];
  return nullptr;
}
OrtStatus* GetTensorShapeElementCount(const OrtTensorTypeAndShapeInfo* i, size_t* out) {
  if (!i || !out) return mk(ORT_INVALID_ARGUMENT, "GetTensorShapeElementCount: null argument");
  size_t n = 1;
  for (size_t k = 0; k < i->ndim; ++k) n *= (size_t)i->dims[k];
  *out = n;
  return nullptr;
}

// THE hot path.  Inputs are matched by name like ORT does, so callers may pass them in any order.
OrtStatus* Run(OrtSession* s, const OrtRunOptions*, const char* const* in_names, const OrtValue* const* in, size_t n_in,
               const char* const* out_names, size_t n_out, OrtValue** out) {
  if (!s || !in_names || !in || !out_names || !out) return mk(ORT_INVALID_ARGUMENT, "Run: null argument");
  if (n_out != 1 || !out_names[0] || strcmp(out_names[0], "logits") != 0)
    return mk(ORT_INVALID_ARGUMENT, "Run: the model has exactly one output named 'logits'");
  const OrtValue *ids = nullptr, *mask = nullptr;
  for (size_t i = 0; i < n_in; ++i) {
    if (!in_names[i] || !in[i]) return mk(ORT_INVALID_ARGUMENT, "Run: null input");
    if (!strcmp(in_names[i], "input_ids")) ids = in[i];
    else if (!strcmp(in_names[i], "attention_mask")) mask = in[i];
    else return mk(ORT_INVALID_ARGUMENT, std::string("Run: Invalid input name: ") + in_names[i]);
  }
  if (!ids || !mask) return mk(ORT_INVALID_ARGUMENT, "Run: missing input (need input_ids and attention_mask)");
  if (ids->type != ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64 || mask->type != ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64)
    return mk(ORT_INVALID_ARGUMENT, "Run: inputs must be tensor(int64)");
  if (ids->ndim != 2 || mask->ndim != 2 || ids->dims[0] != mask->dims[0] || ids->dims[1] != mask->dims[1])
    return mk(ORT_INVALID_ARGUMENT, "Run: inputs must both be [batch_size, sequence_length]");
  const int B = (int)ids->dims[0], S = (int)ids->dims[1];
  const int C = glc_num_classes(s->model, (const int64_t*)ids->data, B, S);
  if (C < 0) return mk(ORT_FAIL, glc_last_error());
  OrtValue* v = new OrtValue();
  v->type = ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT;
  v->ndim = 2;
  v->dims[0] = B;
  v->dims[1] = C;
  v->owned = true;
  const size_t n = (size_t)B * C;
  v->data = malloc((n ? n : 1) * sizeof(float));
  int cc = 0;
  int rc = glc_run(s->model, (const int64_t*)ids->data, (const int64_t*)mask->data, B, S, (float*)v->data, n, &cc);
  if (rc != GLC_OK) {
    OrtStatus* st = mk(ORT_FAIL, glc_last_error());
    ReleaseValue(v);
    return st;
  }
  *out = v;
  return nullptr;
}

const OrtApi g_api = {
    CreateStatus, GetErrorCode, GetErrorMessage, ReleaseStatus,
    CreateEnv, ReleaseEnv,
    CreateSessionOptions, ReleaseSessionOptions, SetIntraOpNumThreads, SetInterOpNumThreads, SetSessionGraphOptimizationLevel,
    CreateSession, ReleaseSession, SessionGetInputCount, SessionGetOutputCount, SessionGetInputName, SessionGetOutputName,
    CreateRunOptions, ReleaseRunOptions, Run,
    GetAllocatorWithDefaultOptions, AllocatorFree, CreateCpuMemoryInfo, ReleaseMemoryInfo,
    CreateTensorWithDataAsOrtValue, CreateTensorAsOrtValue, ReleaseValue, IsTensor, GetTensorMutableData,
    GetTensorTypeAndShape, ReleaseTensorTypeAndShapeInfo, GetTensorElementType, GetDimensionsCount, GetDimensions,
    GetTensorShapeElementCount,
};

const OrtApi* GetApi(uint32_t version) {
  // like ORT: a caller built against a newer API than the library gets NULL
  return (version >= 1 && version <= ORT_API_VERSION) ? &g_api : nullptr;
}
const char* GetVersionString() { return "1.19.2-gliclass-b200"; }
const OrtApiBase g_base = {GetApi, GetVersionString};

}  // namespace

extern "C" {
const OrtApiBase* OrtGetApiBase(void) { return &g_base; }

OrtStatus* OrtSessionOptionsAppendExecutionProvider_CUDA(OrtSessionOptions* o, int device_id) {
  if (!o) return mk(ORT_INVALID_ARGUMENT, "OrtSessionOptionsAppendExecutionProvider_CUDA: null options");
  if (device_id < 0) return mk(ORT_INVALID_ARGUMENT, "OrtSessionOptionsAppendExecutionProvider_CUDA: negative device id");
  o->cuda_device = device_id;
  return nullptr;
}
}
