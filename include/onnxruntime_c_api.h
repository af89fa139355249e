/* onnxruntime_c_api.h — source-compatible SUBSET of the ONNX Runtime C API, implemented by
 * libgliclass_b200.so (csrc/ort_shim.cc) on top of the native engine (gliclass_b200.h).
 *
 * Purpose: the unchanged reference sources (main.c, src/model.c, src/postprocessor.c,
 * src/parallel_processor.c) #include "onnxruntime_c_api.h" and call g_ort->... directly
 * (SURVEY.md §8b, App. D).  Compiling them against THIS header and linking against
 * libgliclass_b200.so swaps ORT for the B200 engine without touching a line of the reference.
 * It is source compatibility for the members listed below — not binary compatibility with the
 * real ORT vtable (whose slots are ordered differently and number in the hundreds).
 *
 * Each member cites the reference call site it serves.
 */
#ifndef GLICLASS_B200_ORT_SHIM_H
#define GLICLASS_B200_ORT_SHIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORT_API_VERSION 19            /* the reference pins onnxruntime 1.19.2 (CMakeLists.txt:22-24) */
#define ORT_GLICLASS_B200_SHIM 1

#ifndef ORT_EXPORT
#define ORT_EXPORT __attribute__((visibility("default")))
#endif
#define ORT_API_CALL
#define ORTCHAR_T char

typedef struct OrtEnv OrtEnv;
typedef struct OrtStatus OrtStatus;
typedef struct OrtSession OrtSession;
typedef struct OrtSessionOptions OrtSessionOptions;
typedef struct OrtRunOptions OrtRunOptions;
typedef struct OrtValue OrtValue;
typedef struct OrtMemoryInfo OrtMemoryInfo;
typedef struct OrtTensorTypeAndShapeInfo OrtTensorTypeAndShapeInfo;
typedef OrtStatus* OrtStatusPtr;

typedef enum OrtLoggingLevel {
  ORT_LOGGING_LEVEL_VERBOSE,
  ORT_LOGGING_LEVEL_INFO,
  ORT_LOGGING_LEVEL_WARNING,
  ORT_LOGGING_LEVEL_ERROR,
  ORT_LOGGING_LEVEL_FATAL
} OrtLoggingLevel;

typedef enum OrtErrorCode {
  ORT_OK, ORT_FAIL, ORT_INVALID_ARGUMENT, ORT_NO_SUCHFILE, ORT_NO_MODEL, ORT_ENGINE_ERROR,
  ORT_RUNTIME_EXCEPTION, ORT_INVALID_PROTOBUF, ORT_MODEL_LOADED, ORT_NOT_IMPLEMENTED,
  ORT_INVALID_GRAPH, ORT_EP_FAIL
} OrtErrorCode;

typedef enum ONNXTensorElementDataType {
  ONNX_TENSOR_ELEMENT_DATA_TYPE_UNDEFINED = 0,
  ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT = 1,
  ONNX_TENSOR_ELEMENT_DATA_TYPE_UINT8 = 2,
  ONNX_TENSOR_ELEMENT_DATA_TYPE_INT8 = 3,
  ONNX_TENSOR_ELEMENT_DATA_TYPE_UINT16 = 4,
  ONNX_TENSOR_ELEMENT_DATA_TYPE_INT16 = 5,
  ONNX_TENSOR_ELEMENT_DATA_TYPE_INT32 = 6,
  ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64 = 7
} ONNXTensorElementDataType;

typedef enum OrtAllocatorType { OrtInvalidAllocator = -1, OrtDeviceAllocator = 0, OrtArenaAllocator = 1 } OrtAllocatorType;
typedef enum OrtMemType { OrtMemTypeCPUInput = -2, OrtMemTypeCPUOutput = -1, OrtMemTypeCPU = -1, OrtMemTypeDefault = 0 } OrtMemType;
typedef enum GraphOptimizationLevel {
  ORT_DISABLE_ALL = 0, ORT_ENABLE_BASIC = 1, ORT_ENABLE_EXTENDED = 2, ORT_ENABLE_ALL = 99
} GraphOptimizationLevel;

/* model.c:139,186 — allocator->Free(allocator, output_name) */
typedef struct OrtAllocator {
  uint32_t version;
  void* (*Alloc)(struct OrtAllocator* this_, size_t size);
  void (*Free)(struct OrtAllocator* this_, void* p);
  const OrtMemoryInfo* (*Info)(const struct OrtAllocator* this_);
} OrtAllocator;

typedef struct OrtApi {
  /* status — model.c:42-47 pattern; postprocessor.c:42,155 */
  OrtStatus* (*CreateStatus)(OrtErrorCode code, const char* msg);
  OrtErrorCode (*GetErrorCode)(const OrtStatus* status);
  const char* (*GetErrorMessage)(const OrtStatus* status);
  void (*ReleaseStatus)(OrtStatus* status);

  /* environment — model.c:290; main.c:96,187 */
  OrtStatus* (*CreateEnv)(OrtLoggingLevel level, const char* logid, OrtEnv** out);
  void (*ReleaseEnv)(OrtEnv* env);

  /* session options — model.c:223,232,242,261,278 */
  OrtStatus* (*CreateSessionOptions)(OrtSessionOptions** out);
  void (*ReleaseSessionOptions)(OrtSessionOptions* opts);
  OrtStatus* (*SetIntraOpNumThreads)(OrtSessionOptions* opts, int n);
  OrtStatus* (*SetInterOpNumThreads)(OrtSessionOptions* opts, int n);
  OrtStatus* (*SetSessionGraphOptimizationLevel)(OrtSessionOptions* opts, GraphOptimizationLevel level);

  /* session — model.c:269 (load model.onnx -> GPU weights); main.c:186 */
  OrtStatus* (*CreateSession)(const OrtEnv* env, const ORTCHAR_T* model_path, const OrtSessionOptions* opts,
                              OrtSession** out);
  void (*ReleaseSession)(OrtSession* session);
  OrtStatus* (*SessionGetInputCount)(const OrtSession* session, size_t* out);
  OrtStatus* (*SessionGetOutputCount)(const OrtSession* session, size_t* out);               /* model.c:150 */
  OrtStatus* (*SessionGetInputName)(const OrtSession* session, size_t index, OrtAllocator* allocator, char** value);
  OrtStatus* (*SessionGetOutputName)(const OrtSession* session, size_t index, OrtAllocator* allocator,
                                     char** value);                                            /* model.c:159 */

  /* run — model.c:130,173-182,190: THE hot path */
  OrtStatus* (*CreateRunOptions)(OrtRunOptions** out);
  void (*ReleaseRunOptions)(OrtRunOptions* opts);
  OrtStatus* (*Run)(OrtSession* session, const OrtRunOptions* run_options, const char* const* input_names,
                    const OrtValue* const* inputs, size_t input_len, const char* const* output_names,
                    size_t output_names_len, OrtValue** outputs);

  /* allocator / memory info — model.c:41,61,139 */
  OrtStatus* (*GetAllocatorWithDefaultOptions)(OrtAllocator** out);
  OrtStatus* (*AllocatorFree)(OrtAllocator* allocator, void* p);
  OrtStatus* (*CreateCpuMemoryInfo)(OrtAllocatorType type, OrtMemType mem_type, OrtMemoryInfo** out);
  void (*ReleaseMemoryInfo)(OrtMemoryInfo* info);

  /* values — model.c:52-60,97; main.c:174-175; parallel_processor.c:88; postprocessor.c:39-75,154 */
  OrtStatus* (*CreateTensorWithDataAsOrtValue)(const OrtMemoryInfo* info, void* p_data, size_t p_data_len,
                                               const int64_t* shape, size_t shape_len,
                                               ONNXTensorElementDataType type, OrtValue** out);
  OrtStatus* (*CreateTensorAsOrtValue)(OrtAllocator* allocator, const int64_t* shape, size_t shape_len,
                                       ONNXTensorElementDataType type, OrtValue** out);
  void (*ReleaseValue)(OrtValue* value);
  OrtStatus* (*IsTensor)(const OrtValue* value, int* out);
  OrtStatus* (*GetTensorMutableData)(OrtValue* value, void** out);
  OrtStatus* (*GetTensorTypeAndShape)(const OrtValue* value, OrtTensorTypeAndShapeInfo** out);
  void (*ReleaseTensorTypeAndShapeInfo)(OrtTensorTypeAndShapeInfo* info);
  OrtStatus* (*GetTensorElementType)(const OrtTensorTypeAndShapeInfo* info, ONNXTensorElementDataType* out);
  OrtStatus* (*GetDimensionsCount)(const OrtTensorTypeAndShapeInfo* info, size_t* out);
  OrtStatus* (*GetDimensions)(const OrtTensorTypeAndShapeInfo* info, int64_t* dim_values, size_t dim_values_length);
  OrtStatus* (*GetTensorShapeElementCount)(const OrtTensorTypeAndShapeInfo* info, size_t* out);
} OrtApi;

typedef struct OrtApiBase {
  const OrtApi* (*GetApi)(uint32_t version);       /* model.c:304 */
  const char* (*GetVersionString)(void);
} OrtApiBase;

ORT_EXPORT const OrtApiBase* OrtGetApiBase(void);

/* model.c:253 (only under -DUSE_CUDA): records the device; the engine always runs on GPU. */
ORT_EXPORT OrtStatus* OrtSessionOptionsAppendExecutionProvider_CUDA(OrtSessionOptions* options, int device_id);

#ifdef __cplusplus
}
#endif
#endif /* GLICLASS_B200_ORT_SHIM_H */
