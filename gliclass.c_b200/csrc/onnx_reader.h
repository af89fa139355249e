// Minimal ONNX (protobuf wire format) reader: just enough of ModelProto/GraphProto/NodeProto/
// TensorProto to harvest the initializers, node names/scopes and scalar Constants that the
// GLiClass engine needs.  This replaces the model-parsing half of ORT's CreateSession
// (reference src/model.c:269) for one model family; it is NOT a generic ONNX executor
// (SURVEY.md App. C explains why a generic executor is the wrong design here).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace glc {

struct OnnxTensor {
  std::string name;
  std::vector<int64_t> dims;
  int data_type = 0;            // 1 = float, 7 = int64, 6 = int32, 9 = bool, 11 = double
  const uint8_t* raw = nullptr; // points into the file buffer (raw_data) — little endian
  size_t raw_bytes = 0;
  std::vector<float> float_data;   // when stored as repeated float
  std::vector<int64_t> int64_data; // when stored as repeated int64
  // element count; -1 on a negative dimension or int64 overflow (crafted files)
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : dims) {
      if (d < 0 || (d > 0 && n > INT64_MAX / d)) return -1;
      n *= d;
    }
    return n;
  }
  bool external = false;        // data_location = EXTERNAL: raw points into OnnxGraph::ext_files after load()
  std::string ext_location;
  uint64_t ext_offset = 0, ext_length = 0;
};

struct OnnxAttr {
  std::string name;
  int64_t i = 0;
  float f = 0.f;
  bool has_t = false;
  OnnxTensor t;
  std::vector<int64_t> ints;
};

struct OnnxNode {
  std::string name, op_type;
  std::vector<std::string> inputs, outputs;
  std::vector<OnnxAttr> attrs;
  const OnnxAttr* attr(const char* n) const {
    for (auto& a : attrs) if (a.name == n) return &a;
    return nullptr;
  }
};

struct OnnxGraph {
  std::vector<uint8_t> file;   // owns the bytes; tensors point into it
  std::unordered_map<std::string, std::vector<uint8_t>> ext_files;   // external-data sidecars, keyed by relative location
  int64_t ir_version = 0;
  int64_t opset = 0;
  std::string producer;
  std::vector<OnnxNode> nodes;
  std::vector<OnnxTensor> initializers;
  std::vector<std::string> input_names, output_names;
  std::unordered_map<std::string, int> init_by_name;
  std::unordered_map<std::string, int> producer_of;   // value name -> node index

  // throws std::runtime_error on malformed input
  void load(const std::string& path);
  // Follow Identity chains to an initializer (torch's exporter dedups identical tensors and
  // leaves Identity(src)->name aliases).  Returns nullptr when `value` is not a constant.
  const OnnxTensor* resolve(const std::string& value) const;
  // Constant node / initializer scalar helpers
  bool scalar_int(const std::string& value, int64_t* out) const;
  bool scalar_float(const std::string& value, float* out) const;
};

// Copies a float tensor (raw or float_data) into `out` (numel floats).
void tensor_to_float(const OnnxTensor& t, float* out);

}  // namespace glc
