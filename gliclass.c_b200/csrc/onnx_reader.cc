// See onnx_reader.h.  Field numbers are the standard onnx.proto3 ones:
//   ModelProto{ir_version=1, producer_name=2, graph=7, opset_import=8{domain=1, version=2}}
//   GraphProto{node=1, name=2, initializer=5, input=11, output=12}
//   NodeProto{input=1, output=2, name=3, op_type=4, attribute=5}
//   AttributeProto{name=1, f=2, i=3, s=4, t=5, ints=8, type=20}
//   TensorProto{dims=1, data_type=2, float_data=4, int64_data=7, name=8, raw_data=9,
//               external_data=13, data_location=14}
//   ValueInfoProto{name=1}
#include "onnx_reader.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace glc {
namespace {

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (true) {
      if (p >= end) throw std::runtime_error("onnx: truncated varint");
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) break;
      shift += 7;
      if (shift > 63) throw std::runtime_error("onnx: varint too long");
    }
    return v;
  }
  // returns field number, sets wire type
  uint32_t tag(int* wt) {
    uint64_t t = varint();
    *wt = (int)(t & 7);
    return (uint32_t)(t >> 3);
  }
  Cursor sub() {
    uint64_t n = varint();
    if ((uint64_t)(end - p) < n) throw std::runtime_error("onnx: truncated length-delimited field");
    Cursor c{p, p + n};
    p += n;
    return c;
  }
  std::string str() {
    Cursor c = sub();
    return std::string((const char*)c.p, (size_t)(c.end - c.p));
  }
  void skip(int wt) {
    switch (wt) {
      case 0: varint(); break;
      case 1: if (end - p < 8) throw std::runtime_error("onnx: truncated fixed64"); p += 8; break;
      case 2: sub(); break;
      case 5: if (end - p < 4) throw std::runtime_error("onnx: truncated fixed32"); p += 4; break;
      default: throw std::runtime_error("onnx: unsupported wire type");
    }
  }
};

void parse_tensor(Cursor c, OnnxTensor* t) {
  while (!c.done()) {
    int wt;
    uint32_t f = c.tag(&wt);
    if (f == 1) {  // dims: packed or repeated varint
      if (wt == 2) { Cursor s = c.sub(); while (!s.done()) t->dims.push_back((int64_t)s.varint()); }
      else t->dims.push_back((int64_t)c.varint());
    } else if (f == 2 && wt == 0) {
      t->data_type = (int)c.varint();
    } else if (f == 4) {  // float_data
      if (wt == 2) {
        Cursor s = c.sub();
        size_t n = (size_t)(s.end - s.p) / 4;
        size_t old = t->float_data.size();
        t->float_data.resize(old + n);
        memcpy(t->float_data.data() + old, s.p, n * 4);
      } else if (wt == 5) {
        if (c.end - c.p < 4) throw std::runtime_error("onnx: truncated fixed32 in float_data");
        float v; memcpy(&v, c.p, 4); c.p += 4; t->float_data.push_back(v);
      } else c.skip(wt);
    } else if (f == 7) {  // int64_data
      if (wt == 2) { Cursor s = c.sub(); while (!s.done()) t->int64_data.push_back((int64_t)s.varint()); }
      else t->int64_data.push_back((int64_t)c.varint());
    } else if (f == 8 && wt == 2) {
      t->name = c.str();
    } else if (f == 9 && wt == 2) {
      Cursor s = c.sub();
      t->raw = s.p;
      t->raw_bytes = (size_t)(s.end - s.p);
    } else if (f == 13 && wt == 2) {   // external_data: repeated StringStringEntryProto{key=1, value=2}
      t->external = true;
      Cursor s = c.sub();
      std::string key, val;
      while (!s.done()) {
        int w2;
        uint32_t f2 = s.tag(&w2);
        if (f2 == 1 && w2 == 2) key = s.str();
        else if (f2 == 2 && w2 == 2) val = s.str();
        else s.skip(w2);
      }
      if (key == "location") t->ext_location = val;
      else if (key == "offset") t->ext_offset = strtoull(val.c_str(), nullptr, 10);
      else if (key == "length") t->ext_length = strtoull(val.c_str(), nullptr, 10);
    } else if (f == 14 && wt == 0) {
      if (c.varint() == 1) t->external = true;
    } else {
      c.skip(wt);
    }
  }
}

void parse_attr(Cursor c, OnnxAttr* a) {
  while (!c.done()) {
    int wt;
    uint32_t f = c.tag(&wt);
    if (f == 1 && wt == 2) a->name = c.str();
    else if (f == 2 && wt == 5) {
      if (c.end - c.p < 4) throw std::runtime_error("onnx: truncated fixed32 attribute");
      memcpy(&a->f, c.p, 4);
      c.p += 4;
    }
    else if (f == 3 && wt == 0) a->i = (int64_t)c.varint();
    else if (f == 5 && wt == 2) { a->has_t = true; parse_tensor(c.sub(), &a->t); }
    else if (f == 8) {
      if (wt == 2) { Cursor s = c.sub(); while (!s.done()) a->ints.push_back((int64_t)s.varint()); }
      else a->ints.push_back((int64_t)c.varint());
    } else c.skip(wt);
  }
}

void parse_node(Cursor c, OnnxNode* n) {
  while (!c.done()) {
    int wt;
    uint32_t f = c.tag(&wt);
    if (f == 1 && wt == 2) n->inputs.push_back(c.str());
    else if (f == 2 && wt == 2) n->outputs.push_back(c.str());
    else if (f == 3 && wt == 2) n->name = c.str();
    else if (f == 4 && wt == 2) n->op_type = c.str();
    else if (f == 5 && wt == 2) { n->attrs.emplace_back(); parse_attr(c.sub(), &n->attrs.back()); }
    else c.skip(wt);
  }
}

std::string parse_value_info_name(Cursor c) {
  std::string name;
  while (!c.done()) {
    int wt;
    uint32_t f = c.tag(&wt);
    if (f == 1 && wt == 2) name = c.str();
    else c.skip(wt);
  }
  return name;
}

void parse_graph(Cursor c, OnnxGraph* g) {
  while (!c.done()) {
    int wt;
    uint32_t f = c.tag(&wt);
    if (f == 1 && wt == 2) { g->nodes.emplace_back(); parse_node(c.sub(), &g->nodes.back()); }
    else if (f == 5 && wt == 2) { g->initializers.emplace_back(); parse_tensor(c.sub(), &g->initializers.back()); }
    else if (f == 11 && wt == 2) g->input_names.push_back(parse_value_info_name(c.sub()));
    else if (f == 12 && wt == 2) g->output_names.push_back(parse_value_info_name(c.sub()));
    else c.skip(wt);
  }
}

}  // namespace

void OnnxGraph::load(const std::string& path) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) throw std::runtime_error("onnx: cannot open " + path);
  fseek(fp, 0, SEEK_END);
  long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (sz <= 0) { fclose(fp); throw std::runtime_error("onnx: empty file " + path); }
  file.resize((size_t)sz);
  size_t got = fread(file.data(), 1, (size_t)sz, fp);
  fclose(fp);
  if (got != (size_t)sz) throw std::runtime_error("onnx: short read " + path);

  Cursor c{file.data(), file.data() + file.size()};
  bool saw_graph = false;
  while (!c.done()) {
    int wt;
    uint32_t f = c.tag(&wt);
    if (f == 1 && wt == 0) ir_version = (int64_t)c.varint();
    else if (f == 2 && wt == 2) producer = c.str();
    else if (f == 7 && wt == 2) { parse_graph(c.sub(), this); saw_graph = true; }
    else if (f == 8 && wt == 2) {
      Cursor s = c.sub();
      std::string domain; int64_t ver = 0;
      while (!s.done()) {
        int w2; uint32_t f2 = s.tag(&w2);
        if (f2 == 1 && w2 == 2) domain = s.str();
        else if (f2 == 2 && w2 == 0) ver = (int64_t)s.varint();
        else s.skip(w2);
      }
      if (domain.empty() || domain == "ai.onnx") opset = ver;
    } else c.skip(wt);
  }
  if (!saw_graph) throw std::runtime_error("onnx: no graph in " + path);
  // external-data tensors (models > 2 GB): raw bytes live in a sidecar file next to model.onnx
  const std::string dir = path.find_last_of('/') == std::string::npos ? std::string(".") : path.substr(0, path.find_last_of('/'));
  for (size_t i = 0; i < initializers.size(); ++i) {
    OnnxTensor& t = initializers[i];
    for (int64_t d : t.dims)
      if (d < 0) throw std::runtime_error("onnx: negative dimension in " + t.name);
    if (t.numel() < 0) throw std::runtime_error("onnx: element count overflow in " + t.name);
    if (t.external) {
      if (t.ext_location.empty() || t.ext_location.find("..") != std::string::npos || t.ext_location[0] == '/')
        throw std::runtime_error("onnx: external tensor " + t.name + " has no usable relative location");
      auto it = ext_files.find(t.ext_location);
      if (it == ext_files.end()) {
        const std::string ep = dir + "/" + t.ext_location;
        FILE* ef = fopen(ep.c_str(), "rb");
        if (!ef) throw std::runtime_error("onnx: cannot open external data file " + ep);
        fseek(ef, 0, SEEK_END);
        const long esz = ftell(ef);
        fseek(ef, 0, SEEK_SET);
        std::vector<uint8_t> buf((size_t)(esz > 0 ? esz : 0));
        const size_t egot = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), ef);
        fclose(ef);
        if (egot != buf.size()) throw std::runtime_error("onnx: short read " + ep);
        it = ext_files.emplace(t.ext_location, std::move(buf)).first;
      }
      const std::vector<uint8_t>& buf = it->second;
      const uint64_t len = t.ext_length ? t.ext_length : (buf.size() > t.ext_offset ? buf.size() - t.ext_offset : 0);
      if (t.ext_offset > buf.size() || len > buf.size() - t.ext_offset)
        throw std::runtime_error("onnx: external tensor " + t.name + " lies outside " + t.ext_location);
      t.raw = buf.data() + t.ext_offset;
      t.raw_bytes = (size_t)len;
    }
    init_by_name[t.name] = (int)i;
  }
  for (size_t i = 0; i < nodes.size(); ++i)
    for (auto& o : nodes[i].outputs) producer_of[o] = (int)i;
}

const OnnxTensor* OnnxGraph::resolve(const std::string& value) const {
  std::string cur = value;
  for (int hop = 0; hop < 64; ++hop) {
    auto it = init_by_name.find(cur);
    if (it != init_by_name.end()) return &initializers[it->second];
    auto pn = producer_of.find(cur);
    if (pn == producer_of.end()) return nullptr;
    const OnnxNode& n = nodes[pn->second];
    if (n.op_type == "Identity" && n.inputs.size() == 1) { cur = n.inputs[0]; continue; }
    if (n.op_type == "Constant") {
      const OnnxAttr* a = n.attr("value");
      if (a && a->has_t) return &a->t;
      return nullptr;
    }
    return nullptr;
  }
  return nullptr;
}

bool OnnxGraph::scalar_int(const std::string& value, int64_t* out) const {
  const OnnxTensor* t = resolve(value);
  if (!t || t->numel() != 1) return false;
  if (t->data_type == 7) {
    if (t->raw && t->raw_bytes >= 8) { memcpy(out, t->raw, 8); return true; }
    if (!t->int64_data.empty()) { *out = t->int64_data[0]; return true; }
  } else if (t->data_type == 6) {
    if (t->raw && t->raw_bytes >= 4) { int32_t v; memcpy(&v, t->raw, 4); *out = v; return true; }
  }
  return false;
}

bool OnnxGraph::scalar_float(const std::string& value, float* out) const {
  const OnnxTensor* t = resolve(value);
  if (!t || t->numel() != 1) return false;
  if (t->data_type == 1) {
    if (t->raw && t->raw_bytes >= 4) { memcpy(out, t->raw, 4); return true; }
    if (!t->float_data.empty()) { *out = t->float_data[0]; return true; }
  } else if (t->data_type == 11 && t->raw && t->raw_bytes >= 8) {
    double d; memcpy(&d, t->raw, 8); *out = (float)d; return true;
  }
  return false;
}

void tensor_to_float(const OnnxTensor& t, float* out) {
  int64_t n = t.numel();
  if (n < 0) throw std::runtime_error("onnx: element count overflow in " + t.name);
  if (t.data_type == 16 && t.raw) {   // bfloat16 raw
    if ((int64_t)t.raw_bytes != n * 2) throw std::runtime_error("onnx: raw size mismatch for " + t.name);
    for (int64_t i = 0; i < n; ++i) {
      uint16_t h;
      memcpy(&h, t.raw + 2 * i, 2);
      const uint32_t u = (uint32_t)h << 16;
      memcpy(out + i, &u, 4);
    }
    return;
  }
  if (t.data_type != 1) throw std::runtime_error("onnx: tensor " + t.name + " is not float32");
  if (t.raw) {
    if ((int64_t)t.raw_bytes != n * 4) throw std::runtime_error("onnx: raw size mismatch for " + t.name);
    memcpy(out, t.raw, (size_t)n * 4);
  } else {
    if ((int64_t)t.float_data.size() != n) throw std::runtime_error("onnx: float_data size mismatch for " + t.name);
    memcpy(out, t.float_data.data(), (size_t)n * 4);
  }
}

}  // namespace glc
