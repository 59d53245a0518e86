// scanner/util/proto_lite.h -- a ~150-line proto3 wire-format reader/writer.
// Scanner passes op arguments (KernelConfig.args, new_stream(args)) as serialized protobuf
// messages (SURVEY 8b "Argument encoding").  protobuf-C++ is not a dependency of scanner-b200;
// tools/scn_protoc.py generates a header-only class per message on top of these primitives with
// the accessor spellings protoc's C++ output has (ParseFromArray, x(), set_x(), add_x() ...), so
// reference op sources that `#include "my_op.pb.h"` keep compiling.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace scanner {
namespace protolite {

enum Wire : uint32_t { VARINT = 0, FIXED64 = 1, LEN = 2, FIXED32 = 5 };

class Reader {
 public:
  Reader(const void* data, size_t size)
    : p_((const uint8_t*)data), end_((const uint8_t*)data + size) {}
  bool done() const { return p_ >= end_; }
  bool ok() const { return ok_; }

  // next tag; false at end of buffer or on malformed input (check ok())
  bool next(uint32_t& field, uint32_t& wire) {
    if (p_ >= end_) return false;
    uint64_t tag;
    if (!varint(tag)) return false;
    field = (uint32_t)(tag >> 3);
    wire = (uint32_t)(tag & 7);
    if (field == 0) ok_ = false;
    return ok_;
  }
  bool varint(uint64_t& v) {
    v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p_ >= end_) return fail();
      const uint8_t b = *p_++;
      v |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) return true;
    }
    return fail();
  }
  bool fixed32(uint32_t& v) {
    if (end_ - p_ < 4) return fail();
    memcpy(&v, p_, 4);
    p_ += 4;
    return true;
  }
  bool fixed64(uint64_t& v) {
    if (end_ - p_ < 8) return fail();
    memcpy(&v, p_, 8);
    p_ += 8;
    return true;
  }
  bool bytes(const uint8_t*& data, size_t& size) {
    uint64_t n;
    if (!varint(n)) return false;
    if ((uint64_t)(end_ - p_) < n) return fail();
    data = p_;
    size = (size_t)n;
    p_ += n;
    return true;
  }
  bool skip(uint32_t wire) {
    uint64_t v64;
    uint32_t v32;
    const uint8_t* d;
    size_t n;
    switch (wire) {
      case VARINT: return varint(v64);
      case FIXED64: return fixed64(v64);
      case LEN: return bytes(d, n);
      case FIXED32: return fixed32(v32);
      default: return fail();
    }
  }

 private:
  bool fail() {
    ok_ = false;
    return false;
  }
  const uint8_t* p_;
  const uint8_t* end_;
  bool ok_ = true;
};

class Writer {
 public:
  void tag(uint32_t field, uint32_t wire) { raw_varint(((uint64_t)field << 3) | wire); }
  void raw_varint(uint64_t v) {
    while (v >= 0x80) {
      out_.push_back((char)(v | 0x80));
      v >>= 7;
    }
    out_.push_back((char)v);
  }
  void varint_field(uint32_t f, uint64_t v) {
    tag(f, VARINT);
    raw_varint(v);
  }
  void fixed32_field(uint32_t f, uint32_t v) {
    tag(f, FIXED32);
    out_.append((const char*)&v, 4);
  }
  void fixed64_field(uint32_t f, uint64_t v) {
    tag(f, FIXED64);
    out_.append((const char*)&v, 8);
  }
  void bytes_field(uint32_t f, const void* d, size_t n) {
    tag(f, LEN);
    raw_varint(n);
    out_.append((const char*)d, n);
  }
  std::string& str() { return out_; }

 private:
  std::string out_;
};

inline uint64_t zigzag(int64_t v) { return ((uint64_t)v << 1) ^ (uint64_t)(v >> 63); }
inline int64_t unzigzag(uint64_t v) { return (int64_t)(v >> 1) ^ -(int64_t)(v & 1); }

}  // namespace protolite
}  // namespace scanner
