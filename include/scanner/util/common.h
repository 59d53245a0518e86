// scanner/util/common.h -- scanner-b200's re-declaration of the basic types Scanner op authors
// see (reference scanner/util/common.h:33-118).  Source-compatible subset: the scalar aliases,
// DeviceType / DeviceHandle / CPU_DEVICE, ColumnType, FrameType, Result, RESULT_ERROR and a
// glog-shaped LOG()/VLOG()/LOG_IF() so plugin sources written against the reference compile
// unchanged.  The reference gets these enums and `Result` from protobuf-generated headers
// (scanner/metadata.proto, types.proto); protobuf-C++ is not a dependency here, so they are
// plain C++ with the same names, enumerators and accessor spellings.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace scanner {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i8 = int8_t;
using i16 = int16_t;
using i32 = int32_t;
using i64 = int64_t;
using f32 = float;
using f64 = double;

namespace proto {
// Values match the reference's proto enums (scanner/metadata.proto) so serialized metadata
// written by either implementation reads back in the other.
enum DeviceType : int { CPU = 0, GPU = 1 };
enum ColumnType : int { Bytes = 0, Video = 1 };
enum FrameType : int { U8 = 0, F32 = 1, F64 = 2, U16 = 3 };

// proto::Result{bool success; string msg} (reference scanner/metadata.proto).
class Result {
 public:
  bool success() const { return success_; }
  void set_success(bool s) { success_ = s; }
  const std::string& msg() const { return msg_; }
  void set_msg(const std::string& m) { msg_ = m; }
  void CopyFrom(const Result& o) { *this = o; }

 private:
  bool success_ = false;
  std::string msg_;
};
}  // namespace proto

// The reference spells enumerators both as DeviceType::CPU and proto::DeviceType::CPU.
struct DeviceType {
  static constexpr proto::DeviceType CPU = proto::CPU;
  static constexpr proto::DeviceType GPU = proto::GPU;
  DeviceType() = default;
  constexpr DeviceType(proto::DeviceType v) : v_(v) {}
  constexpr operator proto::DeviceType() const { return v_; }
  proto::DeviceType v_ = proto::CPU;
};
struct ColumnType {
  static constexpr proto::ColumnType Bytes = proto::Bytes;
  static constexpr proto::ColumnType Video = proto::Video;
  ColumnType() = default;
  constexpr ColumnType(proto::ColumnType v) : v_(v) {}
  constexpr operator proto::ColumnType() const { return v_; }
  proto::ColumnType v_ = proto::Bytes;
};
struct FrameType {
  static constexpr proto::FrameType U8 = proto::U8;
  static constexpr proto::FrameType F32 = proto::F32;
  static constexpr proto::FrameType F64 = proto::F64;
  static constexpr proto::FrameType U16 = proto::U16;
  FrameType() = default;
  constexpr FrameType(proto::FrameType v) : v_(v) {}
  constexpr operator proto::FrameType() const { return v_; }
  proto::FrameType v_ = proto::U8;
};
using proto::Result;

// Where a buffer lives / where a kernel runs (reference common.h:58-88).
struct DeviceHandle {
  DeviceHandle(DeviceType type_, i32 id_) : type(type_), id(id_) {}
  DeviceHandle() = default;

  bool operator==(const DeviceHandle& o) const {
    return (proto::DeviceType)type == (proto::DeviceType)o.type && id == o.id;
  }
  bool operator!=(const DeviceHandle& o) const { return !(*this == o); }
  bool operator<(const DeviceHandle& o) const {
    if ((int)(proto::DeviceType)type != (int)(proto::DeviceType)o.type)
      return (int)(proto::DeviceType)type < (int)(proto::DeviceType)o.type;
    return id < o.id;
  }
  bool can_copy_to(const DeviceHandle&) const { return true; }  // NVSwitch: any GPU pair is peer
  bool is_same_address_space(const DeviceHandle& o) const {
    const bool cpu = (proto::DeviceType)type == proto::CPU;
    const bool ocpu = (proto::DeviceType)o.type == proto::CPU;
    return (cpu && ocpu) || (!cpu && !ocpu && id == o.id);
  }
  bool is_gpu() const { return (proto::DeviceType)type == proto::GPU; }

  DeviceType type = DeviceType::CPU;
  i32 id = 0;
};

inline std::ostream& operator<<(std::ostream& os, const DeviceHandle& h) {
  return os << (h.is_gpu() ? "GPU:" : "CPU:") << h.id;
}

static const DeviceHandle CPU_DEVICE = {DeviceType::CPU, 0};

struct Interval {
  i32 start;
  i32 end;
};

// ------------------------------------------------------------------------------------------
// glog-shaped logging.  LOG(FATAL) aborts, exactly like the reference treats kernel errors
// (SURVEY 8b "Errors": runtime errors are process death, handled as a worker failure).
namespace logging {
enum Severity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
int& verbosity();  // SCANNER_VLOG env, default 0
class Message {
 public:
  Message(const char* file, int line, Severity s) : sev_(s) {
    static const char* names[] = {"I", "W", "E", "F"};
    ss_ << names[s] << " " << file << ":" << line << "] ";
  }
  ~Message() {
    ss_ << "\n";
    std::cerr << ss_.str();
    if (sev_ == FATAL) {
      std::cerr.flush();
      std::abort();
    }
  }
  std::ostream& stream() { return ss_; }

 private:
  std::ostringstream ss_;
  Severity sev_;
};
struct Voidify {
  void operator&(std::ostream&) {}
};
}  // namespace logging

#define LOG(sev__) \
  ::scanner::logging::Message(__FILE__, __LINE__, ::scanner::logging::sev__).stream()
#define LOG_IF(sev__, cond__) \
  !(cond__) ? (void)0 : ::scanner::logging::Voidify() & LOG(sev__)
#define VLOG(n__) LOG_IF(INFO, ::scanner::logging::verbosity() >= (n__))
#define CHECK(cond__) LOG_IF(FATAL, !(cond__)) << "Check failed: " #cond__ " "

#define RESULT_ERROR(result__, str__, ...)          \
  {                                                 \
    char errstr__[1024];                            \
    snprintf(errstr__, 1024, str__, ##__VA_ARGS__); \
    LOG(ERROR) << errstr__;                         \
    (result__)->set_success(false);                 \
    (result__)->set_msg(errstr__);                  \
  }

}  // namespace scanner
