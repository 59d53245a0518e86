// scanner/api/op.h -- REGISTER_OP: declares an op's columns and temporal behaviour to the engine
// (reference scanner/api/op.h:36-136, op.cpp:23-61).  Same builder vocabulary:
//   .input(name[,type]) .frame_input(name) .variadic_inputs() .output(name[,type[,type_name]])
//   .frame_output(name) .stencil({..}) .bounded_state(warmup) .unbounded_state()
//   .protobuf_name("XArgs") .stream_protobuf_name("XArgs")
#pragma once
#include <string>
#include <tuple>
#include <vector>

#include "scanner/util/common.h"

namespace scanner {
namespace internal {

class OpBuilder;

class OpRegistration {
 public:
  OpRegistration(const OpBuilder& builder);
};

class OpBuilder {
 public:
  friend class OpRegistration;
  OpBuilder(const std::string& name) : name_(name) {}

  OpBuilder& variadic_inputs() {
    if (!input_columns_.empty())
      LOG(FATAL) << "Op " << name_ << " cannot have both fixed and variadic inputs";
    variadic_inputs_ = true;
    return *this;
  }
  OpBuilder& input(const std::string& name, ColumnType type = ColumnType::Bytes) {
    if (variadic_inputs_)
      LOG(FATAL) << "Op " << name_ << " cannot have both fixed and variadic inputs";
    input_columns_.push_back(std::make_tuple(name, type));
    return *this;
  }
  OpBuilder& frame_input(const std::string& name) { return input(name, ColumnType::Video); }
  OpBuilder& output(const std::string& name, ColumnType type = ColumnType::Bytes,
                    std::string type_name = "") {
    output_columns_.push_back(std::make_tuple(name, type, type_name));
    return *this;
  }
  OpBuilder& frame_output(const std::string& name) { return output(name, ColumnType::Video); }
  OpBuilder& stencil(const std::vector<int>& stencil = {0}) {
    can_stencil_ = true;
    preferred_stencil_ = stencil;
    return *this;
  }
  OpBuilder& bounded_state(i32 warmup = 0) {
    if (has_unbounded_state_)
      LOG(FATAL) << "Op " << name_ << " was already declared to have unbounded state";
    has_bounded_state_ = true;
    warmup_ = warmup;
    return *this;
  }
  OpBuilder& unbounded_state() {
    if (has_bounded_state_)
      LOG(FATAL) << "Op " << name_ << " was already declared to have bounded state";
    has_unbounded_state_ = true;
    return *this;
  }
  OpBuilder& protobuf_name(std::string protobuf_name) {
    protobuf_name_ = protobuf_name;
    return *this;
  }
  OpBuilder& stream_protobuf_name(std::string protobuf_name) {
    stream_protobuf_name_ = protobuf_name;
    return *this;
  }

 private:
  std::string name_;
  bool variadic_inputs_ = false;
  std::vector<std::tuple<std::string, ColumnType>> input_columns_;
  std::vector<std::tuple<std::string, ColumnType, std::string>> output_columns_;
  bool can_stencil_ = false;
  std::vector<int> preferred_stencil_ = {0};
  bool has_bounded_state_ = false;
  i32 warmup_ = 0;
  bool has_unbounded_state_ = false;
  std::string protobuf_name_;
  std::string stream_protobuf_name_;
};
}  // namespace internal

#define REGISTER_OP(name__) REGISTER_OP_HELPER(__COUNTER__, name__)
#define REGISTER_OP_HELPER(uid__, name__) REGISTER_OP_UID(uid__, name__)
#define REGISTER_OP_UID(uid__, name__)                                                    \
  static ::scanner::internal::OpRegistration op_registration_##uid__ __attribute__((unused)) = \
      ::scanner::internal::OpBuilder(#name__)

}  // namespace scanner
