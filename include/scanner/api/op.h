// scanner/api/op.h -- REGISTER_OP: how an op library tells the engine what an op looks like.
//
// Source-compatible with the reference's registration vocabulary (scanner/api/op.h:36-136) so that
// op sources compile unchanged:
//
//   REGISTER_OP(Histogram).frame_input("frame").output("histogram", ColumnType::Bytes, "Histogram");
//   REGISTER_OP(Window).input("col").output("out").stencil({-1, 0, 1});
//   REGISTER_OP(Resize).frame_input("frame").frame_output("frame").stream_protobuf_name("ResizeArgs");
//
// The chain fills an OpDeclaration; the registration object built from it at static-initialisation
// time (i.e. inside dlopen of the op library) files the declaration with the engine's op registry.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "scanner/util/common.h"

namespace scanner {

// Everything the engine knows about an op before any kernel exists.
struct OpDeclaration {
  struct Column {
    std::string name;
    ColumnType type;
    std::string type_name;  // outputs only: stored as the column's type name ("Histogram", ...)
  };

  std::string name;
  std::vector<Column> inputs;   // empty and `variadic`: any number of byte columns
  std::vector<Column> outputs;
  bool variadic = false;
  bool stencils = false;                // .stencil() was called: the op may look at neighbouring rows
  std::vector<int> default_stencil{0};
  int warmup = -1;                      // >= 0: bounded state, that many rows are re-run at a task start
  bool unbounded = false;               // needs every earlier row of its stream
  std::string args_message;             // proto3 message carried in KernelConfig::args
  std::string stream_args_message;      // proto3 message handed to new_stream()
};

namespace internal {

class OpBuilder {
 public:
  explicit OpBuilder(std::string name) { decl_.name = std::move(name); }

  // ---- inputs: named columns, or variadic (never both)
  OpBuilder& input(const std::string& name, ColumnType type = ColumnType::Bytes) {
    exclusive(decl_.variadic, "fixed and variadic inputs");
    decl_.inputs.push_back({name, type, ""});
    return *this;
  }
  OpBuilder& frame_input(const std::string& name) { return input(name, ColumnType::Video); }
  OpBuilder& variadic_inputs() {
    exclusive(!decl_.inputs.empty(), "fixed and variadic inputs");
    decl_.variadic = true;
    return *this;
  }

  // ---- outputs
  OpBuilder& output(const std::string& name, ColumnType type = ColumnType::Bytes, std::string type_name = "") {
    decl_.outputs.push_back({name, type, std::move(type_name)});
    return *this;
  }
  OpBuilder& frame_output(const std::string& name) { return output(name, ColumnType::Video); }

  // ---- temporal behaviour
  OpBuilder& stencil(const std::vector<int>& offsets = {0}) {
    decl_.stencils = true;
    decl_.default_stencil = offsets;
    return *this;
  }
  OpBuilder& bounded_state(i32 warmup = 0) {
    exclusive(decl_.unbounded, "bounded and unbounded state");
    decl_.warmup = warmup;
    return *this;
  }
  OpBuilder& unbounded_state() {
    exclusive(decl_.warmup >= 0, "bounded and unbounded state");
    decl_.unbounded = true;
    return *this;
  }

  // ---- argument messages
  OpBuilder& protobuf_name(std::string message) {
    decl_.args_message = std::move(message);
    return *this;
  }
  OpBuilder& stream_protobuf_name(std::string message) {
    decl_.stream_args_message = std::move(message);
    return *this;
  }

  const OpDeclaration& declaration() const { return decl_; }

 private:
  void exclusive(bool clash, const char* what) const {
    if (clash) LOG(FATAL) << "Op " << decl_.name << " cannot have both " << what;
  }
  OpDeclaration decl_;
};

// Constructing one registers the op (registry.cpp); a duplicate name is fatal, as upstream.
struct OpRegistration {
  OpRegistration(const OpBuilder& builder);
};

}  // namespace internal

#define SCN_API_PASTE2(a__, b__) a__##b__
#define SCN_API_PASTE(a__, b__) SCN_API_PASTE2(a__, b__)
// One static registration object per use; the builder chain that follows initialises it.
#define REGISTER_OP(name__)                                                                        \
  static const ::scanner::internal::OpRegistration SCN_API_PASTE(scn_registered_op_, __COUNTER__) \
      __attribute__((unused)) = ::scanner::internal::OpBuilder(#name__)

}  // namespace scanner
