// registry.h -- op / kernel registries keyed by name (reference scanner/engine/op_registry.*,
// kernel_registry.* :21-39, op_info.h, kernel_factory.h).  Populated by the static
// REGISTER_OP / REGISTER_KERNEL objects of every plugin .so the engine dlopens.
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "scanner/api/kernel.h"
#include "scanner/api/op.h"

namespace scanner {
namespace internal {

struct ColumnDesc {
  std::string name;
  proto::ColumnType type;
  std::string type_name;
};

struct OpInfo {
  std::string name;
  bool variadic_inputs = false;
  std::vector<ColumnDesc> input_columns;
  std::vector<ColumnDesc> output_columns;
  bool can_stencil = false;
  std::vector<int> preferred_stencil = {0};
  bool has_bounded_state = false;
  i32 warmup = 0;
  bool has_unbounded_state = false;
  std::string protobuf_name;
  std::string stream_protobuf_name;
};

struct KernelFactory {
  std::string op_name;
  proto::DeviceType device_type = proto::CPU;
  i32 max_devices = 1;
  std::map<std::string, proto::DeviceType> input_devices;
  std::map<std::string, proto::DeviceType> output_devices;
  std::map<std::string, FrameLayout> input_layouts;  // columns also accepted in a non-HWC layout
  bool can_batch = false;
  i32 preferred_batch_size = 1;
  KernelConstructor constructor;

  BaseKernel* new_instance(const KernelConfig& config) const { return constructor(config); }
};

class OpRegistry {
 public:
  Result add_op(const std::string& name, OpInfo info);
  const OpInfo* get_op_info(const std::string& name) const;
  bool has_op(const std::string& name) const;
  std::vector<std::string> names() const;

 private:
  mutable std::mutex mu_;
  std::map<std::string, OpInfo> ops_;
};

class KernelRegistry {
 public:
  // key = name + "_cpu" / "_gpu" (reference kernel_registry.cpp:36-39)
  void add_kernel(const std::string& name, KernelFactory factory);
  bool has_kernel(const std::string& name, proto::DeviceType type) const;
  const KernelFactory* get_kernel(const std::string& name, proto::DeviceType type) const;

 private:
  static std::string key(const std::string& name, proto::DeviceType t) {
    return name + (t == proto::GPU ? "_gpu" : "_cpu");
  }
  mutable std::mutex mu_;
  std::map<std::string, KernelFactory> kernels_;
};

OpRegistry* get_op_registry();
KernelRegistry* get_kernel_registry();

// dlopen(RTLD_NOW | RTLD_LOCAL) a plugin library; its static registrations run during the call
// (reference worker.cpp:732-744).  Returns an error Result if the library cannot be loaded.
Result load_op_library(const std::string& so_path);

// A kernel's way to fail the run without aborting the process (used by host-language kernels,
// callback_op.cpp): the message is parked in a slot of the calling thread and collected by the
// evaluate loop right after the kernel call returns.
void raise_kernel_error(const std::string& msg);
bool take_kernel_error(std::string* msg);

}  // namespace internal
}  // namespace scanner
