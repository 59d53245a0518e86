#include "registry.h"

#include <dlfcn.h>

namespace scanner {

namespace logging {
int& verbosity() {
  static int v = [] {
    const char* e = getenv("SCANNER_VLOG");
    return e ? atoi(e) : 0;
  }();
  return v;
}
}  // namespace logging

namespace internal {

OpRegistry* get_op_registry() {
  static OpRegistry* r = new OpRegistry();
  return r;
}
KernelRegistry* get_kernel_registry() {
  static KernelRegistry* r = new KernelRegistry();
  return r;
}

Result OpRegistry::add_op(const std::string& name, OpInfo info) {
  Result result;
  std::lock_guard<std::mutex> g(mu_);
  if (ops_.count(name)) {
    RESULT_ERROR(&result, "Attempted to re-register op %s", name.c_str());
    return result;
  }
  if (info.input_columns.empty() && !info.variadic_inputs) {
    RESULT_ERROR(&result, "Attempted to register op %s with empty input columns.", name.c_str());
    return result;
  }
  if (info.output_columns.empty()) {
    RESULT_ERROR(&result, "Attempted to register op %s with empty output columns.", name.c_str());
    return result;
  }
  ops_[name] = std::move(info);
  result.set_success(true);
  return result;
}
const OpInfo* OpRegistry::get_op_info(const std::string& name) const {
  std::lock_guard<std::mutex> g(mu_);
  auto it = ops_.find(name);
  return it == ops_.end() ? nullptr : &it->second;
}
bool OpRegistry::has_op(const std::string& name) const { return get_op_info(name) != nullptr; }
std::vector<std::string> OpRegistry::names() const {
  std::lock_guard<std::mutex> g(mu_);
  std::vector<std::string> out;
  for (auto& kv : ops_) out.push_back(kv.first);
  return out;
}

void KernelRegistry::add_kernel(const std::string& name, KernelFactory factory) {
  std::lock_guard<std::mutex> g(mu_);
  const std::string k = key(name, factory.device_type);
  if (kernels_.count(k)) LOG(WARNING) << "kernel " << k << " registered twice; keeping the first";
  else kernels_[k] = std::move(factory);
}
bool KernelRegistry::has_kernel(const std::string& name, proto::DeviceType type) const {
  return get_kernel(name, type) != nullptr;
}
const KernelFactory* KernelRegistry::get_kernel(const std::string& name,
                                                proto::DeviceType type) const {
  std::lock_guard<std::mutex> g(mu_);
  auto it = kernels_.find(key(name, type));
  return it == kernels_.end() ? nullptr : &it->second;
}

Result load_op_library(const std::string& so_path) {
  Result result;
  void* handle = dlopen(so_path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!handle) {
    RESULT_ERROR(&result, "Failed to load op library %s: %s", so_path.c_str(), dlerror());
    return result;
  }
  result.set_success(true);
  return result;
}

// ---- registration objects constructed by REGISTER_OP / REGISTER_KERNEL ---------------------
OpRegistration::OpRegistration(const OpBuilder& builder) {
  const OpDeclaration& d = builder.declaration();
  OpInfo info;
  info.name = d.name;
  info.variadic_inputs = d.variadic;
  for (const auto& c : d.inputs) info.input_columns.push_back({c.name, (proto::ColumnType)c.type, ""});
  for (const auto& c : d.outputs) info.output_columns.push_back({c.name, (proto::ColumnType)c.type, c.type_name});
  info.can_stencil = d.stencils;
  info.preferred_stencil = d.default_stencil;
  info.has_bounded_state = d.warmup >= 0;
  info.warmup = d.warmup >= 0 ? d.warmup : 0;
  info.has_unbounded_state = d.unbounded;
  info.protobuf_name = d.args_message;
  info.stream_protobuf_name = d.stream_args_message;
  Result r = get_op_registry()->add_op(d.name, info);
  LOG_IF(FATAL, !r.success()) << "Failed to register op " << d.name << ": " << r.msg();
}

KernelRegistration::KernelRegistration(const KernelBuilder& builder) {
  const KernelDeclaration& d = builder.declaration();
  KernelFactory f;
  f.op_name = d.op_name;
  f.device_type = (proto::DeviceType)d.device;
  f.max_devices = d.max_devices;
  for (const auto& kv : d.input_devices) f.input_devices[kv.first] = (proto::DeviceType)kv.second;
  for (const auto& kv : d.output_devices) f.output_devices[kv.first] = (proto::DeviceType)kv.second;
  f.input_layouts = d.input_layouts;
  f.can_batch = d.batches;
  f.preferred_batch_size = d.batch_size;
  f.constructor = d.make;
  get_kernel_registry()->add_kernel(d.op_name, std::move(f));
}

}  // namespace internal
}  // namespace scanner
