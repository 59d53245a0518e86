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
OpRegistration::OpRegistration(const OpBuilder& b) {
  OpInfo info;
  info.name = b.name_;
  info.variadic_inputs = b.variadic_inputs_;
  for (auto& c : b.input_columns_)
    info.input_columns.push_back({std::get<0>(c), (proto::ColumnType)std::get<1>(c), ""});
  for (auto& c : b.output_columns_)
    info.output_columns.push_back({std::get<0>(c), (proto::ColumnType)std::get<1>(c), std::get<2>(c)});
  info.can_stencil = b.can_stencil_;
  info.preferred_stencil = b.preferred_stencil_;
  info.has_bounded_state = b.has_bounded_state_;
  info.warmup = b.warmup_;
  info.has_unbounded_state = b.has_unbounded_state_;
  info.protobuf_name = b.protobuf_name_;
  info.stream_protobuf_name = b.stream_protobuf_name_;
  Result r = get_op_registry()->add_op(b.name_, info);
  LOG_IF(FATAL, !r.success()) << "Failed to register op " << b.name_ << ": " << r.msg();
}

KernelRegistration::KernelRegistration(const KernelBuilder& b) {
  KernelFactory f;
  f.op_name = b.name_;
  f.device_type = (proto::DeviceType)b.device_type_;
  f.max_devices = b.num_devices_;
  for (auto& kv : b.input_devices_) f.input_devices[kv.first] = (proto::DeviceType)kv.second;
  for (auto& kv : b.output_devices_) f.output_devices[kv.first] = (proto::DeviceType)kv.second;
  f.input_layouts = b.input_layouts_;
  f.can_batch = b.can_batch_;
  f.preferred_batch_size = b.preferred_batch_size_;
  f.constructor = b.constructor_;
  get_kernel_registry()->add_kernel(b.name_, std::move(f));
}

}  // namespace internal
}  // namespace scanner
