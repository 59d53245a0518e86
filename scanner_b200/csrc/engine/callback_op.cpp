// callback_op.cpp -- ops whose kernel is a host-language callable (include/scn_engine.h "kernels
// written in the host language").  The reference runs `@scannerpy.register_python_op` kernels through
// an embedded interpreter (scanner/engine/python_kernel.cpp:49-361: one PythonKernel C++ object per
// kernel instance forwarding new_stream / reset / execute, inputs converted to numpy arrays or bytes,
// outputs copied into new_buffer / new_frame blocks); here the engine is a library INSIDE the host
// process, so the same forwarding goes through one C function pointer per op and the host side
// (scanner_b200/pyops.py) does the conversions.  The callback runs on the evaluate thread of the
// pipeline instance that owns the kernel; a ctypes callback takes the GIL itself.
//
// Failure model: the reference treats an exception in a Python kernel as a worker failure; in-process
// that would take the user's interpreter down, so a failing callback marks the kernel error slot of
// the calling thread instead and the evaluate loop turns it into a failed run with the message.
#include <atomic>
#include <cstring>

#include "registry.h"
#include "scanner/api/kernel.h"
#include "scanner/util/memory.h"
#include "scn_engine.h"

namespace scanner {
namespace internal {

namespace {
thread_local std::string t_kernel_error;
thread_local bool t_kernel_failed = false;
}  // namespace

void raise_kernel_error(const std::string& msg) {
  if (t_kernel_failed) return;  // the first error is the informative one
  t_kernel_failed = true;
  t_kernel_error = msg;
}
bool take_kernel_error(std::string* msg) {
  if (!t_kernel_failed) return false;
  if (msg) *msg = t_kernel_error;
  t_kernel_failed = false;
  t_kernel_error.clear();
  return true;
}

}  // namespace internal

void report_kernel_error(const std::string& message) { internal::raise_kernel_error(message); }

namespace internal {

namespace {

struct CallbackOp {
  std::string name;
  scn_kernel_callback cb = nullptr;
  void* user = nullptr;
  std::vector<bool> out_is_frame;
};

struct EmitCtx {
  BatchedElements* out;
  const CallbackOp* op;
  size_t rows;
  std::string error;
};

std::atomic<int64_t> g_next_instance{1};

class CallbackKernel : public BaseKernel {
 public:
  CallbackKernel(const KernelConfig& config, const CallbackOp* op)
    : BaseKernel(config), op_(op), instance_(g_next_instance++), device_(config.devices[0]), node_id_(config.node_id) {
    scn_cb_call c = base_call(SCN_CB_CONSTRUCT);
    c.args = config.args.data();
    c.args_size = config.args.size();
    constructed_ = invoke(c, &ctor_error_);
  }
  ~CallbackKernel() override {
    if (!constructed_) return;
    scn_cb_call c = base_call(SCN_CB_DESTROY);
    std::string ignored;
    invoke(c, &ignored);
  }

  void validate(proto::Result* result) override {
    if (constructed_) {
      result->set_success(true);
    } else {
      RESULT_ERROR(result, "%s", ctor_error_.c_str());
    }
  }
  void fetch_resources(proto::Result* result) override { lifecycle(SCN_CB_FETCH_RESOURCES, result); }
  void setup_with_resources(proto::Result* result) override { lifecycle(SCN_CB_SETUP_WITH_RESOURCES, result); }

  void new_stream(const std::vector<u8>& args) override {
    scn_cb_call c = base_call(SCN_CB_NEW_STREAM);
    c.args = args.data();
    c.args_size = args.size();
    std::string err;
    if (!invoke(c, &err)) raise_kernel_error(err);
  }
  void reset() override {
    scn_cb_call c = base_call(SCN_CB_RESET);
    std::string err;
    if (!invoke(c, &err)) raise_kernel_error(err);
  }

  void execute_kernel(const StenciledBatchedElements& in, BatchedElements& out) override {
    const size_t cols = in.size(), rows = cols ? in[0].size() : 0, sten = rows ? in[0][0].size() : 0;
    std::vector<scn_cb_elem> elems(cols * rows * sten);
    size_t k = 0;
    for (size_t c = 0; c < cols; ++c)
      for (size_t r = 0; r < rows; ++r)
        for (size_t s = 0; s < sten; ++s, ++k) {
          const Element& e = in[c][r][s];
          scn_cb_elem& d = elems[k];
          memset(&d, 0, sizeof(d));
          d.index = e.index;
          d.frame_type = -1;
          if (e.is_null()) continue;
          if (e.is_frame) {
            const Frame* f = e.as_const_frame();
            d.data = f->data;
            d.size = f->size();
            d.shape[0] = f->shape[0];
            d.shape[1] = f->shape[1];
            d.shape[2] = f->shape[2];
            d.frame_type = (int32_t)(proto::FrameType)f->type;
          } else {
            d.data = e.buffer;
            d.size = e.size;
          }
        }
    EmitCtx ctx{&out, op_, rows, ""};
    scn_cb_call c = base_call(SCN_CB_EXECUTE);
    c.n_cols = (int)cols;
    c.n_rows = (int)rows;
    c.n_stencil = (int)sten;
    c.elems = elems.data();
    c.out = &ctx;
    std::string err;
    if (!invoke(c, &err)) {
      raise_kernel_error(err);
    } else if (!ctx.error.empty()) {
      raise_kernel_error("Op " + op_->name + ": " + ctx.error);
    }
  }

 private:
  scn_cb_call base_call(int event) const {
    scn_cb_call c;
    memset(&c, 0, sizeof(c));
    c.event = event;
    c.instance = instance_;
    c.device_type = (int)(proto::DeviceType)device_.type;
    c.device_id = device_.id;
    c.node_id = node_id_;
    return c;
  }
  bool invoke(const scn_cb_call& c, std::string* err) const {
    char buf[4096];
    buf[0] = 0;
    const int rc = op_->cb(op_->user, &c, buf, sizeof(buf));
    if (rc == 0) return true;
    buf[sizeof(buf) - 1] = 0;
    *err = buf[0] ? std::string(buf) : "kernel callback of op " + op_->name + " failed (code " + std::to_string(rc) + ")";
    return false;
  }
  void lifecycle(int event, proto::Result* result) {
    scn_cb_call c = base_call(event);
    std::string err;
    if (invoke(c, &err)) {
      result->set_success(true);
    } else {
      RESULT_ERROR(result, "%s", err.c_str());
    }
  }

  const CallbackOp* op_;
  const int64_t instance_;
  const DeviceHandle device_;
  const i32 node_id_;
  bool constructed_ = false;
  std::string ctor_error_;
};

EmitCtx* emit_ctx(void* out, int col, std::string* why) {
  EmitCtx* ctx = (EmitCtx*)out;
  if (!ctx || !ctx->out) {
    *why = "null output handle";
    return nullptr;
  }
  if (col < 0 || (size_t)col >= ctx->out->size()) {
    ctx->error = "output column " + std::to_string(col) + " out of range";
    *why = ctx->error;
    return nullptr;
  }
  if ((*ctx->out)[col].size() >= ctx->rows) {
    ctx->error = "more than " + std::to_string(ctx->rows) + " elements emitted for output column " + std::to_string(col);
    *why = ctx->error;
    return nullptr;
  }
  return ctx;
}

}  // namespace
}  // namespace internal
}  // namespace scanner

using namespace scanner;
using namespace scanner::internal;

extern "C" {

int scn_cb_emit_bytes(void* out, int col, const uint8_t* data, size_t size) {
  std::string why;
  EmitCtx* ctx = emit_ctx(out, col, &why);
  if (!ctx) return -1;
  if (ctx->op->out_is_frame[(size_t)col] && size != 0) {
    ctx->error = "output column " + std::to_string(col) + " is a frame column: emit a frame";
    return -1;
  }
  if (!data || size == 0) {
    (*ctx->out)[col].push_back(Element());  // null row
    return 0;
  }
  u8* buf = new_buffer(CPU_DEVICE, size);
  memcpy(buf, data, size);
  insert_element((*ctx->out)[col], buf, size);
  return 0;
}

int scn_cb_emit_frame(void* out, int col, const uint8_t* data, int height, int width, int channels, int frame_type) {
  std::string why;
  EmitCtx* ctx = emit_ctx(out, col, &why);
  if (!ctx) return -1;
  if (!ctx->op->out_is_frame[(size_t)col]) {
    ctx->error = "output column " + std::to_string(col) + " is a bytes column: emit bytes";
    return -1;
  }
  if (!data || height <= 0 || width <= 0 || channels <= 0 || frame_type < 0 || frame_type > (int)proto::U16) {
    ctx->error = "bad frame emitted for output column " + std::to_string(col);
    return -1;
  }
  Frame* f = new_frame(CPU_DEVICE, FrameInfo(height, width, channels, FrameType((proto::FrameType)frame_type)));
  memcpy(f->data, data, f->size());
  insert_frame((*ctx->out)[col], f);
  return 0;
}

}  // extern "C"

namespace scanner {
namespace internal {

Result register_callback_op(const scn_cb_op_desc& d, scn_kernel_callback cb, void* user) {
  Result r;
  if (!d.name || !d.name[0] || !cb || d.n_outputs <= 0 || (d.n_inputs <= 0 && !d.variadic_inputs) ||
      (d.n_inputs > 0 && (!d.input_names || !d.input_is_frame)) || !d.output_names || !d.output_is_frame ||
      (d.n_stencil > 0 && !d.stencil) || d.batch < 1 || (d.bounded_state >= 0 && d.unbounded_state)) {
    RESULT_ERROR(&r, "register_callback_op: bad op description");
    return r;
  }
  OpInfo info;
  info.name = d.name;
  info.variadic_inputs = d.variadic_inputs != 0;
  if (!info.variadic_inputs)
    for (int i = 0; i < d.n_inputs; ++i)
      info.input_columns.push_back({d.input_names[i], d.input_is_frame[i] ? proto::Video : proto::Bytes, ""});
  // the ops and kernels registries keep the CallbackOp alive for the life of the process, like the
  // static registration objects of a dlopen'ed plugin
  CallbackOp* op = new CallbackOp();
  op->name = d.name;
  op->cb = cb;
  op->user = user;
  for (int i = 0; i < d.n_outputs; ++i) {
    const bool frame = d.output_is_frame[i] != 0;
    info.output_columns.push_back({d.output_names[i], frame ? proto::Video : proto::Bytes,
                                   d.output_type_names && d.output_type_names[i] ? d.output_type_names[i] : ""});
    op->out_is_frame.push_back(frame);
  }
  if (d.n_stencil > 0) {
    info.can_stencil = true;
    info.preferred_stencil.assign(d.stencil, d.stencil + d.n_stencil);
  }
  if (d.bounded_state >= 0) {
    info.has_bounded_state = true;
    info.warmup = d.bounded_state;
  }
  info.has_unbounded_state = d.unbounded_state != 0;
  r = get_op_registry()->add_op(info.name, info);
  if (!r.success()) {
    delete op;
    return r;
  }
  // Host-language kernels read and write host memory whichever device type they are scheduled
  // as (python_kernel.cpp copies to the CPU first): a GPU registration pins every named column
  // to the CPU and lets the engine do the moves.
  for (int dev = 0; dev < 2; ++dev) {
    if (dev == 1 && (!d.also_gpu || info.variadic_inputs)) continue;
    KernelFactory f;
    f.op_name = info.name;
    f.device_type = dev == 0 ? proto::CPU : proto::GPU;
    f.max_devices = 1;
    f.can_batch = d.batch > 1;
    f.preferred_batch_size = d.batch;
    if (dev == 1) {
      for (auto& c : info.input_columns) f.input_devices[c.name] = proto::CPU;
      for (auto& c : info.output_columns) f.output_devices[c.name] = proto::CPU;
    }
    f.constructor = [op](const KernelConfig& cfg) -> BaseKernel* { return new CallbackKernel(cfg, op); };
    get_kernel_registry()->add_kernel(info.name, std::move(f));
  }
  r.set_success(true);
  return r;
}

}  // namespace internal
}  // namespace scanner
