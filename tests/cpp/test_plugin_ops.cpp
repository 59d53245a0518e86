// test_plugin_ops.cpp -- a CPU op library written against scanner-b200's plugin API exactly the
// way a Scanner user writes one against the reference's (cf. the shape of the reference's
// tests/test_ops.cpp: REGISTER_OP / REGISTER_KERNEL, Kernel / BatchedKernel / StenciledKernel,
// new_buffer, insert_element, VideoKernel::check_frame).  TEST INFRASTRUCTURE: the CPU Histogram
// and Resize kernels here call the oracle (oracle/scn_oracle.c) -- they exist to exercise the
// engine's plumbing on a box without a GPU, they are not part of the product.
#include <cstring>

#include "scanner/api/kernel.h"
#include "scanner/api/op.h"
#include "scanner/util/memory.h"
#include "test_args.pb.h"

extern "C" {
void orc_hist16_u8c3(const uint8_t* frame, int width, int height, int32_t* out48);
void orc_resize_bilinear_u8(const uint8_t* src, int sw, int sh, int cn, uint8_t* dst, int dw, int dh);
void orc_blur_u8c3(const uint8_t* src, int width, int height, int kernel_size, uint8_t* dst);
}

namespace scanner {

// ---- stateful counter, the reference's TestIncrement semantics (test_ops.cpp:173-236):
// emits 0,1,2,... since the last reset; resets itself when the input index is not consecutive.
class TestIncrementKernel : public Kernel {
 public:
  TestIncrementKernel(const KernelConfig& config) : Kernel(config), device_(config.devices[0]) {}
  void reset() override { next_int_ = 0; }
  void execute(const Elements& in, Elements& out) override {
    if (last_row_ + 1 != in[0].index) {
      last_row_ = in[0].index - 1;
      reset();
    }
    last_row_++;
    u8* buffer = new_buffer(device_, sizeof(i64));
    *((i64*)buffer) = next_int_++;
    insert_element(out[0], buffer, sizeof(i64));
  }

 private:
  DeviceHandle device_;
  i64 next_int_ = 0;
  i64 last_row_ = 0;
};

REGISTER_OP(TestIncrementUnbounded).input("ignore").output("integer").unbounded_state();
REGISTER_OP(TestIncrementUnboundedFrame).frame_input("ignore").output("integer").unbounded_state();
REGISTER_OP(TestIncrementBounded).input("ignore").output("integer").bounded_state();
REGISTER_OP(TestIncrementBoundedFrame).frame_input("ignore").output("integer").bounded_state();
REGISTER_KERNEL(TestIncrementUnbounded, TestIncrementKernel).device(DeviceType::CPU).num_devices(1);
REGISTER_KERNEL(TestIncrementUnboundedFrame, TestIncrementKernel).device(DeviceType::CPU).num_devices(1);
REGISTER_KERNEL(TestIncrementBounded, TestIncrementKernel).device(DeviceType::CPU).num_devices(1);
REGISTER_KERNEL(TestIncrementBoundedFrame, TestIncrementKernel).device(DeviceType::CPU).num_devices(1);

// ---- stencil probe: output = the int64 payloads of the whole window, concatenated
class TestWindowKernel : public StenciledKernel {
 public:
  TestWindowKernel(const KernelConfig& config) : StenciledKernel(config) {}
  void execute(const StenciledElements& in, Elements& out) override {
    const Elements& window = in[0];
    u8* buffer = new_buffer(CPU_DEVICE, window.size() * sizeof(i64));
    for (size_t i = 0; i < window.size(); ++i) memcpy(buffer + i * 8, window[i].buffer, 8);
    insert_element(out[0], buffer, window.size() * sizeof(i64));
  }
};
REGISTER_OP(TestWindow).input("col").output("window").stencil({-1, 0, 1});
REGISTER_KERNEL(TestWindow, TestWindowKernel).device(DeviceType::CPU).num_devices(1);

// ---- batch probe: every output row carries (row payload, size of the batch it was computed in)
class TestBatchKernel : public BatchedKernel {
 public:
  TestBatchKernel(const KernelConfig& config) : BatchedKernel(config) {}
  void execute(const BatchedElements& in, BatchedElements& out) override {
    const i32 n = (i32)num_rows(in[0]);
    u8* block = new_block_buffer_size(CPU_DEVICE, 16, n);
    for (i32 i = 0; i < n; ++i) {
      memcpy(block + i * 16, in[0][i].buffer, 8);
      const i64 b = n;
      memcpy(block + i * 16 + 8, &b, 8);
      insert_element(out[0], block + i * 16, 16);
    }
  }
};
REGISTER_OP(TestBatch).input("col").output("pair");
REGISTER_KERNEL(TestBatch, TestBatchKernel).device(DeviceType::CPU).batch(4).num_devices(1);

// ---- two outputs, one of which callers may leave unread
class TestTwoOutKernel : public Kernel {
 public:
  TestTwoOutKernel(const KernelConfig& config) : Kernel(config) {}
  void execute(const Elements& in, Elements& out) override {
    for (int c = 0; c < 2; ++c) {
      u8* b = new_buffer(CPU_DEVICE, 8);
      i64 v;
      memcpy(&v, in[0].buffer, 8);
      v = c == 0 ? v * 2 : v + 1000;
      memcpy(b, &v, 8);
      insert_element(out[c], b, 8);
    }
  }
};
REGISTER_OP(TestTwoOut).input("col").output("twice").output("plus1000");
REGISTER_KERNEL(TestTwoOut, TestTwoOutKernel).device(DeviceType::CPU).num_devices(1);

// ---- args: init args (protobuf_name) scale, per-stream args (stream_protobuf_name) offset
class TestAffineKernel : public Kernel {
 public:
  TestAffineKernel(const KernelConfig& config) : Kernel(config) {
    TestScaleArgs a;
    if (!a.ParseFromArray(config.args.data(), (int)config.args.size()) || config.args.empty()) {
      RESULT_ERROR(&valid_, "Could not parse TestScaleArgs");
      return;
    }
    scale_ = a.scale();
    valid_.set_success(true);
  }
  void validate(Result* r) override { r->CopyFrom(valid_); }
  void new_stream(const std::vector<u8>& args) override {
    TestOffsetArgs a;
    a.ParseFromArray(args.data(), (int)args.size());
    offset_ = a.offset();
  }
  void execute(const Elements& in, Elements& out) override {
    i64 v;
    memcpy(&v, in[0].buffer, 8);
    v = v * scale_ + offset_;
    u8* b = new_buffer(CPU_DEVICE, 8);
    memcpy(b, &v, 8);
    insert_element(out[0], b, 8);
  }

 private:
  Result valid_;
  i64 scale_ = 1, offset_ = 0;
};
REGISTER_OP(TestAffine).input("col").output("out").protobuf_name("TestScaleArgs").stream_protobuf_name("TestOffsetArgs");
REGISTER_KERNEL(TestAffine, TestAffineKernel).device(DeviceType::CPU).num_devices(1);

// ---- a C++ kernel that meets bad data: scanner::report_kernel_error instead of LOG(FATAL)
class TestRefuseValueKernel : public Kernel {
 public:
  TestRefuseValueKernel(const KernelConfig& config) : Kernel(config) {
    TestScaleArgs a;  // the value that cannot be processed travels in the `scale` field
    a.ParseFromArray(config.args.data(), (int)config.args.size());
    refused_ = a.scale();
  }
  void execute(const Elements& in, Elements& out) override {
    i64 v;
    memcpy(&v, in[0].buffer, 8);
    if (v == refused_) {
      report_kernel_error("TestRefuseValue cannot process the value " + std::to_string(v) + " (row " +
                          std::to_string(in[0].index) + ")");
      out[0] = Element();  // still one element per row: null
      return;
    }
    u8* b = new_buffer(CPU_DEVICE, 8);
    memcpy(b, &v, 8);
    insert_element(out[0], b, 8);
  }

 private:
  i64 refused_ = -1;
};
REGISTER_OP(TestRefuseValue).input("col").output("out").protobuf_name("TestScaleArgs");
REGISTER_KERNEL(TestRefuseValue, TestRefuseValueKernel).device(DeviceType::CPU).num_devices(1);

// ---- frame ops on the CPU through the ORACLE (plumbing tests only)
class TestHistogramOracleKernel : public BatchedKernel {
 public:
  TestHistogramOracleKernel(const KernelConfig& config) : BatchedKernel(config) {}
  void execute(const BatchedElements& in, BatchedElements& out) override {
    const i32 n = (i32)num_rows(in[0]);
    u8* block = new_block_buffer_size(CPU_DEVICE, 192, n);
    for (i32 i = 0; i < n; ++i) {
      const Frame* f = in[0][i].as_const_frame();
      orc_hist16_u8c3(f->data, f->width(), f->height(), (int32_t*)(block + i * 192));
      insert_element(out[0], block + i * 192, 192);
    }
  }
};
REGISTER_OP(TestHistogramOracle).frame_input("frame").output("histogram", ColumnType::Bytes, "Histogram");
REGISTER_KERNEL(TestHistogramOracle, TestHistogramOracleKernel).device(DeviceType::CPU).batch(8).num_devices(1);

class TestResizeOracleKernel : public BatchedKernel, public VideoKernel {
 public:
  TestResizeOracleKernel(const KernelConfig& config) : BatchedKernel(config) {}
  void new_stream(const std::vector<u8>& args) override { args_.ParseFromArray(args.data(), (int)args.size()); }
  void new_frame_info() override { ++frame_info_changes_; }
  void execute(const BatchedElements& in, BatchedElements& out) override {
    check_frame(CPU_DEVICE, in[0][0]);
    const i32 n = (i32)num_rows(in[0]);
    FrameInfo info(args_.height(), args_.width(), 3, FrameType::U8);
    std::vector<Frame*> frames = new_frames(CPU_DEVICE, info, n);
    for (i32 i = 0; i < n; ++i) {
      const Frame* f = in[0][i].as_const_frame();
      orc_resize_bilinear_u8(f->data, f->width(), f->height(), 3, frames[i]->data, args_.width(), args_.height());
      insert_frame(out[0], frames[i]);
    }
  }

 private:
  TestSizeArgs args_;
  int frame_info_changes_ = 0;
};
REGISTER_OP(TestResizeOracle).frame_input("frame").frame_output("frame").stream_protobuf_name("TestSizeArgs");
REGISTER_KERNEL(TestResizeOracle, TestResizeOracleKernel).device(DeviceType::CPU).batch(4).num_devices(1);

// frame stencil: |frame[t+1] - frame[t]| summed, as one f64 (a miniature of the OpticalFlow op's
// {0,1} stencil, reference test_ops.cpp:104-107)
class TestFrameDiffKernel : public StenciledKernel {
 public:
  TestFrameDiffKernel(const KernelConfig& config) : StenciledKernel(config) {}
  void execute(const StenciledElements& in, Elements& out) override {
    const Frame* a = in[0][0].as_const_frame();
    const Frame* b = in[0][1].as_const_frame();
    f64 acc = 0;
    for (size_t i = 0; i < a->size(); ++i) acc += std::abs((int)a->data[i] - (int)b->data[i]);
    u8* buf = new_buffer(CPU_DEVICE, 8);
    memcpy(buf, &acc, 8);
    insert_element(out[0], buf, 8);
  }
};
REGISTER_OP(TestFrameDiff).frame_input("frame").output("diff").stencil({0, 1});
REGISTER_KERNEL(TestFrameDiff, TestFrameDiffKernel).device(DeviceType::CPU).num_devices(1);

// a kernel that produces the wrong number of rows -> the engine must report it
class TestBadCountKernel : public BatchedKernel {
 public:
  TestBadCountKernel(const KernelConfig& config) : BatchedKernel(config) {}
  void execute(const BatchedElements& in, BatchedElements& out) override {
    for (size_t i = 0; i + 1 < in[0].size() + 1 && i < 1; ++i) insert_element(out[0], new_buffer(CPU_DEVICE, 1), 1);
  }
};
REGISTER_OP(TestBadCount).input("col").output("out");
REGISTER_KERNEL(TestBadCount, TestBadCountKernel).device(DeviceType::CPU).batch(3).num_devices(1);

}  // namespace scanner
