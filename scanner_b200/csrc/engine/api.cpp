// api.cpp -- out-of-line parts of scanner/api/{frame,kernel}.h (reference frame.cpp:22-118,
// kernel.cpp:34-109).
#include "scanner/api/frame.h"
#include "scanner/api/kernel.h"

namespace scanner {

size_t size_of_frame_type(FrameType type) {
  switch ((proto::FrameType)type) {
    case proto::U8: return 1;
    case proto::U16: return 2;
    case proto::F32: return 4;
    case proto::F64: return 8;
  }
  return 1;
}

FrameInfo::FrameInfo(int s0, int s1, int s2, FrameType t) : type(t) {
  shape[0] = s0;
  shape[1] = s1;
  shape[2] = s2;
}
FrameInfo::FrameInfo(const std::vector<int> shapes, FrameType t) : type(t) {
  for (size_t i = 0; i < shapes.size() && i < (size_t)FRAME_DIMS; ++i) shape[i] = shapes[i];
}
bool FrameInfo::operator==(const FrameInfo& o) const {
  return (proto::FrameType)type == (proto::FrameType)o.type && shape[0] == o.shape[0] &&
         shape[1] == o.shape[1] && shape[2] == o.shape[2] && layout == o.layout;
}
size_t FrameInfo::size() const {
  return size_of_frame_type(type) * (size_t)shape[0] * (size_t)shape[1] * (size_t)shape[2];
}

Frame::Frame(FrameInfo info, u8* b) : type(info.type), data(b), layout(info.layout) {
  memcpy(shape, info.shape, sizeof(shape));
}

Frame* new_frame(DeviceHandle device, FrameInfo info) {
  return new Frame(info, new_buffer(device, info.size()));
}

std::vector<Frame*> new_frames(DeviceHandle device, FrameInfo info, i32 num) {
  std::vector<Frame*> frames;
  if (num <= 0) return frames;
  u8* base = new_block_buffer_size(device, info.size(), num);
  frames.reserve(num);
  for (i32 i = 0; i < num; ++i) frames.push_back(new Frame(info, base + (size_t)i * info.size()));
  return frames;
}

// ---- element ownership -------------------------------------------------------------------------
Element add_element_ref(DeviceHandle device, Element& element) {
  Element second;
  if (element.is_null()) return second;
  if (element.is_frame) {
    Frame* frame = element.as_frame();
    add_buffer_ref(device, frame->data);
    second = Element(new Frame(frame->as_frame_info(), frame->data));
  } else {
    add_buffer_ref(device, element.buffer);
    second = element;
  }
  second.index = element.index;
  return second;
}

void delete_element(DeviceHandle device, Element& element) {
  if (element.is_null()) return;
  if (!element.is_frame) {
    delete_buffer(device, element.buffer);
    return;
  }
  Frame* frame = element.as_frame();
  delete_buffer(device, frame->data);
  delete frame;
}

// ---- calling-convention adapters: the engine always hands column -> batch -> stencil ---------
void StenciledBatchedKernel::execute_kernel(const StenciledBatchedElements& in,
                                            BatchedElements& out) {
  execute(in, out);
}

void BatchedKernel::execute_kernel(const StenciledBatchedElements& in, BatchedElements& out) {
  BatchedElements flat(in.size());
  for (size_t c = 0; c < in.size(); ++c) {
    flat[c].reserve(in[c].size());
    for (const Elements& stencil : in[c]) flat[c].push_back(stencil[0]);
  }
  execute(flat, out);
}

void StenciledKernel::execute_kernel(const StenciledBatchedElements& in, BatchedElements& out) {
  StenciledElements one(in.size());
  for (size_t c = 0; c < in.size(); ++c) one[c] = in[c][0];
  Elements row(out.size());
  execute(one, row);
  for (size_t c = 0; c < row.size(); ++c) out[c].push_back(row[c]);
}

void Kernel::execute_kernel(const StenciledBatchedElements& in, BatchedElements& out) {
  Elements one;
  one.reserve(in.size());
  for (const auto& col : in) one.push_back(col[0][0]);
  Elements row(out.size());
  execute(one, row);
  for (size_t c = 0; c < row.size(); ++c) out[c].push_back(row[c]);
}

void VideoKernel::check_frame(const DeviceHandle&, const Element& element) {
  const Frame* frame = element.as_const_frame();
  if (!(frame->as_frame_info() == frame_info_)) {
    frame_info_ = frame->as_frame_info();
    new_frame_info();
  }
}

void VideoKernel::check_frame_info(const DeviceHandle& device, const Element& element) {
  FrameInfo info;
  memcpy_buffer((u8*)&info, CPU_DEVICE, element.buffer, device, sizeof(FrameInfo));
  if (!(info == frame_info_)) {
    frame_info_ = info;
    new_frame_info();
  }
}

}  // namespace scanner
