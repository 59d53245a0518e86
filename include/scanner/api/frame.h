// scanner/api/frame.h -- Frame / FrameInfo, the dense-HWC image element type of the plugin API
// (reference scanner/api/frame.h:34-82, frame.cpp:22-118).  Same fields, same accessors, same
// allocation helpers; the memory comes from scanner-b200's stream-ordered allocator.
#pragma once
#include <cstring>
#include <vector>

#include "scanner/util/common.h"
#include "scanner/util/memory.h"

namespace scanner {

size_t size_of_frame_type(FrameType type);

const i32 FRAME_DIMS = 3;

struct FrameInfo {
  FrameInfo() = default;
  FrameInfo(int shape0, int shape1, int shape2, FrameType type);
  FrameInfo(const std::vector<int> shapes, FrameType type);

  bool operator==(const FrameInfo& other) const;
  bool operator!=(const FrameInfo& other) const { return !(*this == other); }

  size_t size() const;
  int width() const { return shape[1]; }     // valid for (height, width, channels)
  int height() const { return shape[0]; }
  int channels() const { return shape[2]; }

  int shape[FRAME_DIMS] = {0, 0, 0};
  FrameType type = FrameType::U8;
};

class Frame {
 public:
  Frame(FrameInfo info, u8* buffer);

  FrameInfo as_frame_info() const { return FrameInfo(shape[0], shape[1], shape[2], type); }
  size_t size() const { return as_frame_info().size(); }
  int width() const { return shape[1]; }
  int height() const { return shape[0]; }
  int channels() const { return shape[2]; }

  int shape[FRAME_DIMS];
  FrameType type;
  u8* data;
};

Frame* new_frame(DeviceHandle device, FrameInfo info);
// `num` frames carved out of ONE block allocation (contiguous: frame i at base + i*info.size()),
// which is what lets a batched GPU kernel treat its outputs as a strided array.
std::vector<Frame*> new_frames(DeviceHandle device, FrameInfo info, i32 num);

}  // namespace scanner
