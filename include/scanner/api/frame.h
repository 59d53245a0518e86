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

// scanner-b200 extension (not in the reference): how the bytes of a frame element are laid out.
//   HWC   dense height x width x channels, what every reference op expects (and the default)
//   NV12  a decoder surface kept in its native form: shape = (H*3/2, W, 1) U8 -- W x H luma rows
//         followed by W x H/2 rows of interleaved Cb,Cr (the convention cv::COLOR_YUV2RGB_NV12
//         takes).  The decode stage emits it ONLY for a video column whose every consumer kernel
//         registered `.input_layout(col, FrameLayout::NV12)`; such kernels must produce what they
//         would have produced from the RGB24 frame the reference's decoder delivers
//         (scanner/util/image.cu:109-200).
enum class FrameLayout : u8 { HWC = 0, NV12 = 1 };

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

  // NV12 surface of a width x height picture (both even)
  static FrameInfo nv12(int width, int height) {
    FrameInfo f(height + height / 2, width, 1, FrameType::U8);
    f.layout = FrameLayout::NV12;
    return f;
  }

  int shape[FRAME_DIMS] = {0, 0, 0};
  FrameType type = FrameType::U8;
  FrameLayout layout = FrameLayout::HWC;
};

class Frame {
 public:
  Frame(FrameInfo info, u8* buffer);

  FrameInfo as_frame_info() const {
    FrameInfo f(shape[0], shape[1], shape[2], type);
    f.layout = layout;
    return f;
  }
  size_t size() const { return as_frame_info().size(); }
  int width() const { return shape[1]; }
  int height() const { return shape[0]; }
  int channels() const { return shape[2]; }

  int shape[FRAME_DIMS];
  FrameType type;
  u8* data;
  FrameLayout layout = FrameLayout::HWC;  // extension, see above
};

Frame* new_frame(DeviceHandle device, FrameInfo info);
// `num` frames carved out of ONE block allocation (contiguous: frame i at base + i*info.size()),
// which is what lets a batched GPU kernel treat its outputs as a strided array.
std::vector<Frame*> new_frames(DeviceHandle device, FrameInfo info, i32 num);

}  // namespace scanner
