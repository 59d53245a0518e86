// h264.h -- H.264 Annex-B bitstream utilities of the decode stage.
//  * index_bytestream(): the sample/keyframe index the load stage slices by.  Restates what the
//    reference builds at ingest (scanner/video/h264_byte_stream_index_creator.cpp:52-229 using
//    scanner/util/h264.h:135-438): per access unit byte offset + size, keyframe (IDR) frame
//    indices, coded size from the SPS, and the SPS/PPS "metadata packets" replayed on a seek.
//  * write_ipcm_stream(): a minimal, self-contained encoder for synthetic inputs (no x264 /
//    NVENC is available offline): Baseline profile, one slice per picture, every macroblock
//    I_PCM (lossless; decoders reproduce the source planes bit-exactly), IDR + SPS + PPS at each
//    GOP start, the other pictures either P slices of I_PCM macroblocks ("pcm") or a single
//    mb_skip_run ("skip"), or -- "bidir", Main profile with POC type 0 -- I_PCM anchors at the
//    even GOP positions and non-reference B pictures (one B_Skip run, spatial direct: the rounded
//    mean of the two anchors) at the odd ones, coded AFTER their later anchor so that decode
//    order differs from display order.  Used by tests and bench.py to make the H.264 of BASELINE.json's
//    configs.
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "scanner/util/common.h"

namespace scanner {
namespace internal {

struct H264Index {
  i32 width = 0, height = 0;              // display size (after SPS cropping)
  i32 coded_width = 0, coded_height = 0;  // multiples of 16
  std::vector<u64> sample_offsets;        // per frame: first byte of its access unit
  std::vector<u64> sample_sizes;
  std::vector<i64> keyframe_indices;      // frame indices of IDR pictures
  std::vector<u8> metadata_packets;       // first SPS + PPS, with start codes
  // false when the SPS rules picture reordering out (pic_order_cnt_type 2, or a VUI with
  // max_num_reorder_frames == 0): display position k of a GOP is then its k-th sample and the
  // decoder need not be fed past the last wanted picture.
  bool may_reorder = false;
  i64 frames() const { return (i64)sample_offsets.size(); }
};

// Scans an Annex-B byte stream.  A new access unit starts at the first SPS/PPS/SEI/AUD NAL after
// a VCL NAL, or at a VCL NAL whose first_mb_in_slice is 0.
// parameter_sets_only: accept a buffer holding just SPS/PPS (a stored descriptor's metadata
// packets) and fill the picture geometry without requiring any picture.
Result index_bytestream(const u8* data, size_t size, H264Index& out, bool parameter_sets_only = false);

// An index that did not come from index_bytestream (a stored VideoDescriptor) against the byte
// stream it describes: every sample inside [0, stream_size), keyframes ascending, inside the frame
// range and starting at frame 0.
Result check_index(const H264Index& index, size_t stream_size);

enum class SynthNonKey { Pcm = 0, Skip = 1, Bidir = 2 };

// fill(frame_index, y, u, v): writes the 4:2:0 planes of one frame (y: w*h, u/v: (w/2)*(h/2)).
using PlaneFiller = std::function<void(i64, u8*, u8*, u8*)>;

// width/height must be even; frames at index k*gop are IDR.
void write_ipcm_stream(i32 width, i32 height, i64 frames, i32 gop, SynthNonKey non_key,
                       const PlaneFiller& fill, std::vector<u8>& out);

}  // namespace internal
}  // namespace scanner
