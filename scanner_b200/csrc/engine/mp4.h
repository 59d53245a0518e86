// mp4.h -- ISO base media (ISO/IEC 14496-12) container support for ingest, H.264 tracks only.
// The reference demuxes with libavformat + the h264_mp4toannexb bitstream filter
// (scanner/engine/ingest.cpp:54-168, :228-300); neither is available here, so this is a small
// stand-alone reader for what ingest needs from an .mp4/.mov file:
//   moov/trak[vide]/mdia/{mdhd, minf/stbl/{stsd(avc1|avc3 + avcC), stsz, stsc, stco|co64, stss}}
// and the same conversion the filter does: length-prefixed NAL units -> Annex-B start codes, with
// the avcC parameter sets put in front of every sync sample that does not carry its own.
// A writer for the same subset exists so tests and benchmarks can produce real .mp4 files from
// the synthetic encoder (and so FFmpeg can be used as an independent reader of them).
#pragma once
#include <string>
#include <vector>

#include "h264.h"
#include "scanner/util/common.h"

namespace scanner {
namespace internal {

struct Mp4Track {
  i32 width = 0, height = 0;        // from the sample entry
  u32 timescale = 0;                // mdhd
  u64 duration = 0;                 // mdhd, in timescale units
  i64 samples = 0;
  i64 sync_samples = 0;
  std::vector<u8> annexb;           // the elementary stream, one access unit per sample
};

// true if the buffer starts with a plausible top-level box ('ftyp', 'moov', 'mdat', 'free', ...)
bool looks_like_mp4(const u8* data, size_t size);

// Extracts the first H.264 video track.  Errors: no moov / no avc track / truncated tables /
// sample outside the file / fragmented file (moof) / unsupported codec (hev1, mp4v, ...).
Result demux_mp4(const u8* data, size_t size, Mp4Track& out);

// Wraps an Annex-B stream (as indexed by index_bytestream) into a non-fragmented .mp4: one sample
// per access unit, parameter sets moved into avcC, sync table from the IDR list.
// fps = fps_num / fps_den.
Result mux_mp4(const u8* annexb, size_t size, const H264Index& index, i32 fps_num, i32 fps_den,
               std::vector<u8>& out);

}  // namespace internal
}  // namespace scanner
