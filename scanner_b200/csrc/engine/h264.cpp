// h264.cpp -- see h264.h.  Syntax element names follow ITU-T H.264 (7.3.x).
#include "h264.h"

#include <cstring>

namespace scanner {
namespace internal {

namespace {

// ---- reading -------------------------------------------------------------------------------
struct Nal {
  size_t start;        // offset of the start code
  size_t payload;      // offset of the NAL header byte
  size_t end;          // one past the last byte of the NAL
  int type;
  int ref_idc;
};

// next start code 00 00 01 at or after `pos`; returns size if none
size_t find_start_code(const u8* d, size_t size, size_t pos) {
  for (size_t i = pos; i + 2 < size; ++i) {
    if (d[i + 2] > 1) {
      i += 2;  // cannot be part of a start code ending here or in the next two positions
      continue;
    }
    if (d[i] == 0 && d[i + 1] == 0 && d[i + 2] == 1) return i;
  }
  return size;
}

// RBSP bit reader with emulation-prevention removal (00 00 03 -> 00 00)
class BitReader {
 public:
  BitReader(const u8* d, size_t n) {
    buf_.reserve(n);
    int zeros = 0;
    for (size_t i = 0; i < n; ++i) {
      if (zeros >= 2 && d[i] == 3) {
        zeros = 0;
        continue;
      }
      buf_.push_back(d[i]);
      zeros = d[i] == 0 ? zeros + 1 : 0;
    }
  }
  bool ok() const { return ok_; }
  u32 u(int n) {
    u32 v = 0;
    for (int i = 0; i < n; ++i) {
      if (pos_ >= buf_.size() * 8) {
        ok_ = false;
        return 0;
      }
      v = (v << 1) | ((buf_[pos_ >> 3] >> (7 - (pos_ & 7))) & 1);
      ++pos_;
    }
    return v;
  }
  u32 ue() {
    int zeros = 0;
    while (ok_ && u(1) == 0) {
      if (++zeros > 32) {
        ok_ = false;
        return 0;
      }
    }
    if (!ok_) return 0;
    return (zeros ? ((1u << zeros) - 1 + u(zeros)) : 0);
  }
  i32 se() {
    const u32 k = ue();
    return (k & 1) ? (i32)((k + 1) >> 1) : -(i32)(k >> 1);
  }

 private:
  std::vector<u8> buf_;
  size_t pos_ = 0;
  bool ok_ = true;
};

void skip_scaling_list(BitReader& br, int n) {
  int last = 8, next = 8;
  for (int j = 0; j < n; ++j) {
    if (next != 0) next = (last + br.se() + 256) % 256;
    last = next == 0 ? last : next;
  }
}

void skip_hrd(BitReader& br) {
  const u32 cpb = br.ue() + 1;
  br.u(4);
  br.u(4);
  for (u32 i = 0; i < cpb && br.ok(); ++i) {
    br.ue();
    br.ue();
    br.u(1);
  }
  br.u(5);
  br.u(5);
  br.u(5);
  br.u(5);
}

// VUI (E.1.1) up to bitstream_restriction; reorder = max_num_reorder_frames, or -1 if not coded.
bool parse_vui_reorder(BitReader& br, i64& reorder) {
  reorder = -1;
  if (br.u(1)) {                          // aspect_ratio_info_present_flag
    if (br.u(8) == 255) {                 // Extended_SAR
      br.u(16);
      br.u(16);
    }
  }
  if (br.u(1)) br.u(1);                   // overscan
  if (br.u(1)) {                          // video_signal_type
    br.u(3);
    br.u(1);
    if (br.u(1)) {
      br.u(8);
      br.u(8);
      br.u(8);
    }
  }
  if (br.u(1)) {                          // chroma_loc_info
    br.ue();
    br.ue();
  }
  if (br.u(1)) {                          // timing_info
    br.u(32);
    br.u(32);
    br.u(1);
  }
  const u32 nal_hrd = br.u(1);
  if (nal_hrd) skip_hrd(br);
  const u32 vcl_hrd = br.u(1);
  if (vcl_hrd) skip_hrd(br);
  if (nal_hrd || vcl_hrd) br.u(1);        // low_delay_hrd_flag
  br.u(1);                                // pic_struct_present_flag
  if (!br.ok()) return false;
  if (br.u(1)) {                          // bitstream_restriction_flag
    br.u(1);
    br.ue();
    br.ue();
    br.ue();
    br.ue();
    const u32 r = br.ue();                // max_num_reorder_frames
    br.ue();                              // max_dec_frame_buffering
    if (!br.ok()) return false;
    reorder = r;
  }
  return br.ok();
}

bool parse_sps(const u8* d, size_t n, H264Index& out) {
  BitReader br(d, n);
  const u32 profile = br.u(8);
  br.u(8);  // constraint flags
  br.u(8);  // level
  br.ue();  // sps id
  u32 chroma_format_idc = 1;
  if (profile == 100 || profile == 110 || profile == 122 || profile == 244 || profile == 44 || profile == 83 ||
      profile == 86 || profile == 118 || profile == 128 || profile == 138 || profile == 139 || profile == 134) {
    chroma_format_idc = br.ue();
    if (chroma_format_idc == 3) br.u(1);
    br.ue();
    br.ue();
    br.u(1);
    if (br.u(1)) {  // seq_scaling_matrix_present_flag
      const int lists = chroma_format_idc != 3 ? 8 : 12;
      for (int i = 0; i < lists; ++i)
        if (br.u(1)) skip_scaling_list(br, i < 6 ? 16 : 64);
    }
  }
  br.ue();  // log2_max_frame_num_minus4
  const u32 poc_type = br.ue();
  if (poc_type == 0) {
    br.ue();
  } else if (poc_type == 1) {
    br.u(1);
    br.se();
    br.se();
    const u32 cyc = br.ue();
    for (u32 i = 0; i < cyc && br.ok(); ++i) br.se();
  }
  br.ue();  // max_num_ref_frames
  br.u(1);
  const u32 w_mbs = br.ue() + 1;
  const u32 h_units = br.ue() + 1;
  const u32 frame_mbs_only = br.u(1);
  if (!frame_mbs_only) br.u(1);
  br.u(1);  // direct_8x8_inference_flag
  u32 cl = 0, cr = 0, ct = 0, cb = 0;
  if (br.u(1)) {
    cl = br.ue();
    cr = br.ue();
    ct = br.ue();
    cb = br.ue();
  }
  if (!br.ok()) return false;
  // Can pictures leave the decoder in another order than they enter it?  Not with POC type 2
  // (output order == decoding order, 8.2.1.3); otherwise only a VUI bitstream_restriction with
  // max_num_reorder_frames == 0 rules it out.
  out.may_reorder = poc_type != 2;
  if (br.u(1) && out.may_reorder) {  // vui_parameters_present_flag
    i64 reorder = -1;
    if (parse_vui_reorder(br, reorder) && reorder == 0) out.may_reorder = false;
  }
  out.coded_width = (i32)(w_mbs * 16);
  out.coded_height = (i32)((2 - frame_mbs_only) * h_units * 16);
  const u32 sub_w = (chroma_format_idc == 1 || chroma_format_idc == 2) ? 2 : 1;
  const u32 sub_h = chroma_format_idc == 1 ? 2 : 1;
  const u32 ux = chroma_format_idc == 0 ? 1 : sub_w;
  const u32 uy = (chroma_format_idc == 0 ? 1 : sub_h) * (2 - frame_mbs_only);
  out.width = out.coded_width - (i32)(ux * (cl + cr));
  out.height = out.coded_height - (i32)(uy * (ct + cb));
  return out.width > 0 && out.height > 0;
}

// ---- writing -------------------------------------------------------------------------------
class BitWriter {
 public:
  void u(int n, u32 v) {
    for (int i = n - 1; i >= 0; --i) bit((v >> i) & 1);
  }
  void ue(u32 v) {
    const u32 k = v + 1;
    int len = 0;
    while ((k >> len) > 1) ++len;
    u(len, 0);
    u(len + 1, k);
  }
  void se(i32 v) { ue(v > 0 ? (u32)(2 * v - 1) : (u32)(-2 * v)); }
  void align_zero() {
    while (nbits_ & 7) bit(0);
  }
  void bytes(const u8* d, size_t n) {  // only when aligned
    buf_.insert(buf_.end(), d, d + n);
    nbits_ += n * 8;
  }
  void trailing() {
    bit(1);
    align_zero();
  }
  const std::vector<u8>& data() const { return buf_; }

 private:
  void bit(u32 b) {
    if ((nbits_ & 7) == 0) buf_.push_back(0);
    if (b) buf_.back() |= (u8)(0x80 >> (nbits_ & 7));
    ++nbits_;
  }
  std::vector<u8> buf_;
  size_t nbits_ = 0;
};

// start code + header byte + RBSP with emulation prevention
void emit_nal(std::vector<u8>& out, int ref_idc, int type, const std::vector<u8>& rbsp) {
  static const u8 sc[4] = {0, 0, 0, 1};
  out.insert(out.end(), sc, sc + 4);
  out.push_back((u8)((ref_idc << 5) | type));
  const size_t base = out.size();
  out.resize(base + rbsp.size() + rbsp.size() / 2 + 8);
  u8* o = out.data() + base;
  size_t w = 0;
  int zeros = 0;
  for (size_t i = 0; i < rbsp.size(); ++i) {
    const u8 b = rbsp[i];
    if (zeros >= 2 && b <= 3) {
      o[w++] = 3;
      zeros = 0;
    }
    o[w++] = b;
    zeros = b == 0 ? zeros + 1 : 0;
  }
  out.resize(base + w);
}

}  // namespace

Result index_bytestream(const u8* data, size_t size, H264Index& out, bool parameter_sets_only) {
  Result r;
  out = H264Index();
  std::vector<Nal> nals;
  size_t pos = find_start_code(data, size, 0);
  while (pos < size) {
    Nal n;
    n.start = (pos > 0 && data[pos - 1] == 0) ? pos - 1 : pos;  // 4-byte start code
    n.payload = pos + 3;
    const size_t next = find_start_code(data, size, n.payload);
    size_t end = next;
    if (next < size && next > n.payload && data[next - 1] == 0) end = next - 1;
    n.end = end;
    if (n.payload < size) {
      n.type = data[n.payload] & 0x1F;
      n.ref_idc = (data[n.payload] >> 5) & 3;
      nals.push_back(n);
    }
    pos = next;
  }
  if (nals.empty()) {
    RESULT_ERROR(&r, "no NAL units found in %zu bytes", size);
    return r;
  }
  bool have_sps = false, have_pps = false, seen_vcl_in_au = false, au_open = false;
  size_t au_start = 0;
  bool au_is_idr = false;
  auto close_au = [&](size_t end) {
    if (!au_open || !seen_vcl_in_au) return;
    if (au_is_idr) out.keyframe_indices.push_back((i64)out.sample_offsets.size());
    out.sample_offsets.push_back(au_start);
    out.sample_sizes.push_back(end - au_start);
  };
  for (size_t i = 0; i < nals.size(); ++i) {
    const Nal& n = nals[i];
    const bool vcl = n.type >= 1 && n.type <= 5;
    bool starts_new = false;
    if (vcl) {
      BitReader br(data + n.payload + 1, std::min<size_t>(n.end - n.payload - 1, 16));
      const u32 first_mb = br.ue();
      starts_new = seen_vcl_in_au && first_mb == 0;
    } else if (n.type >= 6 && n.type <= 9) {
      starts_new = seen_vcl_in_au;
    }
    if (!au_open || starts_new) {
      close_au(n.start);
      au_open = true;
      au_start = n.start;
      seen_vcl_in_au = false;
      au_is_idr = false;
    }
    if (vcl) {
      seen_vcl_in_au = true;
      if (n.type == 5) au_is_idr = true;
    }
    if (n.type == 7 && !have_sps) {
      if (!parse_sps(data + n.payload + 1, n.end - n.payload - 1, out)) {
        RESULT_ERROR(&r, "could not parse SPS");
        return r;
      }
      out.metadata_packets.insert(out.metadata_packets.end(), data + n.start, data + n.end);
      have_sps = true;
    }
    if (n.type == 8 && !have_pps) {
      out.metadata_packets.insert(out.metadata_packets.end(), data + n.start, data + n.end);
      have_pps = true;
    }
  }
  close_au(size);
  if (!have_sps || !have_pps) {
    RESULT_ERROR(&r, "stream has no SPS/PPS");
    return r;
  }
  if (parameter_sets_only) {
    r.set_success(true);
    return r;
  }
  if (out.sample_offsets.empty() || out.keyframe_indices.empty() || out.keyframe_indices[0] != 0) {
    RESULT_ERROR(&r, "stream must start with an IDR picture (%zu frames, %zu keyframes)", out.sample_offsets.size(),
                 out.keyframe_indices.size());
    return r;
  }
  r.set_success(true);
  return r;
}

Result check_index(const H264Index& index, size_t stream_size) {
  Result r;
  if (index.sample_offsets.size() != index.sample_sizes.size()) {
    RESULT_ERROR(&r, "video index: %zu sample offsets but %zu sizes", index.sample_offsets.size(),
                 index.sample_sizes.size());
    return r;
  }
  for (size_t i = 0; i < index.sample_offsets.size(); ++i) {
    const u64 off = index.sample_offsets[i], sz = index.sample_sizes[i];
    if (off > stream_size || sz > stream_size - off) {
      RESULT_ERROR(&r, "video index: sample %zu (offset %llu, %llu bytes) lies outside the %zu-byte stream", i,
                   (unsigned long long)off, (unsigned long long)sz, stream_size);
      return r;
    }
  }
  i64 prev = -1;
  for (i64 k : index.keyframe_indices) {
    if (k <= prev || k >= index.frames()) {
      RESULT_ERROR(&r, "video index: keyframe list is not ascending inside [0, %ld)", (long)index.frames());
      return r;
    }
    prev = k;
  }
  if (index.frames() > 0 && (index.keyframe_indices.empty() || index.keyframe_indices[0] != 0)) {
    RESULT_ERROR(&r, "video index: the first frame is not a keyframe");
    return r;
  }
  if (index.width <= 0 || index.height <= 0 || index.width > index.coded_width || index.height > index.coded_height) {
    RESULT_ERROR(&r, "video index: picture size %dx%d does not fit the coded size %dx%d", index.width, index.height,
                 index.coded_width, index.coded_height);
    return r;
  }
  r.set_success(true);
  return r;
}

void write_ipcm_stream(i32 width, i32 height, i64 frames, i32 gop, SynthNonKey non_key,
                       const PlaneFiller& fill, std::vector<u8>& out) {
  const i32 wmb = (width + 15) / 16, hmb = (height + 15) / 16;
  const i32 cw = wmb * 16, ch = hmb * 16;
  if (gop < 1) gop = 1;
  const bool bidir = non_key == SynthNonKey::Bidir;

  std::vector<u8> sps, pps;
  {
    BitWriter b;
    b.u(8, bidir ? 77 : 66);      // profile_idc: Baseline, or Main when there are B pictures
    b.u(8, bidir ? 0x40 : 0xC0);  // constraint_set0/1
    b.u(8, 51);    // level_idc 5.1 (I_PCM bitrates are far above the level limits anyway)
    b.ue(0);       // seq_parameter_set_id
    b.ue(0);       // log2_max_frame_num_minus4 -> MaxFrameNum 16
    if (bidir) {
      b.ue(0);     // pic_order_cnt_type 0: POC lsb in every slice header
      b.ue(4);     // log2_max_pic_order_cnt_lsb_minus4 -> 8 bits
      b.ue(2);     // max_num_ref_frames: the two anchors around a B picture
    } else {
      b.ue(2);     // pic_order_cnt_type 2: output order == decode order
      b.ue(1);     // max_num_ref_frames
    }
    b.u(1, 0);     // gaps_in_frame_num_value_allowed_flag
    b.ue(wmb - 1);
    b.ue(hmb - 1);
    b.u(1, 1);     // frame_mbs_only_flag
    b.u(1, 1);     // direct_8x8_inference_flag
    const bool crop = cw != width || ch != height;
    b.u(1, crop ? 1 : 0);
    if (crop) {
      b.ue(0);
      b.ue((cw - width) / 2);
      b.ue(0);
      b.ue((ch - height) / 2);
    }
    b.u(1, 0);  // vui_parameters_present_flag
    b.trailing();
    sps = b.data();
  }
  {
    BitWriter b;
    b.ue(0);     // pic_parameter_set_id
    b.ue(0);     // seq_parameter_set_id
    b.u(1, 0);   // entropy_coding_mode_flag: CAVLC
    b.u(1, 0);   // bottom_field_pic_order_in_frame_present_flag
    b.ue(0);     // num_slice_groups_minus1
    b.ue(0);     // num_ref_idx_l0_default_active_minus1
    b.ue(0);     // num_ref_idx_l1_default_active_minus1
    b.u(1, 0);   // weighted_pred_flag
    b.u(2, 0);   // weighted_bipred_idc
    b.se(0);     // pic_init_qp_minus26
    b.se(0);     // pic_init_qs_minus26
    b.se(0);     // chroma_qp_index_offset
    b.u(1, 1);   // deblocking_filter_control_present_flag
    b.u(1, 0);   // constrained_intra_pred_flag
    b.u(1, 0);   // redundant_pic_cnt_present_flag
    b.trailing();
    pps = b.data();
  }

  std::vector<u8> y((size_t)cw * ch), u((size_t)(cw / 2) * (ch / 2)), v(u.size());
  std::vector<u8> ysrc((size_t)width * height), usrc((size_t)(width / 2) * (height / 2)), vsrc(usrc.size());
  u32 idr_id = 0;
  // Coding order.  Bidir: inside a GOP of L pictures the odd display positions that have a later
  // anchor in the same GOP are non-reference B pictures, coded after that anchor (I0 P2 B1 P4 B3 ..).
  std::vector<i64> order;
  order.reserve((size_t)frames);
  for (i64 g0 = 0; g0 < frames; g0 += gop) {
    const i64 len = std::min<i64>(gop, frames - g0);
    if (!bidir) {
      for (i64 k = 0; k < len; ++k) order.push_back(g0 + k);
      continue;
    }
    order.push_back(g0);
    for (i64 k = 1; k < len; k += 2) {
      if (k + 1 < len) order.push_back(g0 + k + 1);
      order.push_back(g0 + k);
    }
  }
  u32 refs_in_gop = 0;  // reference pictures coded so far in this GOP (frame_num)
  for (i64 f : order) {
    const bool key = (f % gop) == 0;
    const i64 in_gop = f % gop;
    const i64 gop_len = std::min<i64>(gop, frames - (f - in_gop));
    const bool bpic = bidir && (in_gop & 1) && in_gop + 1 < gop_len;
    if (key) {
      emit_nal(out, 3, 7, sps);
      emit_nal(out, 3, 8, pps);
      refs_in_gop = 0;
    }
    const bool skip = (!key && non_key == SynthNonKey::Skip) || bpic;
    BitWriter b;
    b.ue(0);                         // first_mb_in_slice
    b.ue(key ? 7 : (bpic ? 6 : 5));  // slice_type: I (7) / B (6) / P (5), "all slices of this type"
    b.ue(0);                         // pic_parameter_set_id
    // frame_num: reference pictures count up; a non-reference picture carries PrevRefFrameNum + 1
    b.u(4, (bidir ? refs_in_gop : (u32)in_gop) % 16);
    if (!bpic) ++refs_in_gop;
    if (key) b.ue(idr_id++ & 0xFFFF);  // idr_pic_id
    if (bidir) b.u(8, (u32)(2 * in_gop) & 0xFF);  // pic_order_cnt_lsb
    if (bpic) b.u(1, 1);  // direct_spatial_mv_pred_flag
    if (!key) {
      b.u(1, 0);  // num_ref_idx_active_override_flag
      b.u(1, 0);  // ref_pic_list_modification_flag_l0
      if (bpic) b.u(1, 0);  // ref_pic_list_modification_flag_l1
    }
    if (key) {
      b.u(1, 0);  // no_output_of_prior_pics_flag
      b.u(1, 0);  // long_term_reference_flag
    } else if (!bpic) {
      b.u(1, 0);  // adaptive_ref_pic_marking_mode_flag (sliding window)
    }
    b.se(0);  // slice_qp_delta
    b.ue(1);  // disable_deblocking_filter_idc = 1 (off)
    if (skip) {
      // mb_skip_run covering the whole picture.  P_Skip copies the previous picture; B_Skip with
      // spatial direct prediction (no coded neighbours, intra co-located picture) is bi-predicted
      // with zero motion: (anchor_before + anchor_after + 1) >> 1 per sample.
      b.ue((u32)(wmb * hmb));
    } else {
      fill(f, ysrc.data(), usrc.data(), vsrc.data());
      // pad the display planes to the coded size by edge replication
      for (i32 r = 0; r < ch; ++r) {
        const u8* s = ysrc.data() + (size_t)std::min(r, height - 1) * width;
        u8* d = y.data() + (size_t)r * cw;
        memcpy(d, s, width);
        for (i32 c = width; c < cw; ++c) d[c] = s[width - 1];
      }
      const i32 w2 = width / 2, h2 = height / 2, cw2 = cw / 2, ch2 = ch / 2;
      for (i32 r = 0; r < ch2; ++r) {
        const size_t so = (size_t)std::min(r, h2 - 1) * w2;
        u8* du = u.data() + (size_t)r * cw2;
        u8* dv = v.data() + (size_t)r * cw2;
        memcpy(du, usrc.data() + so, w2);
        memcpy(dv, vsrc.data() + so, w2);
        for (i32 c = w2; c < cw2; ++c) {
          du[c] = usrc[so + w2 - 1];
          dv[c] = vsrc[so + w2 - 1];
        }
      }
      u8 mb[384];
      for (i32 my = 0; my < hmb; ++my)
        for (i32 mx = 0; mx < wmb; ++mx) {
          if (!key) b.ue(0);        // mb_skip_run = 0 before every coded macroblock of a P slice
          b.ue(key ? 25 : 30);      // mb_type I_PCM (25 in I slices, 5 + 25 in P slices)
          b.align_zero();           // pcm_alignment_zero_bit
          for (i32 r = 0; r < 16; ++r) memcpy(mb + r * 16, y.data() + (size_t)(my * 16 + r) * cw + mx * 16, 16);
          for (i32 r = 0; r < 8; ++r) {
            memcpy(mb + 256 + r * 8, u.data() + (size_t)(my * 8 + r) * cw2 + mx * 8, 8);
            memcpy(mb + 320 + r * 8, v.data() + (size_t)(my * 8 + r) * cw2 + mx * 8, 8);
          }
          b.bytes(mb, 384);
        }
    }
    b.trailing();
    emit_nal(out, key ? 3 : (bpic ? 0 : 2), key ? 5 : 1, b.data());
  }
}

}  // namespace internal
}  // namespace scanner
