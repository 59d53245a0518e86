// mp4.cpp -- see mp4.h.
#include "mp4.h"

#include <cstring>
#include <functional>

namespace scanner {
namespace internal {
namespace {

// ---- big-endian readers over a bounded view -----------------------------------------------------
struct View {
  const u8* p = nullptr;
  size_t n = 0;
  bool has(size_t off, size_t len) const { return off <= n && len <= n - off; }
  u32 be16(size_t o) const { return ((u32)p[o] << 8) | p[o + 1]; }
  u32 be32(size_t o) const { return ((u32)p[o] << 24) | ((u32)p[o + 1] << 16) | ((u32)p[o + 2] << 8) | p[o + 3]; }
  u64 be64(size_t o) const { return ((u64)be32(o) << 32) | be32(o + 4); }
  View sub(size_t off, size_t len) const { return View{p + off, len}; }
};

constexpr u32 fourcc(const char (&s)[5]) {
  return ((u32)(u8)s[0] << 24) | ((u32)(u8)s[1] << 16) | ((u32)(u8)s[2] << 8) | (u32)(u8)s[3];
}

// Calls fn(type, payload) for every box directly inside `v`; stops early when fn returns false.
// Returns false on a malformed box header.
bool for_each_box(const View& v, const std::function<bool(u32, const View&)>& fn) {
  size_t off = 0;
  while (off + 8 <= v.n) {
    u64 size = v.be32(off);
    const u32 type = v.be32(off + 4);
    size_t hdr = 8;
    if (size == 1) {
      if (!v.has(off, 16)) return false;
      size = v.be64(off + 8);
      hdr = 16;
    } else if (size == 0) {
      size = v.n - off;  // box extends to the end of its container
    }
    if (size < hdr || size > v.n - off) return false;
    if (!fn(type, v.sub(off + hdr, (size_t)size - hdr))) return true;
    off += (size_t)size;
  }
  return true;
}

bool find_box(const View& v, u32 want, View& out) {
  bool found = false;
  for_each_box(v, [&](u32 t, const View& b) {
    if (t == want) {
      out = b;
      found = true;
      return false;
    }
    return true;
  });
  return found;
}

struct AvcConfig {
  int nal_length_size = 4;
  std::vector<std::vector<u8>> sps, pps;
};

bool parse_avcc(const View& b, AvcConfig& c) {
  if (b.n < 7 || b.p[0] != 1) return false;
  c.nal_length_size = (b.p[4] & 3) + 1;
  size_t off = 5;
  const int nsps = b.p[off++] & 31;
  for (int i = 0; i < nsps; ++i) {
    if (!b.has(off, 2)) return false;
    const size_t len = b.be16(off);
    off += 2;
    if (!b.has(off, len)) return false;
    c.sps.emplace_back(b.p + off, b.p + off + len);
    off += len;
  }
  if (!b.has(off, 1)) return false;
  const int npps = b.p[off++];
  for (int i = 0; i < npps; ++i) {
    if (!b.has(off, 2)) return false;
    const size_t len = b.be16(off);
    off += 2;
    if (!b.has(off, len)) return false;
    c.pps.emplace_back(b.p + off, b.p + off + len);
    off += len;
  }
  return !c.sps.empty() && !c.pps.empty();
}

const u8 kStartCode[4] = {0, 0, 0, 1};

void put_nal(std::vector<u8>& out, const u8* nal, size_t len) {
  out.insert(out.end(), kStartCode, kStartCode + 4);
  out.insert(out.end(), nal, nal + len);
}

// ---- writer helpers -----------------------------------------------------------------------------
struct Box {
  std::vector<u8> b;
  void u8_(u32 v) { b.push_back((u8)v); }
  void be16(u32 v) {
    b.push_back((u8)(v >> 8));
    b.push_back((u8)v);
  }
  void be32(u32 v) {
    be16(v >> 16);
    be16(v & 0xFFFF);
  }
  void be64(u64 v) {
    be32((u32)(v >> 32));
    be32((u32)v);
  }
  void tag(const char* s) { b.insert(b.end(), s, s + 4); }
  void zeros(size_t n) { b.insert(b.end(), n, 0); }
  void bytes(const std::vector<u8>& v) { b.insert(b.end(), v.begin(), v.end()); }
  // wraps `payload` into a box of `type`
  static std::vector<u8> wrap(const char* type, const std::vector<u8>& payload) {
    Box o;
    o.be32((u32)(payload.size() + 8));
    o.tag(type);
    o.bytes(payload);
    return o.b;
  }
  static std::vector<u8> full(const char* type, u32 version_flags, const std::vector<u8>& payload) {
    Box o;
    o.be32(version_flags);
    o.bytes(payload);
    return wrap(type, o.b);
  }
};

// NAL units of one Annex-B access unit: (pointer, length) without start codes
void split_nals(const u8* p, size_t n, std::vector<std::pair<const u8*, size_t>>& out) {
  size_t i = 0;
  auto start_at = [&](size_t k) -> int {  // length of a start code at k, 0 if none
    if (k + 3 <= n && p[k] == 0 && p[k + 1] == 0 && p[k + 2] == 1) return 3;
    if (k + 4 <= n && p[k] == 0 && p[k + 1] == 0 && p[k + 2] == 0 && p[k + 3] == 1) return 4;
    return 0;
  };
  while (i < n && !start_at(i)) ++i;
  while (i < n) {
    const int sc = start_at(i);
    if (!sc) break;
    const size_t b = i + sc;
    size_t e = b;
    while (e < n && !start_at(e)) ++e;
    size_t len = e - b;
    while (len > 0 && e < n && p[b + len - 1] == 0) --len;  // trailing_zero_8bits belong to the next start code
    if (len) out.emplace_back(p + b, len);
    i = e;
  }
}

}  // namespace

bool looks_like_mp4(const u8* data, size_t size) {
  if (size < 12) return false;
  View v{data, size};
  const u32 t = v.be32(4);
  return t == fourcc("ftyp") || t == fourcc("moov") || t == fourcc("mdat") || t == fourcc("free") ||
         t == fourcc("skip") || t == fourcc("wide") || t == fourcc("styp");
}

Result demux_mp4(const u8* data, size_t size, Mp4Track& out) {
  Result r;
  r.set_success(true);
  View file{data, size};
  View moov;
  bool fragmented = false, ok_boxes;
  ok_boxes = for_each_box(file, [&](u32 t, const View&) {
    if (t == fourcc("moof")) fragmented = true;
    return true;
  });
  if (!find_box(file, fourcc("moov"), moov)) {
    if (!ok_boxes) {
      RESULT_ERROR(&r, "mp4: a top-level box runs past the end of the file before any moov box (truncated file?)");
      return r;
    }
    RESULT_ERROR(&r, "mp4: no moov box (not an ISO base media file, or the index is missing)");
    return r;
  }
  if (fragmented) {
    RESULT_ERROR(&r, "mp4: fragmented files (moof) are not supported");
    return r;
  }

  std::string why = "mp4: no H.264 (avc1/avc3) video track";
  bool done = false;
  for_each_box(moov, [&](u32 t, const View& trak) {
    if (t != fourcc("trak") || done) return true;
    View mdia, hdlr, mdhd, minf, stbl, stsd, stsz, stsc, stco, stss;
    if (!find_box(trak, fourcc("mdia"), mdia) || !find_box(mdia, fourcc("hdlr"), hdlr) || hdlr.n < 12 ||
        hdlr.be32(8) != fourcc("vide"))
      return true;
    if (!find_box(mdia, fourcc("mdhd"), mdhd) || !find_box(mdia, fourcc("minf"), minf) ||
        !find_box(minf, fourcc("stbl"), stbl) || !find_box(stbl, fourcc("stsd"), stsd) ||
        !find_box(stbl, fourcc("stsz"), stsz) || !find_box(stbl, fourcc("stsc"), stsc)) {
      why = "mp4: video track without mdhd/stbl/stsd/stsz/stsc";
      return true;
    }
    const bool co64 = !find_box(stbl, fourcc("stco"), stco);
    if (co64 && !find_box(stbl, fourcc("co64"), stco)) {
      why = "mp4: video track without a chunk offset table";
      return true;
    }
    const bool has_stss = find_box(stbl, fourcc("stss"), stss);

    // mdhd: version 0 = 32-bit times, version 1 = 64-bit
    if (mdhd.n >= 24 && mdhd.p[0] == 0) {
      out.timescale = mdhd.be32(12);
      out.duration = mdhd.be32(16);
    } else if (mdhd.n >= 36 && mdhd.p[0] == 1) {
      out.timescale = mdhd.be32(20);
      out.duration = mdhd.be64(24);
    }

    // stsd: first sample entry must be AVC
    if (stsd.n < 16) {
      why = "mp4: truncated stsd";
      return true;
    }
    const View entry_area = stsd.sub(8, stsd.n - 8);
    AvcConfig cfg;
    bool avc = false;
    std::string codec;
    for_each_box(entry_area, [&](u32 et, const View& e) {
      codec.clear();
      for (int sh = 24; sh >= 0; sh -= 8) {  // printable form of the fourcc for the error message
        const u8 ch = (u8)(et >> sh);
        codec.push_back(ch >= 0x20 && ch < 0x7F ? (char)ch : '?');
      }
      if ((et == fourcc("avc1") || et == fourcc("avc3")) && e.n >= 78) {
        out.width = (i32)e.be16(24);
        out.height = (i32)e.be16(26);
        View avcc;
        if (find_box(e.sub(78, e.n - 78), fourcc("avcC"), avcc) && parse_avcc(avcc, cfg)) avc = true;
      }
      return false;  // only the first entry
    });
    if (!avc) {
      why = "mp4: video track codec '" + codec + "' is not H.264 with an avcC box";
      return true;
    }

    // sample sizes
    if (stsz.n < 12) {
      why = "mp4: truncated stsz";
      return true;
    }
    const u32 fixed_size = stsz.be32(4), nsamples = stsz.be32(8);
    if (fixed_size == 0 && !stsz.has(12, (size_t)nsamples * 4)) {
      why = "mp4: truncated stsz";
      return true;
    }
    // chunk offsets
    if (stco.n < 8) {
      why = "mp4: truncated chunk offset table";
      return true;
    }
    const u32 nchunks = stco.be32(4);
    if (!stco.has(8, (size_t)nchunks * (co64 ? 8 : 4))) {
      why = "mp4: truncated chunk offset table";
      return true;
    }
    // sample-to-chunk runs
    if (stsc.n < 8) {
      why = "mp4: truncated stsc";
      return true;
    }
    const u32 nruns = stsc.be32(4);
    if (!stsc.has(8, (size_t)nruns * 12)) {
      why = "mp4: truncated stsc";
      return true;
    }
    // sync samples (1-based); without the box every sample is a sync sample
    std::vector<bool> sync(nsamples, !has_stss);
    if (has_stss) {
      if (stss.n < 8 || !stss.has(8, (size_t)stss.be32(4) * 4)) {
        why = "mp4: truncated stss";
        return true;
      }
      for (u32 i = 0; i < stss.be32(4); ++i) {
        const u32 s = stss.be32(8 + (size_t)i * 4);
        if (s >= 1 && s <= nsamples) sync[s - 1] = true;
      }
    }

    // walk chunks -> samples
    out.annexb.clear();
    u32 sample = 0;
    for (u32 run = 0; run < nruns && sample < nsamples; ++run) {
      const u32 first = stsc.be32(8 + (size_t)run * 12), per = stsc.be32(12 + (size_t)run * 12);
      const u32 next_first = run + 1 < nruns ? stsc.be32(8 + (size_t)(run + 1) * 12) : nchunks + 1;
      if (first < 1 || next_first < first) {
        why = "mp4: inconsistent stsc";
        return true;
      }
      for (u32 chunk = first; chunk < next_first && chunk <= nchunks && sample < nsamples; ++chunk) {
        u64 off = co64 ? stco.be64(8 + (size_t)(chunk - 1) * 8) : stco.be32(8 + (size_t)(chunk - 1) * 4);
        for (u32 k = 0; k < per && sample < nsamples; ++k, ++sample) {
          const u64 ssize = fixed_size ? fixed_size : stsz.be32(12 + (size_t)sample * 4);
          if (off > size || ssize > size - off) {
            why = "mp4: sample " + std::to_string(sample) + " lies outside the file";
            return true;
          }
          // length-prefixed NAL units -> Annex-B
          const u8* sp = data + off;
          bool has_sps = false;
          {
            u64 q = 0;
            while (q + cfg.nal_length_size <= ssize) {
              u64 len = 0;
              for (int b = 0; b < cfg.nal_length_size; ++b) len = (len << 8) | sp[q + b];
              q += cfg.nal_length_size;
              if (len > ssize - q) break;
              if (len && (sp[q] & 31) == 7) has_sps = true;
              q += len;
            }
          }
          if (sync[sample] && !has_sps) {
            for (auto& s : cfg.sps) put_nal(out.annexb, s.data(), s.size());
            for (auto& p : cfg.pps) put_nal(out.annexb, p.data(), p.size());
          }
          u64 q = 0;
          while (q + cfg.nal_length_size <= ssize) {
            u64 len = 0;
            for (int b = 0; b < cfg.nal_length_size; ++b) len = (len << 8) | sp[q + b];
            q += cfg.nal_length_size;
            if (len > ssize - q) {
              why = "mp4: NAL unit of sample " + std::to_string(sample) + " overruns the sample";
              return true;
            }
            if (len) put_nal(out.annexb, sp + q, (size_t)len);
            q += len;
          }
          if (sync[sample]) ++out.sync_samples;
          off += ssize;
        }
      }
    }
    if (sample != nsamples) {
      why = "mp4: the chunk tables describe " + std::to_string(sample) + " of " + std::to_string(nsamples) + " samples";
      return true;
    }
    out.samples = nsamples;
    done = true;
    return false;
  });
  if (!done) {
    RESULT_ERROR(&r, "%s", why.c_str());
  }
  return r;
}

Result mux_mp4(const u8* annexb, size_t size, const H264Index& index, i32 fps_num, i32 fps_den,
               std::vector<u8>& out) {
  Result r;
  r.set_success(true);
  if (index.frames() == 0 || fps_num <= 0 || fps_den <= 0) {
    RESULT_ERROR(&r, "mux_mp4: empty stream or bad frame rate");
    return r;
  }
  // parameter sets for avcC: from the index's metadata packets
  std::vector<std::pair<const u8*, size_t>> meta;
  split_nals(index.metadata_packets.data(), index.metadata_packets.size(), meta);
  std::vector<u8> sps, pps;
  for (auto& m : meta) {
    if ((m.first[0] & 31) == 7 && sps.empty()) sps.assign(m.first, m.first + m.second);
    if ((m.first[0] & 31) == 8 && pps.empty()) pps.assign(m.first, m.first + m.second);
  }
  if (sps.size() < 4 || pps.empty()) {
    RESULT_ERROR(&r, "mux_mp4: the stream has no SPS/PPS");
    return r;
  }

  // ---- mdat payload: every access unit as 4-byte-length-prefixed NALs, parameter sets and
  // access unit delimiters dropped (they live in avcC)
  const i64 n = index.frames();
  std::vector<u8> mdat;
  std::vector<u32> sample_size((size_t)n);
  for (i64 f = 0; f < n; ++f) {
    const u64 off = index.sample_offsets[(size_t)f], sz = index.sample_sizes[(size_t)f];
    if (off > size || sz > size - off) {
      RESULT_ERROR(&r, "mux_mp4: sample %ld lies outside the stream", (long)f);
      return r;
    }
    std::vector<std::pair<const u8*, size_t>> nals;
    split_nals(annexb + off, (size_t)sz, nals);
    const size_t before = mdat.size();
    for (auto& nal : nals) {
      const int t = nal.first[0] & 31;
      if (t == 7 || t == 8 || t == 9) continue;
      const u32 len = (u32)nal.second;
      const u8 be[4] = {(u8)(len >> 24), (u8)(len >> 16), (u8)(len >> 8), (u8)len};
      mdat.insert(mdat.end(), be, be + 4);
      mdat.insert(mdat.end(), nal.first, nal.first + nal.second);
    }
    sample_size[(size_t)f] = (u32)(mdat.size() - before);
  }

  // ---- ftyp
  Box ftyp;
  ftyp.tag("isom");
  ftyp.be32(0x200);
  ftyp.tag("isom");
  ftyp.tag("iso2");
  ftyp.tag("avc1");
  ftyp.tag("mp41");
  const std::vector<u8> ftyp_box = Box::wrap("ftyp", ftyp.b);
  const bool big = mdat.size() + 16 > 0xFFFFFFFFull;
  const u64 mdat_hdr = big ? 16 : 8;
  const u64 first_sample_off = ftyp_box.size() + mdat_hdr;

  // ---- sample tables
  const u32 timescale = (u32)fps_num, delta = (u32)fps_den;  // one sample lasts fps_den / fps_num s
  const u64 duration = (u64)n * delta;
  Box stts;  // one run: n samples of `delta`
  stts.be32(1);
  stts.be32((u32)n);
  stts.be32(delta);
  Box stss;
  stss.be32((u32)index.keyframe_indices.size());
  for (i64 k : index.keyframe_indices) stss.be32((u32)k + 1);
  Box stsc;  // one chunk holding every sample
  stsc.be32(1);
  stsc.be32(1);
  stsc.be32((u32)n);
  stsc.be32(1);
  Box stsz;
  stsz.be32(0);
  stsz.be32((u32)n);
  for (u32 s : sample_size) stsz.be32(s);
  Box co;
  co.be32(1);
  const bool co64 = first_sample_off > 0xFFFFFFFFull;
  if (co64) co.be64(first_sample_off);
  else co.be32((u32)first_sample_off);

  Box avcc;
  avcc.u8_(1);
  avcc.u8_(sps[1]);
  avcc.u8_(sps[2]);
  avcc.u8_(sps[3]);
  avcc.u8_(0xFC | 3);  // 4-byte NAL lengths
  avcc.u8_(0xE0 | 1);
  avcc.be16((u32)sps.size());
  avcc.bytes(sps);
  avcc.u8_(1);
  avcc.be16((u32)pps.size());
  avcc.bytes(pps);

  Box avc1;  // VisualSampleEntry
  avc1.zeros(6);
  avc1.be16(1);  // data_reference_index
  avc1.zeros(16);
  avc1.be16((u32)index.width);
  avc1.be16((u32)index.height);
  avc1.be32(0x00480000);  // 72 dpi
  avc1.be32(0x00480000);
  avc1.be32(0);
  avc1.be16(1);   // frame_count
  avc1.zeros(32);  // compressorname
  avc1.be16(0x18);
  avc1.be16(0xFFFF);
  avc1.bytes(Box::wrap("avcC", avcc.b));
  Box stsd;
  stsd.be32(1);
  stsd.bytes(Box::wrap("avc1", avc1.b));

  Box stbl;
  stbl.bytes(Box::full("stsd", 0, stsd.b));
  stbl.bytes(Box::full("stts", 0, stts.b));
  stbl.bytes(Box::full("stss", 0, stss.b));
  stbl.bytes(Box::full("stsc", 0, stsc.b));
  stbl.bytes(Box::full("stsz", 0, stsz.b));
  stbl.bytes(Box::full(co64 ? "co64" : "stco", 0, co.b));

  Box url;  // self-contained data reference
  Box dref;
  dref.be32(1);
  dref.bytes(Box::full("url ", 1, url.b));
  Box dinf;
  dinf.bytes(Box::full("dref", 0, dref.b));
  Box vmhd;
  vmhd.zeros(8);
  Box minf;
  minf.bytes(Box::full("vmhd", 1, vmhd.b));
  minf.bytes(Box::wrap("dinf", dinf.b));
  minf.bytes(Box::wrap("stbl", stbl.b));

  Box mdhd;
  mdhd.be32(0);
  mdhd.be32(0);
  mdhd.be32(timescale);
  mdhd.be32((u32)duration);
  mdhd.be16(0x55C4);  // language: und
  mdhd.be16(0);
  Box hdlr;
  hdlr.be32(0);
  hdlr.tag("vide");
  hdlr.zeros(12);
  const char hname[] = "VideoHandler";
  hdlr.b.insert(hdlr.b.end(), hname, hname + sizeof(hname));
  Box mdia;
  mdia.bytes(Box::full("mdhd", 0, mdhd.b));
  mdia.bytes(Box::full("hdlr", 0, hdlr.b));
  mdia.bytes(Box::wrap("minf", minf.b));

  static const u32 kUnity[9] = {0x00010000, 0, 0, 0, 0x00010000, 0, 0, 0, 0x40000000};
  Box tkhd;
  tkhd.be32(0);
  tkhd.be32(0);
  tkhd.be32(1);  // track id
  tkhd.be32(0);
  tkhd.be32((u32)duration);
  tkhd.zeros(8);
  tkhd.be16(0);
  tkhd.be16(0);
  tkhd.be16(0);
  tkhd.be16(0);
  for (u32 m : kUnity) tkhd.be32(m);
  tkhd.be32((u32)index.width << 16);
  tkhd.be32((u32)index.height << 16);
  Box trak;
  trak.bytes(Box::full("tkhd", 3, tkhd.b));  // enabled | in movie
  trak.bytes(Box::wrap("mdia", mdia.b));

  Box mvhd;
  mvhd.be32(0);
  mvhd.be32(0);
  mvhd.be32(timescale);
  mvhd.be32((u32)duration);
  mvhd.be32(0x00010000);  // rate
  mvhd.be16(0x0100);      // volume
  mvhd.zeros(10);
  for (u32 m : kUnity) mvhd.be32(m);
  mvhd.zeros(24);
  mvhd.be32(2);  // next track id
  Box moov;
  moov.bytes(Box::full("mvhd", 0, mvhd.b));
  moov.bytes(Box::wrap("trak", trak.b));
  const std::vector<u8> moov_box = Box::wrap("moov", moov.b);

  // ---- assemble: ftyp, mdat, moov
  out.clear();
  out.reserve(ftyp_box.size() + mdat_hdr + mdat.size() + moov_box.size());
  out.insert(out.end(), ftyp_box.begin(), ftyp_box.end());
  Box mh;
  if (big) {
    mh.be32(1);
    mh.tag("mdat");
    mh.be64(mdat.size() + 16);
  } else {
    mh.be32((u32)(mdat.size() + 8));
    mh.tag("mdat");
  }
  out.insert(out.end(), mh.b.begin(), mh.b.end());
  out.insert(out.end(), mdat.begin(), mdat.end());
  out.insert(out.end(), moov_box.begin(), moov_box.end());
  return r;
}

}  // namespace internal
}  // namespace scanner
