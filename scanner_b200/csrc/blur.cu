// blur.cu -- integer box filter on u8 HWC 3-channel frames (sm_100a).
// Replaces BlurKernel::execute (reference tests/test_ops.cpp:265-294): for interior pixels
// out = (sum of the (fl+fr+1)^2 window) / (fl+fr+1)^2 with fl = ceil(k/2)-1, fr = k/2; the
// reference never writes the border of its freshly allocated frame -- here it is written as 0.
// Integer, bit-exact on the interior.
//
//   box3_kernel   kernel_size 3 (BASELINE configs[2]).  A thread owns one 32-bit column of the byte
//                 image (4 bytes = 1 1/3 pixels) and walks down a strip of rows.  The three taps of
//                 the horizontal pass are byte-shifted views of three neighbouring words (funnel
//                 shifts), summed as packed u16x2 lanes; the vertical pass is a 3-row sliding sum
//                 held in registers; /9 is an exact multiply-shift; four results are stored as one
//                 32-bit word.  No shared memory: neighbouring threads' loads overlap in L1.
//   box_stream_kernel    any kernel_size 2..31 on 16-byte aligned rows: bulk-async (cp.async.bulk + mbarrier) row
//                 ring, sliding horizontal sums, running vertical sums: O(1) work per sample.
//   box_generic_kernel   the rest (unaligned or tiny frames, kernel_size 1): shared-memory tile, direct sums.
#include "scn_common.cuh"

namespace scn {
namespace {

// ---------------------------------------------------------------------------------------------
constexpr int B3_THREADS = 256;
constexpr int B3_ROWS = 64;  // output rows per CTA strip

// packed u16x2 horizontal sums of the 4 bytes of a word column: even = bytes 0,2 ; odd = bytes 1,3
struct H3 {
  uint32_t even, odd;
};

__device__ __forceinline__ H3 hsum3(uint32_t wm, uint32_t w0, uint32_t wp) {
  // byte views: A = bytes [x-3 .. x], B = [x .. x+3], C = [x+3 .. x+6]
  const uint32_t A = __funnelshift_r(wm, w0, 8);
  const uint32_t C = __funnelshift_r(w0, wp, 24);
  H3 h;
  h.even = (A & 0x00FF00FFu) + (w0 & 0x00FF00FFu) + (C & 0x00FF00FFu);
  h.odd = prmt(A, 0u, 0x4341u) + prmt(w0, 0u, 0x4341u) + prmt(C, 0u, 0x4341u);
  return h;
}

// floor(v / 9) for v <= 2295 (3*3*255): (v * 7282) >> 16, exact (checked exhaustively in the tests)
__device__ __forceinline__ uint32_t div9_pack(uint32_t even, uint32_t odd) {
  const uint32_t e0 = ((even & 0xFFFFu) * 7282u) >> 16, e1 = __umulhi(even & 0xFFFF0000u, 7282u);
  const uint32_t o0 = ((odd & 0xFFFFu) * 7282u) >> 16, o1 = __umulhi(odd & 0xFFFF0000u, 7282u);
  return e0 | (o0 << 8) | (e1 << 16) | (o1 << 24);
}

// words_per_row = width*3/4 (requires width % 4 == 0 and 4-byte aligned frames)
__global__ void __launch_bounds__(B3_THREADS)
box3_kernel(PtrBatch src, MutPtrBatch dst, int width, int height, int words_per_row) {
  const int wx = blockIdx.x * B3_THREADS + threadIdx.x;
  if (wx >= words_per_row) return;
  const uint32_t* __restrict__ s = reinterpret_cast<const uint32_t*>(src.p[blockIdx.z]);
  uint32_t* __restrict__ d = reinterpret_cast<uint32_t*>(dst.p[blockIdx.z]);
  const int y0 = blockIdx.y * B3_ROWS;
  const int y1 = min(height, y0 + B3_ROWS);

  // bytes of this word that are interior in x: pixel = byte / 3 must be in [1, width-2]
  uint32_t xmask = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int px = (wx * 4 + b) / 3;
    if (px >= 1 && px <= width - 2) xmask |= 0xFFu << (8 * b);
  }
  const bool has_m = wx > 0, has_p = wx + 1 < words_per_row;

  auto load_h = [&](int y) -> H3 {
    H3 h{0u, 0u};
    if (y >= 0 && y < height) {
      const uint32_t* row = s + (size_t)y * words_per_row + wx;
      const uint32_t w0 = __ldg(row);
      const uint32_t wm = has_m ? __ldg(row - 1) : 0u;
      const uint32_t wp = has_p ? __ldg(row + 1) : 0u;
      h = hsum3(wm, w0, wp);
    }
    return h;
  };

  H3 a = load_h(y0 - 1), b = load_h(y0);
  for (int y = y0; y < y1; ++y) {
    const H3 c = load_h(y + 1);
    uint32_t out = 0;
    if (y >= 1 && y <= height - 2) out = div9_pack(a.even + b.even + c.even, a.odd + b.odd + c.odd) & xmask;
    d[(size_t)y * words_per_row + wx] = out;
    a = b;
    b = c;
  }
}

// ---------------------------------------------------------------------------------------------
constexpr int TW = 64;   // tile width in pixels
constexpr int TH = 32;   // tile height in pixels
constexpr int KMAX = 31;
constexpr int BT = 256;  // 8 warps: warp <-> row, lane <-> byte

__global__ void __launch_bounds__(BT)
box_generic_kernel(PtrBatch src, MutPtrBatch dst, int width, int height, int fl, int fr) {
  extern __shared__ uint8_t smem_raw[];
  const int k = fl + fr + 1;
  const int in_w = TW + k - 1, in_h = TH + k - 1;
  const int in_wb = in_w * 3, out_wb = TW * 3;
  uint8_t* tile = smem_raw;                                                              // in_h x in_wb
  uint16_t* hsum = reinterpret_cast<uint16_t*>(smem_raw + ((in_h * in_wb + 15) & ~15));  // in_h x out_wb
  const uint8_t* __restrict__ s = src.p[blockIdx.z];
  uint8_t* __restrict__ d = dst.p[blockIdx.z];
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_bytes = width * 3;

  for (int r = warp; r < in_h; r += BT / 32) {
    const int sy = y0 - fl + r;
    const bool row_ok = sy >= 0 && sy < height;
    const uint8_t* srow = s + (size_t)(row_ok ? sy : 0) * row_bytes;
    const int xb0 = (x0 - fl) * 3;
    for (int c = lane; c < in_wb; c += 32) {
      const int sxb = xb0 + c;
      tile[r * in_wb + c] = (row_ok && sxb >= 0 && sxb < row_bytes) ? __ldg(srow + sxb) : (uint8_t)0;
    }
  }
  __syncthreads();
  for (int r = warp; r < in_h; r += BT / 32) {
    const uint8_t* trow = tile + r * in_wb;
    for (int c = lane; c < out_wb; c += 32) {
      uint32_t acc = 0;
      for (int j = 0; j < k; ++j) acc += trow[c + j * 3];
      hsum[r * out_wb + c] = (uint16_t)acc;  // <= 31 * 255
    }
  }
  __syncthreads();
  const uint32_t div = (uint32_t)(k * k);
  for (int r = warp; r < TH; r += BT / 32) {
    const int y = y0 + r;
    if (y >= height) break;
    const bool row_in = y >= fl && y < height - fr;
    for (int c = lane; c < out_wb; c += 32) {
      const int xb = x0 * 3 + c;
      if (xb >= row_bytes) break;
      uint32_t acc = 0;
      for (int j = 0; j < k; ++j) acc += hsum[(r + j) * out_wb + c];
      const int x = xb / 3;
      const bool interior = row_in && x >= fl && x < width - fr;
      d[(size_t)y * row_bytes + xb] = interior ? (uint8_t)(acc / div) : (uint8_t)0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// box_stream_kernel: any kernel_size, O(1) work per sample.  A CTA owns a strip of kStripBytes byte
// columns and walks down kStripRows output rows:
//   * each input row segment arrives by ONE bulk-async copy (cp.async.bulk global -> shared, completion on
//     an mbarrier) into a ring of kRowSlots slots, issued kRowSlots - 1 rows ahead by one thread;
//   * horizontal pass: a thread owns 4 pixels (12 bytes); the first pixel's window is summed directly, the
//     next three slide (add the entering sample, drop the leaving one);
//   * vertical pass: running column sums in registers: + this row's horizontal sums, - those of the row
//     that leaves the window (kept in a shared-memory ring of k rows of u16);
//   * out = sum / k^2 by an exact multiply-shift, 12 bytes stored per thread (a warp writes 384 contiguous
//     bytes).  Border samples (window not inside the frame) are written as 0, so whatever a row segment
//     holds beyond the frame never reaches an output.
// Needs 16-byte aligned rows (width % 16 == 0, 16-byte aligned frames).
constexpr int kStripBytes = 3072;  // 1024 pixels: 256 threads x 12 bytes
constexpr int kStripRows = 128;
constexpr int kRowSlots = 4;
constexpr int BS = 256;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(BS)
box_stream_kernel(PtrBatch src, MutPtrBatch dst, int width, int height, int fl, int fr, uint32_t div_magic) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long bars[kRowSlots];
  const int k = fl + fr + 1;
  const int row_bytes = width * 3;
  const int x0b = blockIdx.x * kStripBytes;                       // first output byte column of the strip
  const int out_wb = min(kStripBytes, row_bytes - x0b);
  // bytes of a row this strip reads: [c0, c1), widened to 16-byte boundaries, clipped to the row
  const int c0 = max(0, (x0b - 3 * fl) & ~15);
  const int c1 = min(row_bytes, (x0b + out_wb + 3 * fr + 15) & ~15);
  const int seg = c1 - c0;
  const int lead = x0b - c0;                                      // strip byte 0 sits at segment byte `lead`
  // a slot: kLeftPad bytes nobody writes (reads of a border pixel's window may fall before the segment),
  // then the segment: at most kStripBytes + 2 * (45 + 15) bytes (KMAX = 31: fl, fr <= 15)
  constexpr int kLeftPad = 48;
  constexpr int kSegMax = kLeftPad + kStripBytes + 128;
  uint8_t* rows = smem_raw;                                       // kRowSlots x kSegMax
  uint16_t* hring = reinterpret_cast<uint16_t*>(smem_raw + kRowSlots * kSegMax);  // k x kStripBytes
  const uint8_t* __restrict__ s = src.p[blockIdx.z];
  uint8_t* __restrict__ d = dst.p[blockIdx.z];
  const int y_out0 = blockIdx.y * kStripRows, y_out1 = min(height, y_out0 + kStripRows);
  // input rows needed: [y_out0 - fl, y_out1 + fr), clipped; rows outside the frame only feed border outputs
  const int r_first = max(0, y_out0 - fl), r_last = min(height, y_out1 + fr);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);
  const int t = threadIdx.x;
  if (t == 0) {
    for (int i = 0; i < kRowSlots; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int r) {  // thread 0: bulk copy of input row r into its slot
    const int slot = (r - r_first) % kRowSlots;
    mbar_expect_tx(bar0 + 8 * slot, (uint32_t)seg);
    bulk_g2s(rows_s + slot * kSegMax + kLeftPad, s + (size_t)r * row_bytes + c0, (uint32_t)seg, bar0 + 8 * slot);
  };
  if (t == 0)
    for (int r = r_first; r < min(r_last, r_first + kRowSlots - 1); ++r) issue(r);

  const int b0 = 12 * t;                       // this thread's first output byte in the strip
  const bool active = b0 < out_wb;             // out_wb is a multiple of 12 (row_bytes % 48 == 0, strips of 3072)
  uint32_t vsum[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) vsum[i] = 0;
  // x-interior flags of the 4 pixels
  bool xin[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int px = (x0b + b0) / 3 + p;
    xin[p] = px >= fl && px < width - fr;
  }

  for (int r = r_first; r < r_last; ++r) {
    const int idx = r - r_first, slot = idx % kRowSlots;
    if (t == 0 && r + kRowSlots - 1 < r_last) issue(r + kRowSlots - 1);   // its slot was released by the last barrier
    mbar_wait(bar0 + 8 * slot, (uint32_t)((idx / kRowSlots) & 1));
    uint16_t h[12];
    if (active) {
      const uint8_t* row = rows + slot * kSegMax + kLeftPad + lead + b0;
      // window of pixel 0 summed directly: samples at byte offsets 3 * (j - fl) + ch (a border pixel's window
      // may start in the pad or end in stale bytes: its output is masked); pixels 1..3 slide
      uint32_t a0 = 0, a1 = 0, a2 = 0;
      const int lo = -3 * fl;
      for (int j = 0; j < k; ++j) {
        a0 += row[lo + 3 * j + 0];
        a1 += row[lo + 3 * j + 1];
        a2 += row[lo + 3 * j + 2];
      }
      h[0] = (uint16_t)a0;
      h[1] = (uint16_t)a1;
      h[2] = (uint16_t)a2;
#pragma unroll
      for (int p = 1; p < 4; ++p) {
        a0 += row[3 * (p + fr) + 0] - row[3 * (p - 1 - fl) + 0];
        a1 += row[3 * (p + fr) + 1] - row[3 * (p - 1 - fl) + 1];
        a2 += row[3 * (p + fr) + 2] - row[3 * (p - 1 - fl) + 2];
        h[3 * p + 0] = (uint16_t)a0;
        h[3 * p + 1] = (uint16_t)a1;
        h[3 * p + 2] = (uint16_t)a2;
      }
      // vertical running sums: + this row, - the row leaving the window (input row r - k, if it was fed);
      // the ring holds 12 u16 per thread = three 8-byte words
      uint2* hslot = reinterpret_cast<uint2*>(hring + (size_t)(idx % k) * kStripBytes + b0);
      if (idx >= k) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint2 o = hslot[q];
          vsum[4 * q + 0] -= o.x & 0xFFFFu;
          vsum[4 * q + 1] -= o.x >> 16;
          vsum[4 * q + 2] -= o.y & 0xFFFFu;
          vsum[4 * q + 3] -= o.y >> 16;
        }
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) vsum[i] += h[i];
#pragma unroll
      for (int q = 0; q < 3; ++q)
        hslot[q] = make_uint2((uint32_t)h[4 * q] | ((uint32_t)h[4 * q + 1] << 16), (uint32_t)h[4 * q + 2] | ((uint32_t)h[4 * q + 3] << 16));
      // output row y = r - fr is complete once input row r is in (its window is rows y-fl .. y+fr)
      const int y = r - fr;
      if (y >= y_out0 && y < y_out1) {
        const bool yin = y >= fl && y < height - fr;
        uint32_t o[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) o[i] = (yin && xin[i / 3]) ? __umulhi(vsum[i], div_magic) : 0u;
        uint32_t* dp = reinterpret_cast<uint32_t*>(d + (size_t)y * row_bytes + x0b + b0);
        dp[0] = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        dp[1] = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
        dp[2] = o[8] | (o[9] << 8) | (o[10] << 16) | (o[11] << 24);
      }
    }
    __syncthreads();  // everyone is done with this row's slot: thread 0 may refill it next iteration
  }
  // output rows whose window reaches below the frame are border rows: written as 0 (their input rows never came)
  for (int y = max(y_out0, r_last - fr); y < y_out1; ++y) {
    if (active && y >= 0) {
      uint32_t* dp = reinterpret_cast<uint32_t*>(d + (size_t)y * row_bytes + x0b + b0);
      dp[0] = dp[1] = dp[2] = 0u;
    }
  }
}

int launch_blur(const uint8_t* const* sp, int n, int width, int height, int ksize,
                uint8_t* const* dp, cudaStream_t st) {
  if (n < 0 || width <= 0 || height <= 0 || ksize < 1 || ksize > KMAX) return SCN_E_BADARG;
  if (n == 0) return 0;
  if (!sp || !dp) return SCN_E_BADARG;
  const int fl = (ksize + 1) / 2 - 1;  // ceil(k/2.0) - 1
  const int fr = ksize / 2;
  const int k = fl + fr + 1;
  const int in_w = TW + k - 1, in_h = TH + k - 1;
  const size_t smem = (((size_t)in_h * in_w * 3 + 15) & ~(size_t)15) + (size_t)in_h * TW * 3 * 2;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(box_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         100 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  if (((height + TH - 1) / TH) > 65535) return SCN_E_UNSUPPORTED;
  for (int i0 = 0; i0 < n; i0 += SCN_MAX_PTRS) {
    const int cnt = (n - i0 < SCN_MAX_PTRS) ? (n - i0) : SCN_MAX_PTRS;
    PtrBatch s;
    MutPtrBatch d;
    bool aligned = (width % 4) == 0 && width >= 4 && height >= 3;
    for (int i = 0; i < cnt; ++i) {
      s.p[i] = sp[i0 + i];
      d.p[i] = dp[i0 + i];
      if (((uintptr_t)s.p[i] | (uintptr_t)d.p[i]) & 3) aligned = false;
    }
    bool aligned16 = (width % 16) == 0 && height >= k;
    for (int i = 0; i < cnt; ++i)
      if (((uintptr_t)s.p[i] | (uintptr_t)d.p[i]) & 15) aligned16 = false;
    static const bool stream3 = [] {
      const char* e = getenv("SCN_BLUR3");
      return e && e[0] == 's';  // SCN_BLUR3=stream: kernel_size 3 through the streaming kernel too (measurement switch)
    }();
    if (aligned16 && k >= 2 && (ksize != 3 || stream3)) {
      const size_t smem_s = (size_t)kRowSlots * (48 + kStripBytes + 128) + (size_t)k * kStripBytes * 2;
      static bool attr2 = false;
      if (!attr2) {
        cudaError_t e = cudaFuncSetAttribute(box_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr2 = true;
      }
      const unsigned d2 = (unsigned)(k * k);
      const uint32_t magic = (uint32_t)((1ull << 32) / d2) + 1u;  // floor(v * magic / 2^32) == v / d2 for v <= 255 * d2
      dim3 grid((unsigned)((width * 3 + kStripBytes - 1) / kStripBytes), (unsigned)((height + kStripRows - 1) / kStripRows),
                (unsigned)cnt);
      LaunchScope ls("box_stream_kernel", st);
      box_stream_kernel<<<grid, BS, smem_s, st>>>(s, d, width, height, fl, fr, magic);
    } else if (ksize == 3 && aligned) {
      const int wpr = width * 3 / 4;
      dim3 grid((unsigned)((wpr + B3_THREADS - 1) / B3_THREADS), (unsigned)((height + B3_ROWS - 1) / B3_ROWS),
                (unsigned)cnt);
      LaunchScope ls("box3_kernel", st);
      box3_kernel<<<grid, B3_THREADS, 0, st>>>(s, d, width, height, wpr);
    } else {
      dim3 grid((unsigned)((width + TW - 1) / TW), (unsigned)((height + TH - 1) / TH), (unsigned)cnt);
      LaunchScope ls("box_generic_kernel", st);
      box_generic_kernel<<<grid, BT, smem, st>>>(s, d, width, height, fl, fr);
    }
    int rc = launch_status();
    if (rc) return rc;
  }
  return 0;
}

}  // namespace
}  // namespace scn

extern "C" int scn_box_blur_u8c3(const uint8_t* const* host_src_ptrs, int n, int width,
                                 int height, int kernel_size, uint8_t* const* host_dst_ptrs,
                                 void* stream) {
  return scn::launch_blur(host_src_ptrs, n, width, height, kernel_size, host_dst_ptrs,
                          (cudaStream_t)stream);
}

extern "C" int scn_box_blur_u8c3_strided(const uint8_t* src, size_t stride_bytes, int n,
                                         int width, int height, int kernel_size, uint8_t* dst,
                                         void* stream) {
  if (n < 0) return SCN_E_BADARG;
  if (n == 0) return 0;
  if (!src || !dst) return SCN_E_BADARG;
  const uint8_t* sp[SCN_MAX_PTRS];
  uint8_t* dp[SCN_MAX_PTRS];
  for (int i0 = 0; i0 < n; i0 += SCN_MAX_PTRS) {
    const int cnt = (n - i0 < SCN_MAX_PTRS) ? (n - i0) : SCN_MAX_PTRS;
    for (int i = 0; i < cnt; ++i) {
      sp[i] = src + (size_t)(i0 + i) * stride_bytes;
      dp[i] = dst + (size_t)(i0 + i) * stride_bytes;
    }
    int rc = scn::launch_blur(sp, cnt, width, height, kernel_size, dp, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}
