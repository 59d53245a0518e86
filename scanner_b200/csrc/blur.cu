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
//   box_generic_kernel   any kernel_size <= 31: shared-memory tile, separable direct sums.
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
    if (ksize == 3 && aligned) {
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
