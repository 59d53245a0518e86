// blur.cu -- integer box filter on u8 HWC 3-channel frames (sm_100a).
// Replaces BlurKernel::execute (reference tests/test_ops.cpp:265-294): for interior pixels
// out = (sum of the (fl+fr+1)^2 window) / (fl+fr+1)^2 with fl = ceil(k/2)-1, fr = k/2; the
// reference never writes the border of its freshly allocated frame -- here it is written as 0.
//
// Tiled: a CTA produces a TW x TH pixel tile; it stages the (TW+k-1) x (TH+k-1) source window in
// shared memory with coalesced row reads, runs a separable running sum (horizontal pass into
// shared u16 sums, vertical pass from those), so each source byte is read from HBM/L2 once per
// tile instead of k^2 times.
#include "scn_common.cuh"

namespace scn {
namespace {

constexpr int TW = 64;   // tile width in pixels
constexpr int TH = 32;   // tile height in pixels
constexpr int KMAX = 31;
constexpr int BT = 256;

__global__ void __launch_bounds__(BT)
box_blur_kernel(PtrBatch src, MutPtrBatch dst, int width, int height, int fl, int fr) {
  extern __shared__ uint8_t smem_raw[];
  const int k = fl + fr + 1;
  const int in_w = TW + k - 1;   // pixels
  const int in_h = TH + k - 1;
  const int in_wb = in_w * 3;    // bytes per staged row
  uint8_t* tile = smem_raw;                                          // in_h x in_wb
  uint16_t* hsum = reinterpret_cast<uint16_t*>(smem_raw + ((in_h * in_wb + 15) & ~15));  // in_h x TW*3

  const uint8_t* __restrict__ s = src.p[blockIdx.z];
  uint8_t* __restrict__ d = dst.p[blockIdx.z];
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;

  // stage source window rows [y0-fl, y0+TH+fr) x cols [x0-fl, x0+TW+fr), zero outside the frame
  for (int i = threadIdx.x; i < in_h * in_wb; i += BT) {
    const int r = i / in_wb, cbyte = i - r * in_wb;
    const int sy = y0 - fl + r;
    const int sxb = (x0 - fl) * 3 + cbyte;
    uint8_t v = 0;
    if (sy >= 0 && sy < height && sxb >= 0 && sxb < width * 3) v = __ldg(s + (size_t)sy * width * 3 + sxb);
    tile[i] = v;
  }
  __syncthreads();
  // horizontal window sums: hsum[r][x*3+c] = sum_{j<k} tile[r][(x+j)*3+c]
  for (int i = threadIdx.x; i < in_h * TW * 3; i += BT) {
    const int r = i / (TW * 3), xb = i - r * (TW * 3);
    const uint8_t* t = tile + r * in_wb + xb;
    uint32_t acc = 0;
    for (int j = 0; j < k; ++j) acc += t[j * 3];
    hsum[i] = (uint16_t)acc;   // <= 31*255 = 7905
  }
  __syncthreads();
  const uint32_t div = (uint32_t)(k * k);
  for (int i = threadIdx.x; i < TH * TW * 3; i += BT) {
    const int r = i / (TW * 3), xb = i - r * (TW * 3);
    const int x = x0 + xb / 3, y = y0 + r;
    if (x >= width || y >= height) continue;
    uint32_t acc = 0;
    for (int j = 0; j < k; ++j) acc += hsum[(r + j) * (TW * 3) + xb];
    const bool interior = (y >= fl) && (y < height - fr) && (x >= fl) && (x < width - fr);
    d[((size_t)y * width) * 3 + (size_t)x0 * 3 + xb] = interior ? (uint8_t)(acc / div) : (uint8_t)0;
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
    cudaError_t e = cudaFuncSetAttribute(box_blur_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  if (((height + TH - 1) / TH) > 65535) return SCN_E_UNSUPPORTED;
  for (int i0 = 0; i0 < n; i0 += SCN_MAX_PTRS) {
    const int cnt = (n - i0 < SCN_MAX_PTRS) ? (n - i0) : SCN_MAX_PTRS;
    PtrBatch s;
    MutPtrBatch d;
    for (int i = 0; i < cnt; ++i) {
      s.p[i] = sp[i0 + i];
      d.p[i] = dp[i0 + i];
    }
    dim3 grid((unsigned)((width + TW - 1) / TW), (unsigned)((height + TH - 1) / TH), (unsigned)cnt);
    {
      LaunchScope ls("box_blur_kernel", st);
      box_blur_kernel<<<grid, BT, smem, st>>>(s, d, width, height, fl, fr);
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
