// nv12.cu -- NV12 decoder surface -> dense RGB24 (sm_100a).
// Replaces convertNV12toRGBA / NV12_to_RGB (reference scanner/util/image.cu:109-200,229-239).
// The reference kernel does scalar byte loads and six scalar byte stores per thread; here one
// thread converts 4 horizontally adjacent pixels: one 32-bit luma load, one 32-bit chroma load
// per contributing chroma row, and three 32-bit stores (12 bytes = 4 RGB pixels), so every
// warp-level access is a contiguous 128-byte (loads) / 384-byte (stores) span.
#include "nv12_math.cuh"
#include "scn_common.cuh"

namespace scn {
namespace {

__global__ void __launch_bounds__(256)
nv12_to_rgb_kernel(PtrBatch lumas, PtrBatch chromas, MutPtrBatch rgbs, size_t pitch, int width,
                   int height, size_t rgb_pitch, int quads_per_row, int vec_ok) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (q >= quads_per_row) return;
  const uint8_t* __restrict__ luma = lumas.p[blockIdx.z];
  const uint8_t* __restrict__ chroma = chromas.p[blockIdx.z];
  uint8_t* __restrict__ rgb = rgbs.p[blockIdx.z];
  const int x0 = q * 4;
  uint32_t yy[4], cb[2], cr[2];
  const int yc = y >> 1;
  const bool avg = (y & 1) && yc < ((height >> 1) - 1);
  if (vec_ok) {
    const uint32_t yw = ld_stream_u32(luma + (size_t)y * pitch + x0);
    uint32_t cw = __ldg(reinterpret_cast<const uint32_t*>(chroma + (size_t)yc * pitch + x0));
    yy[0] = yw & 0xFF; yy[1] = (yw >> 8) & 0xFF; yy[2] = (yw >> 16) & 0xFF; yy[3] = yw >> 24;
    cb[0] = cw & 0xFF; cr[0] = (cw >> 8) & 0xFF; cb[1] = (cw >> 16) & 0xFF; cr[1] = cw >> 24;
    if (avg) {
      const uint32_t c2 =
          __ldg(reinterpret_cast<const uint32_t*>(chroma + (size_t)(yc + 1) * pitch + x0));
      cb[0] = (cb[0] + (c2 & 0xFF) + 1) >> 1;
      cr[0] = (cr[0] + ((c2 >> 8) & 0xFF) + 1) >> 1;
      cb[1] = (cb[1] + ((c2 >> 16) & 0xFF) + 1) >> 1;
      cr[1] = (cr[1] + (c2 >> 24) + 1) >> 1;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) yy[i] = (x0 + i < width) ? luma[(size_t)y * pitch + x0 + i] : 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      cb[i] = cr[i] = 0;
      if (x0 + 2 * i < width) chroma_at(chroma, pitch, height, y, x0 + 2 * i, cb[i], cr[i]);
    }
  }
  uint8_t px[12];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const Rgb8 c = yuv_to_rgb(yy[i], cb[i >> 1], cr[i >> 1]);
    px[3 * i + 0] = (uint8_t)c.r;
    px[3 * i + 1] = (uint8_t)c.g;
    px[3 * i + 2] = (uint8_t)c.b;
  }
  uint8_t* o = rgb + (size_t)y * rgb_pitch + (size_t)x0 * 3;
  if (vec_ok && x0 + 4 <= width) {
    uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
    o32[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
    o32[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
    o32[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
  } else {
    for (int i = 0; i < 12; ++i)
      if (x0 + i / 3 < width) o[i] = px[i];
  }
}

// Pitched decoder surface -> packed NV12 element (W x H luma rows, then W x H/2 CbCr rows).
// Row r of the packed image comes from the luma plane for r < H, from the chroma plane after.
// 16 bytes per thread when everything is 16-byte aligned, else bytes.
__global__ void __launch_bounds__(256)
nv12_pack_kernel(PtrBatch lumas, PtrBatch chromas, MutPtrBatch dsts, size_t pitch, int width, int height,
                 int vec16) {
  const uint8_t* __restrict__ luma = lumas.p[blockIdx.z];
  const uint8_t* __restrict__ chroma = chromas.p[blockIdx.z];
  uint8_t* __restrict__ dst = dsts.p[blockIdx.z];
  const int rows = height + (height >> 1);
  if (vec16) {
    const int vpr = width >> 4;  // 16-byte vectors per row
    const long total = (long)vpr * rows;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      const int r = (int)(i / vpr), v = (int)(i - (long)r * vpr);
      const uint8_t* src = (r < height ? luma + (size_t)r * pitch : chroma + (size_t)(r - height) * pitch) + v * 16;
      *reinterpret_cast<uint4*>(dst + (size_t)r * width + v * 16) = ld_stream_u4(src);
    }
  } else {
    const long total = (long)width * rows;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      const int r = (int)(i / width), x = (int)(i - (long)r * width);
      dst[i] = r < height ? luma[(size_t)r * pitch + x] : chroma[(size_t)(r - height) * pitch + x];
    }
  }
}

}  // namespace
}  // namespace scn

extern "C" int scn_nv12_pack(const uint8_t* const* host_luma_ptrs, const uint8_t* const* host_chroma_ptrs,
                             size_t pitch, int n, int width, int height, uint8_t* const* host_dst_ptrs,
                             void* stream) {
  using namespace scn;
  if (n < 0 || width <= 0 || height <= 0 || (width & 1) || (height & 1) || pitch < (size_t)width)
    return SCN_E_BADARG;
  if (n == 0) return 0;
  if (!host_luma_ptrs || !host_chroma_ptrs || !host_dst_ptrs) return SCN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  for (int i0 = 0; i0 < n; i0 += SCN_MAX_PTRS) {
    const int cnt = (n - i0 < SCN_MAX_PTRS) ? (n - i0) : SCN_MAX_PTRS;
    PtrBatch l, c;
    MutPtrBatch d;
    int vec16 = ((pitch & 15) == 0) && ((width & 15) == 0);
    for (int i = 0; i < cnt; ++i) {
      l.p[i] = host_luma_ptrs[i0 + i];
      c.p[i] = host_chroma_ptrs[i0 + i];
      d.p[i] = host_dst_ptrs[i0 + i];
      if (((uintptr_t)l.p[i] | (uintptr_t)c.p[i] | (uintptr_t)d.p[i]) & 15) vec16 = 0;
    }
    const long work = vec16 ? (long)(width >> 4) * (height + height / 2) : (long)width * (height + height / 2);
    long gx = (work + 255) / 256;
    const long cap = (long)sm_count() * 8 / cnt + 1;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, 1, (unsigned)cnt);
    {
      LaunchScope ls("nv12_pack_kernel", st);
      nv12_pack_kernel<<<grid, 256, 0, st>>>(l, c, d, pitch, width, height, vec16);
    }
    int rc = launch_status();
    if (rc) return rc;
  }
  return 0;
}

extern "C" int scn_nv12_to_rgb24(const uint8_t* const* host_luma_ptrs,
                                 const uint8_t* const* host_chroma_ptrs, size_t pitch, int n,
                                 int width, int height, uint8_t* const* host_rgb_ptrs,
                                 size_t rgb_pitch, void* stream) {
  using namespace scn;
  if (n < 0 || width <= 0 || height <= 0 || (width & 1) || (height & 1) ||
      pitch < (size_t)width || rgb_pitch < (size_t)width * 3)
    return SCN_E_BADARG;
  if (n == 0) return 0;
  if (!host_luma_ptrs || !host_chroma_ptrs || !host_rgb_ptrs) return SCN_E_BADARG;
  if (height > 65535) return SCN_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int quads = (width + 3) / 4;
  for (int i0 = 0; i0 < n; i0 += SCN_MAX_PTRS) {
    const int cnt = (n - i0 < SCN_MAX_PTRS) ? (n - i0) : SCN_MAX_PTRS;
    PtrBatch l, c;
    MutPtrBatch r;
    int vec_ok = ((pitch & 3) == 0) && ((rgb_pitch & 3) == 0);
    for (int i = 0; i < cnt; ++i) {
      l.p[i] = host_luma_ptrs[i0 + i];
      c.p[i] = host_chroma_ptrs[i0 + i];
      r.p[i] = host_rgb_ptrs[i0 + i];
      if (((uintptr_t)l.p[i] | (uintptr_t)c.p[i] | (uintptr_t)r.p[i]) & 3) vec_ok = 0;
    }
    dim3 grid((unsigned)((quads + 255) / 256), (unsigned)height, (unsigned)cnt);
    {
      LaunchScope ls("nv12_to_rgb_kernel", st);
      nv12_to_rgb_kernel<<<grid, 256, 0, st>>>(l, c, r, pitch, width, height, rgb_pitch, quads,
                                             vec_ok);
    }
    int rc = launch_status();
    if (rc) return rc;
  }
  return 0;
}
