// nv12_math.cuh -- the reference's NV12 -> RGB arithmetic (scanner/util/image.cu:67-102,
// :109-200), written once for every kernel that consumes decoder surfaces.
//   Y' = Y<<2, C' = (C<<2)-512 (10-bit widen), float matrix {1.1644,0,1.596; 1.1644,-.3918,
//   -.813; 1.1644,2.0172,0}, clamp [0,1023], truncate, >>2.  On odd luma rows (except the last
//   chroma row) chroma is the rounded average of the two neighbouring chroma rows (:133-151).
// The products are contracted the way nvcc's default -fmad contracts the reference source:
// fma(cr,k2, fma(cb,k1, y*k0)); terms with a zero coefficient are dropped (x + (+-0) == x).
#pragma once
#include <stdint.h>

namespace scn {

struct Rgb8 {
  uint32_t r, g, b;
};

__device__ __forceinline__ uint32_t pack10(float v) {
  v = fminf(fmaxf(v, 0.0f), 1023.f);
  return ((uint32_t)v) >> 2;
}

__device__ __forceinline__ Rgb8 yuv_to_rgb(uint32_t y, uint32_t cb, uint32_t cr) {
  const float l = (float)(y << 2);
  const float fcb = (float)((int)(cb << 2) - 512);
  const float fcr = (float)((int)(cr << 2) - 512);
  const float ly = __fmul_rn(l, 1.1644f);
  Rgb8 o;
  o.r = pack10(__fmaf_rn(fcr, 1.596f, ly));
  o.g = pack10(__fmaf_rn(fcr, -0.813f, __fmaf_rn(fcb, -0.3918f, ly)));
  o.b = pack10(__fmaf_rn(fcb, 2.0172f, ly));
  return o;
}

// chroma sample pair for luma row y at (even) column xc
__device__ __forceinline__ void chroma_at(const uint8_t* __restrict__ chroma, size_t pitch,
                                          int height, int y, int xc, uint32_t& cb,
                                          uint32_t& cr) {
  const int yc = y >> 1;
  const uint8_t* c0 = chroma + (size_t)yc * pitch + xc;
  cb = c0[0];
  cr = c0[1];
  if ((y & 1) && yc < ((height >> 1) - 1)) {
    cb = (cb + c0[pitch] + 1) >> 1;
    cr = (cr + c0[pitch + 1] + 1) >> 1;
  }
}

__device__ __forceinline__ Rgb8 nv12_pixel(const uint8_t* __restrict__ luma,
                                           const uint8_t* __restrict__ chroma, size_t pitch,
                                           int height, int x, int y) {
  uint32_t cb, cr;
  chroma_at(chroma, pitch, height, y, x & ~1, cb, cr);
  return yuv_to_rgb(luma[(size_t)y * pitch + x], cb, cr);
}

}  // namespace scn
