/*
 * scn_oracle.c -- CPU restatement of the reference's per-frame pixel ops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (scanner_b200/, the
 * C-ABI in include/) may link, import or call this file.  It exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the
 * CUDA path against an independent scalar implementation.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the scanner-research/scanner tree).  The pixel arithmetic of Histogram and
 * Resize lives in OpenCV (pinned 4.2.0 by the reference's deps.sh:643), which
 * is not vendored: those two functions restate OpenCV's published algorithm and
 * are pinned against cv2 outputs committed under tests/golden/ (see
 * oracle/make_golden.py).  Blur and NV12->RGB are fully open-coded in the
 * reference and are restated from its source directly.
 *
 * Build: see oracle/Makefile (plain gcc, -ffp-contract=off so that float
 * expressions round exactly as written).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Histogram -- tests/test_ops.cpp:13-59.
 * cv::calcHist(&img,1,{j},Mat(),hist,1,&BINS(16),range [0,256)) per channel j,
 * converted to CV_32S.  With 16 uniform bins over [0,256) the bin of an 8-bit
 * value v is floor(v * 16 / 256) == v >> 4.  Output layout: int32[3][16],
 * channel-major (test_ops.cpp:40-42: output_buf + j*BINS*sizeof(int)). */
ORC_API void orc_hist16_u8c3(const uint8_t* frame, int width, int height,
                             int32_t* out48) {
  memset(out48, 0, 48 * sizeof(int32_t));
  size_t npix = (size_t)width * (size_t)height;
  for (size_t p = 0; p < npix; ++p) {
    out48[0 * 16 + (frame[3 * p + 0] >> 4)]++;
    out48[1 * 16 + (frame[3 * p + 1] >> 4)]++;
    out48[2 * 16 + (frame[3 * p + 2] >> 4)]++;
  }
}

/* ------------------------------------------------------------------------- */
/* Resize target size -- tests/test_ops.cpp:126-147 (ResizeKernel::execute).
 * width/height/min/preserve_aspect are the ResizeArgs fields
 * (tests/test_ops.proto:8-14); the interpolation string is ignored by the
 * reference (test_ops.cpp:156 calls cv::resize with the default). */
ORC_API void orc_resize_target(int src_w, int src_h, int arg_w, int arg_h,
                               int arg_min, int arg_preserve_aspect,
                               int* out_w, int* out_h) {
  int tw = arg_w, th = arg_h;
  if (arg_preserve_aspect) {
    if (tw == 0) {
      tw = src_w * th / src_h;
    } else {
      th = src_h * tw / src_w;
    }
  }
  if (arg_min) {
    if (src_w <= tw && src_h <= th) {
      tw = src_w;
      th = src_h;
    }
  }
  *out_w = tw;
  *out_h = th;
}

/* ------------------------------------------------------------------------- */
/* Resize -- tests/test_ops.cpp:156: cv::resize(img, out, Size(w,h)) i.e.
 * INTER_LINEAR on CV_8UC3.  OpenCV (modules/imgproc/src/resize.cpp, 4.x):
 *   - if both scale factors are exactly 2 the call is re-routed to the
 *     INTER_AREA fast path: (a+b+c+d+2)>>2 over each 2x2 block;
 *   - otherwise: per destination column dx
 *        fx = (float)((dx+0.5)*scale_x - 0.5); sx = floor(fx); fx -= sx;
 *        sx<0 -> (sx,fx)=(0,0); sx>=sw-1 -> (sx,fx)=(sw-1,0)
 *        alpha = { rint((1.f-fx)*2048), rint(fx*2048) }   (saturate_cast<short>)
 *     per destination row dy the same WITHOUT the clamp-with-zeroed-weight:
 *        sy = floor(fy), beta from fy, and the two source rows are
 *        clip(sy,0,sh-1), clip(sy+1,0,sh-1);
 *     horizontal pass in int:  H = S[sx]*a0 + S[sx+1]*a1      (scale 2^11)
 *     vertical pass (u8 specialisation, identical in the SIMD and scalar tails):
 *        D = ( ((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2 ) >> 2
 * Pinned bit-exact against cv2.resize 4.13.0 for up- and down-scales
 * (tests/golden/resize_*.npz). */
static void orc_linear_tab(int ssz, int dsz, int clamp_zero, int* ofs,
                           short* coef /* 2 per entry */) {
  double inv_scale = (double)dsz / (double)ssz;
  double scale = 1.0 / inv_scale;
  for (int d = 0; d < dsz; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_zero) {
      if (s < 0) {
        f = 0.f;
        s = 0;
      }
      if (s >= ssz - 1) {
        f = 0.f;
        s = ssz - 1;
      }
    }
    ofs[d] = s;
    float c0 = (1.f - f) * 2048.f;
    float c1 = f * 2048.f;
    coef[2 * d + 0] = (short)lrintf(c0);
    coef[2 * d + 1] = (short)lrintf(c1);
  }
}

ORC_API void orc_resize_bilinear_u8(const uint8_t* src, int sw, int sh, int cn,
                                    uint8_t* dst, int dw, int dh) {
  if (dw <= 0 || dh <= 0) return;
  if (sw == 2 * dw && sh == 2 * dh) {
    for (int y = 0; y < dh; ++y) {
      const uint8_t* r0 = src + (size_t)(2 * y) * sw * cn;
      const uint8_t* r1 = r0 + (size_t)sw * cn;
      for (int x = 0; x < dw; ++x)
        for (int c = 0; c < cn; ++c) {
          int v = r0[(2 * x) * cn + c] + r0[(2 * x + 1) * cn + c] +
                  r1[(2 * x) * cn + c] + r1[(2 * x + 1) * cn + c];
          dst[((size_t)y * dw + x) * cn + c] = (uint8_t)((v + 2) >> 2);
        }
    }
    return;
  }
  int* xofs = (int*)malloc(sizeof(int) * dw);
  int* yofs = (int*)malloc(sizeof(int) * dh);
  short* xa = (short*)malloc(sizeof(short) * 2 * dw);
  short* yb = (short*)malloc(sizeof(short) * 2 * dh);
  int* row0 = (int*)malloc(sizeof(int) * dw * cn);
  int* row1 = (int*)malloc(sizeof(int) * dw * cn);
  orc_linear_tab(sw, dw, 1, xofs, xa);
  orc_linear_tab(sh, dh, 0, yofs, yb);
  for (int y = 0; y < dh; ++y) {
    int sy0 = yofs[y], sy1 = yofs[y] + 1;
    if (sy0 < 0) sy0 = 0;
    if (sy0 > sh - 1) sy0 = sh - 1;
    if (sy1 < 0) sy1 = 0;
    if (sy1 > sh - 1) sy1 = sh - 1;
    const uint8_t* s0 = src + (size_t)sy0 * sw * cn;
    const uint8_t* s1 = src + (size_t)sy1 * sw * cn;
    for (int x = 0; x < dw; ++x) {
      int sx0 = xofs[x];
      int sx1 = sx0 + 1 < sw ? sx0 + 1 : sw - 1;
      int a0 = xa[2 * x], a1 = xa[2 * x + 1];
      for (int c = 0; c < cn; ++c) {
        row0[x * cn + c] = s0[sx0 * cn + c] * a0 + s0[sx1 * cn + c] * a1;
        row1[x * cn + c] = s1[sx0 * cn + c] * a0 + s1[sx1 * cn + c] * a1;
      }
    }
    int b0 = yb[2 * y], b1 = yb[2 * y + 1];
    uint8_t* d = dst + (size_t)y * dw * cn;
    for (int i = 0; i < dw * cn; ++i) {
      d[i] = (uint8_t)((((b0 * (row0[i] >> 4)) >> 16) +
                        ((b1 * (row1[i] >> 4)) >> 16) + 2) >>
                       2);
    }
  }
  free(xofs);
  free(yofs);
  free(xa);
  free(yb);
  free(row0);
  free(row1);
}

/* ------------------------------------------------------------------------- */
/* Blur -- tests/test_ops.cpp:239-310.  A box filter (sigma is parsed and never
 * used): filter_left = ceil(k/2.0)-1, filter_right = k/2 (:252-253); only
 * interior pixels y in [fl, H-fr), x in [fl, W-fr) are written (:278-279) with
 * value = (sum over the (fl+fr+1)^2 window as u32) / ((fl+fr+1)^2) (integer
 * division, :288-290).  The reference leaves the border of its freshly
 * allocated output frame UNINITIALISED; this restatement (and the CUDA path)
 * defines the border as 0 -- parity is claimed on the interior only. */
ORC_API void orc_blur_u8c3(const uint8_t* src, int width, int height,
                           int kernel_size, uint8_t* dst) {
  int fl = (int)ceil(kernel_size / 2.0) - 1;
  int fr = kernel_size / 2;
  memset(dst, 0, (size_t)width * height * 3);
  uint32_t div = (uint32_t)((fr + fl + 1) * (fr + fl + 1));
  for (int y = fl; y < height - fr; ++y)
    for (int x = fl; x < width - fr; ++x)
      for (int c = 0; c < 3; ++c) {
        uint32_t value = 0;
        for (int ry = -fl; ry < fr + 1; ++ry)
          for (int rx = -fl; rx < fr + 1; ++rx)
            value += src[((size_t)(y + ry) * width + (x + rx)) * 3 + c];
        dst[((size_t)y * width + x) * 3 + c] = (uint8_t)(value / div);
      }
}

/* ------------------------------------------------------------------------- */
/* NV12 -> RGB24 -- scanner/util/image.cu:67-102 (matrix, clamp, pack) and
 * :109-200 (pixel fetch, odd-row chroma averaging).
 *   Y' = Y<<2;  C' = (C<<2) - 512 with, on odd luma rows that are not in the
 *   last chroma row, C = (C[r] + C[r+1] + 1) >> 1   (:133-151)
 *   R = Y'*1.1644 + Cb'*0      + Cr'*1.596
 *   G = Y'*1.1644 + Cb'*-.3918 + Cr'*-.813
 *   B = Y'*1.1644 + Cb'*2.0172 + Cr'*0
 *   clamp to [0,1023], truncate to uint, >>2   (:92-102)
 * The reference is CUDA source compiled by nvcc with its default -fmad=true,
 * under which `a*b + c*d + e*f` contracts to fma(e,f, fma(c,d, a*b)); that
 * contraction is written out explicitly here so the CPU and GPU agree bit for
 * bit.  (There is no -16 luma offset in the reference, :74.) */
static inline uint8_t orc_pack10(float v) {
  v = fminf(fmaxf(v, 0.0f), 1023.f);
  return (uint8_t)(((uint32_t)v) >> 2);
}

ORC_API void orc_nv12_to_rgb24(const uint8_t* luma, const uint8_t* chroma,
                               size_t pitch, int width, int height,
                               uint8_t* rgb, size_t rgb_pitch) {
  for (int y = 0; y < height; ++y) {
    int yc = y >> 1;
    for (int x = 0; x < width; ++x) {
      int xc = x & ~1;
      uint32_t cb = chroma[(size_t)yc * pitch + xc];
      uint32_t cr = chroma[(size_t)yc * pitch + xc + 1];
      if ((y & 1) && yc < ((height >> 1) - 1)) {
        cb = (cb + chroma[(size_t)(yc + 1) * pitch + xc] + 1) >> 1;
        cr = (cr + chroma[(size_t)(yc + 1) * pitch + xc + 1] + 1) >> 1;
      }
      float l = (float)((uint32_t)luma[(size_t)y * pitch + x] << 2);
      float fcb = (float)((int)(cb << 2) - 512);
      float fcr = (float)((int)(cr << 2) - 512);
      float r = fmaf(fcr, 1.596f, fmaf(fcb, 0.0f, l * 1.1644f));
      float g = fmaf(fcr, -0.813f, fmaf(fcb, -0.3918f, l * 1.1644f));
      float b = fmaf(fcr, 0.0f, fmaf(fcb, 2.0172f, l * 1.1644f));
      uint8_t* o = rgb + (size_t)y * rgb_pitch + (size_t)x * 3;
      o[0] = orc_pack10(r);
      o[1] = orc_pack10(g);
      o[2] = orc_pack10(b);
    }
  }
}

/* ------------------------------------------------------------------------- */
/* The C2 DAG of BASELINE.json configs[1] on one decoded surface:
 * NV12 -> RGB24 -> { Histogram, Resize(dw,dh) }.  Restated as the plain
 * composition of the three functions above (that is what the reference's
 * pipeline does: decoder output column feeds both ops,
 * scanner/engine/evaluate_worker.cpp:710-1261). */
ORC_API void orc_nv12_hist_resize(const uint8_t* luma, const uint8_t* chroma,
                                  size_t pitch, int width, int height,
                                  int32_t* out48, uint8_t* resized, int dw,
                                  int dh, uint8_t* scratch_rgb) {
  orc_nv12_to_rgb24(luma, chroma, pitch, width, height, scratch_rgb,
                    (size_t)width * 3);
  orc_hist16_u8c3(scratch_rgb, width, height, out48);
  if (resized) orc_resize_bilinear_u8(scratch_rgb, width, height, 3, resized, dw, dh);
}

/* Index column -- scanner/engine/ingest.cpp:337-345: row i is the
 * little-endian int64 i. */
ORC_API void orc_index_column(int64_t start, int64_t n, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    uint64_t v = (uint64_t)(start + i);
    for (int b = 0; b < 8; ++b) out[i * 8 + b] = (uint8_t)(v >> (8 * b));
  }
}
