// nv12_csa.cuh -- histogram of the RGB image of an NV12 decoder surface without materialising it
// (the Histogram half of BASELINE configs[1], fused with the reference's NV12->RGB arithmetic,
// scanner/util/image.cu:67-200 + tests/test_ops.cpp:19-49).
//
// A thread converts a "unit" = 16 pixels x 2 rows (one chroma row pair) per step from four 128-bit
// loads (2 luma, 2 chroma).  The colour math runs on the FMA pipe in a 2^-11 scaled domain, which
// is exact (power-of-two scaling commutes with IEEE rounding) and makes the lower clamp free:
//     byte -> float      PRMT builds the bits of 2^23 + byte, one FFMA/FADD removes the bias
//     ly  = Y * (4*1.1644*2^-11)                      (== fl(Y'*1.1644) * 2^-11)
//     r   = fma.sat(cr, 4*1.596*2^-11, ly)            (.sat == max(.,0); min(.,1) never binds)
//     g   = fma.sat(cr, 4*-.813*2^-11, fma(cb, 4*-.3918*2^-11, ly));  b likewise
//     bin = floor(min(v,1023)/64) = floor(32 * min(x, 1023*2^-11))   via fma.rm(x, 32, 2^23)
// Four bins of one channel are packed into the low half of a 32-bit word (nibbles) with IMADs,
// decoded to one-hot bytes with PRMT as an 8-entry LUT and counted with the bit-sliced carry-save adders of
// hist_csa.cuh (6 accumulators: R,G,B x bins 0-7 / 8-14; bin 15 recovered from the total).
// ~11 integer-pipe + ~12 FMA-pipe operations per pixel.  Bit-exact with the three-pass path.
#pragma once
#include "csa_core.cuh"
#include "nv12_math.cuh"
#include "scn_common.cuh"

namespace scn {
namespace nvcsa {

#ifndef NVCSA_THREADS
#define NVCSA_THREADS 256
#endif
constexpr int kThreads = NVCSA_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kPlanes = 10;
constexpr int kHi = kPlanes - 3;   // planes 3..9
constexpr int kMaxSteps = 127;     // 127 * 8 words per accumulator <= 1023

using Acc = csa::Acc8<kHi>;
using csa::acc_clear;
using csa::finish_span;
using csa::fold_step;
using csa::push8;

// 2^-11 scaled constants (exact power-of-two rescalings of image.cu's matrix, times 4 for the
// 8->10 bit widening of the inputs)
constexpr float kS = 1.0f / 2048.0f;
constexpr float kCY = 4.0f * 1.1644f * kS;
constexpr float kKR = 4.0f * 1.596f * kS;
constexpr float kKG1 = 4.0f * -0.3918f * kS;
constexpr float kKG2 = 4.0f * -0.813f * kS;
constexpr float kKB = 4.0f * 2.0172f * kS;
constexpr float kMagic = 8388608.0f;        // 2^23

__device__ __forceinline__ float byte_magic(uint32_t word, uint32_t sel) {
  // bits of (2^23 + byte): byte -> mantissa LSBs, 0x4B exponent byte on top
  return __uint_as_float(prmt(word, 0x4B000000u, sel));
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float r;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2: one issue slot, two FMAs) -------------
using f2 = unsigned long long;  // two floats in an aligned register pair
__device__ __forceinline__ f2 pack2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2 fma2_rm(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 splat(float v) { return pack2(v, v); }

// Two pixels at a time (pixel j of luma word A with pixel j of luma word B), all on the FMA pipe:
//   x  = fma.sat(c, k, ly)                    value in the 2^-11 domain, lower clamp from .sat
//   y  = fma.rm(x, 32, 2^23)      [x2]        2^23 + m,  m = floor(v / 64) in 0..31
//   q  = fma.sat(y, 1/16, 1/16 - 2^19)        (m + 1) / 16 clamped to 1: (min(m, 15) + 1) / 16, exact
//   S  = fma(q, 16^(j+1), S)      [x2]        S starts at 2^23 - 0x1111, so after pixels j = 0..3 the
//                                             float's low 16 bits ARE the four bin nibbles
// -- the upper clamp costs no integer-pipe FMNMX, the packing no IMAD chain, and PRMT only reads
// the low half of the selector, so the exponent bits of S need no masking.
constexpr float kQScale = 0.0625f;
constexpr float kQBias = 0.0625f - 524288.0f;     // 1/16 - 2^19, exactly representable
constexpr float kSInit = kMagic - 4369.0f;        // 2^23 - 0x1111

struct S3 {
  f2 r, g, b;  // (word A, word B) accumulators per channel
};

struct Chroma2 {   // per (Cb,Cr) sample pair of words A and B: Cr, and the two products image.cu rounds on their own
  f2 tg, tb;       // fl(cb * kKG1), fl(cb * kKB)   (A, B)
  float crA, crB;
};
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

template <int J>  // J = pixel index inside the 4-pixel words
__device__ __forceinline__ void pixel_pair(S3& s, float yA, float yB, const Chroma2& c) {
  constexpr float w = (float)(16 << (4 * J));  // 16^(J+1)
  // image.cu's  y*k0 + cb*k1 + cr*k2  as nvcc contracts it:  fma(cr,k2, fma(y,k0, fl(cb*k1)))  (nv12_math.cuh)
  const f2 ym = pack2(yA, yB);
  const f2 yf = add2(ym, splat(-kMagic));                     // exact
  const f2 ly = fma2(ym, splat(kCY), splat(-kMagic * kCY));   // == fl(yf * kCY)  (R: the middle product is +-0)
  const f2 gi = fma2(yf, splat(kCY), c.tg);
  const f2 bi = fma2(yf, splat(kCY), c.tb);                   // B before the clamp
  float lyA, lyB, giA, giB, biA, biB;
  unpack2(ly, lyA, lyB);
  unpack2(gi, giA, giB);
  unpack2(bi, biA, biB);
  const f2 r = pack2(fma_sat(c.crA, kKR, lyA), fma_sat(c.crB, kKR, lyB));
  const f2 g = pack2(fma_sat(c.crA, kKG2, giA), fma_sat(c.crB, kKG2, giB));
  const f2 b = pack2(fma_sat(biA, 1.0f, 0.0f), fma_sat(biB, 1.0f, 0.0f));
  const f2 k32 = splat(32.0f), mg = splat(kMagic);
  float ra, rb, ga, gb, ba, bb;
  unpack2(fma2_rm(r, k32, mg), ra, rb);
  unpack2(fma2_rm(g, k32, mg), ga, gb);
  unpack2(fma2_rm(b, k32, mg), ba, bb);
  const f2 qr = pack2(fma_sat(ra, kQScale, kQBias), fma_sat(rb, kQScale, kQBias));
  const f2 qg = pack2(fma_sat(ga, kQScale, kQBias), fma_sat(gb, kQScale, kQBias));
  const f2 qb = pack2(fma_sat(ba, kQScale, kQBias), fma_sat(bb, kQScale, kQBias));
  s.r = fma2(qr, splat(w), s.r);
  s.g = fma2(qg, splat(w), s.g);
  s.b = fma2(qb, splat(w), s.b);
}

// `lut_lo` is csa::kLutLo held in a vector register: PRMT takes only one immediate, and as a literal
// (or any value ptxas can prove uniform) the low LUT word sits in a uniform register and is copied
// with an IMAD before every PRMT -- 36 issue slots per step
template <int K>
__device__ __forceinline__ void count4(Acc& A, Acc& B, uint32_t z, uint32_t& cA, uint32_t& cB, uint32_t lut_lo) {
  push8<K>(A, prmt(lut_lo, csa::kLutHiA, z), cA);
  push8<K>(B, prmt(lut_lo, csa::kLutHiB, z ^ 0x8888u), cB);
}

// 8 pixels: luma words yA, yB (4 px each) with their chroma words cA, cB (2 Cb,Cr pairs each);
// feeds pushes K and K+1 of every accumulator
template <int K>
__device__ __forceinline__ void eight(Acc (&A)[3], Acc (&B)[3], uint32_t (&cyA)[3], uint32_t (&cyB)[3], uint32_t yA,
                                      uint32_t yB, uint32_t cA, uint32_t cB, uint32_t lut_lo) {
  const f2 bias = splat(-(kMagic + 128.0f));
  float cb0A, cb0B, cr0A, cr0B, cb1A, cb1B, cr1A, cr1B;
  unpack2(add2(pack2(byte_magic(cA, 0x7440u), byte_magic(cB, 0x7440u)), bias), cb0A, cb0B);
  unpack2(add2(pack2(byte_magic(cA, 0x7441u), byte_magic(cB, 0x7441u)), bias), cr0A, cr0B);
  unpack2(add2(pack2(byte_magic(cA, 0x7442u), byte_magic(cB, 0x7442u)), bias), cb1A, cb1B);
  unpack2(add2(pack2(byte_magic(cA, 0x7443u), byte_magic(cB, 0x7443u)), bias), cr1A, cr1B);
  S3 s{splat(kSInit), splat(kSInit), splat(kSInit)};
  const f2 cb0 = pack2(cb0A, cb0B), cb1 = pack2(cb1A, cb1B);
  const Chroma2 c0{mul2(cb0, splat(kKG1)), mul2(cb0, splat(kKB)), cr0A, cr0B};
  const Chroma2 c1{mul2(cb1, splat(kKG1)), mul2(cb1, splat(kKB)), cr1A, cr1B};
  pixel_pair<0>(s, byte_magic(yA, 0x7440u), byte_magic(yB, 0x7440u), c0);
  pixel_pair<1>(s, byte_magic(yA, 0x7441u), byte_magic(yB, 0x7441u), c0);
  pixel_pair<2>(s, byte_magic(yA, 0x7442u), byte_magic(yB, 0x7442u), c1);
  pixel_pair<3>(s, byte_magic(yA, 0x7443u), byte_magic(yB, 0x7443u), c1);
  float lo, hi;
  unpack2(s.r, lo, hi);
  count4<K>(A[0], B[0], __float_as_uint(lo), cyA[0], cyB[0], lut_lo);
  count4<K + 1>(A[0], B[0], __float_as_uint(hi), cyA[0], cyB[0], lut_lo);
  unpack2(s.g, lo, hi);
  count4<K>(A[1], B[1], __float_as_uint(lo), cyA[1], cyB[1], lut_lo);
  count4<K + 1>(A[1], B[1], __float_as_uint(hi), cyA[1], cyB[1], lut_lo);
  unpack2(s.b, lo, hi);
  count4<K>(A[2], B[2], __float_as_uint(lo), cyA[2], cyB[2], lut_lo);
  count4<K + 1>(A[2], B[2], __float_as_uint(hi), cyA[2], cyB[2], lut_lo);
}

__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b) {  // per byte (a + b + 1) >> 1
  return (a | b) - (((a ^ b) & 0xFEFEFEFEu) >> 1);
}

// csa::kLutLo once per lane: loaded with a per-lane index so the value is not provably uniform
__device__ const uint32_t g_lut_lo[32] = {
    csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo,
    csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo,
    csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo,
    csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo, csa::kLutLo};
__device__ uint4 g_zero_page[1];  // what inactive lanes load (zero-initialised module memory)

struct Params {
  PtrBatch luma, chroma;
  size_t pitch;
  int width, height;
  uint32_t units_per_row;     // width / 16
  uint32_t units_per_frame;   // units_per_row * height / 2
  uint32_t steps_per_frame;   // ceil(units_per_frame / 32)
  uint64_t total_steps;
};

__device__ __forceinline__ uint32_t lane_count(const Acc& a, int lane) {
  uint32_t pl[kPlanes + 5];
  csa::planes_of(a, pl);
  csa::warp_sum<kPlanes>(pl);
  uint32_t v = csa::extract_lane<kPlanes>(pl, lane);
  // the four byte slots of a word carry the same channel: fold them onto lanes 0..7
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

static __global__ void __launch_bounds__(kThreads, 1)
nv12_hist_csa_kernel(const Params prm, int32_t* __restrict__ out) {
  __shared__ int sh[kWarps][48];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* h = sh[warp];
  const uint64_t gwarp = (uint64_t)blockIdx.x * kWarps + warp;
  const uint64_t nwarps = (uint64_t)gridDim.x * kWarps;
  uint64_t g0 = prm.total_steps * gwarp / nwarps;
  const uint64_t g1 = prm.total_steps * (gwarp + 1) / nwarps;
  const int last_crow = (prm.height >> 1) - 1;
#ifdef NVCSA_LITERAL_LUT
  const uint32_t lut_lo = csa::kLutLo;
#else
  const uint32_t lut_lo = g_lut_lo[lane];
#endif

  while (g0 < g1) {
    const uint32_t frame = (uint32_t)(g0 / prm.steps_per_frame);
    const uint32_t s0 = (uint32_t)(g0 - (uint64_t)frame * prm.steps_per_frame);
    uint32_t ns = prm.steps_per_frame - s0;
    if ((uint64_t)ns > g1 - g0) ns = (uint32_t)(g1 - g0);
    if (ns > (uint32_t)kMaxSteps) ns = kMaxSteps;
    const uint8_t* __restrict__ luma = prm.luma.p[frame];
    const uint8_t* __restrict__ chroma = prm.chroma.p[frame];

    for (int i = lane; i < 48; i += 32) h[i] = 0;
    __syncwarp();
    Acc A[3], B[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      acc_clear(A[c]);
      acc_clear(B[c]);
    }
    uint32_t padded_units = 0;

    // Loads are unconditional and never sit in a divergent region: ptxas makes the first
    // instruction after a reconvergence point wait for every load issued inside it, which
    // turned the prefetch of the next step into a blocking load (r01 ncu: 30 % of all stall
    // samples on that one instruction).  Inactive lanes (only in a frame's last step) read a
    // zero page instead, and the step index is clamped instead of predicated.
    auto load_unit = [&](uint32_t step, uint4& ya, uint4& yb, uint4& c0, uint4& c1) -> bool {
      const uint32_t unit = (s0 + step) * 32u + (uint32_t)lane;
      const bool active = unit < prm.units_per_frame;
      const uint32_t yp = unit / prm.units_per_row;
      const uint32_t xs = unit - yp * prm.units_per_row;
      const uint8_t* lp = luma + (size_t)(2 * yp) * prm.pitch + (size_t)xs * 16;
      const uint8_t* cp = chroma + (size_t)yp * prm.pitch + (size_t)xs * 16;
      const size_t cnext = ((int)yp < last_crow) ? prm.pitch : 0;
      const uint8_t* zero = reinterpret_cast<const uint8_t*>(g_zero_page);
      const uint8_t* p0 = active ? lp : zero;
      const uint8_t* p1 = active ? lp + prm.pitch : zero;
      const uint8_t* p2 = active ? cp : zero;
      const uint8_t* p3 = active ? cp + cnext : zero;
      ya = ld_stream_u4(p0);
      yb = ld_stream_u4(p1);
      c0 = ld_stream_u4(p2);
      c1 = ld_stream_u4(p3);
      return active;
    };

    uint4 ya, yb, c0, c1;
    bool act = load_unit(0, ya, yb, c0, c1);
    for (uint32_t step = 0; step < ns; ++step) {
      uint4 nya, nyb, nc0, nc1;
      const uint32_t nstep = step + 1 < ns ? step + 1 : step;  // last iteration re-reads its own unit
      const bool nact = load_unit(nstep, nya, nyb, nc0, nc1);
      padded_units += act ? 0u : 1u;
      // odd luma row: rounded average of the two neighbouring chroma rows (image.cu:133-151)
      const uint4 ca = make_uint4(avg4(c0.x, c1.x), avg4(c0.y, c1.y), avg4(c0.z, c1.z), avg4(c0.w, c1.w));
      uint32_t cA[3], cB[3];
      eight<0>(A, B, cA, cB, ya.x, ya.y, c0.x, c0.y, lut_lo);
      eight<2>(A, B, cA, cB, ya.z, ya.w, c0.z, c0.w, lut_lo);
      eight<4>(A, B, cA, cB, yb.x, yb.y, ca.x, ca.y, lut_lo);
      eight<6>(A, B, cA, cB, yb.z, yb.w, ca.z, ca.w, lut_lo);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        fold_step(A[c], cA[c], (int)step);
        fold_step(B[c], cB[c], (int)step);
      }
      ya = nya;
      yb = nyb;
      c0 = nc0;
      c1 = nc1;
      act = nact;
    }

    // ---- flush the span
    const uint32_t values = ns * 32u * 32u;  // pixels (per channel) this warp fed, padding included
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      finish_span(A[c], (int)ns);
      finish_span(B[c], (int)ns);
      uint32_t ca = lane_count(A[c], lane);  // lanes 0..7: bins 0..7 (+ n15 each)
      uint32_t cb = lane_count(B[c], lane);  // lanes 0..6: bins 8..14
      uint32_t sum = (lane < 8) ? ca + cb : 0u;
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      const uint32_t n15 = (sum - values) / 7u;
      if (lane < 8) {
        ca -= n15;
        if (lane == 7) cb = n15;
        if (ca) atomicAdd(&h[c * 16 + lane], (int)ca);
        if (cb) atomicAdd(&h[c * 16 + 8 + lane], (int)cb);
      }
    }
    // padding units were fed as all-zero bytes: remove what 32 such pixels each contributed
    const uint32_t pad_total = __reduce_add_sync(0xffffffffu, padded_units);
    __syncwarp();
    if (pad_total && lane == 0) {
      const Rgb8 z = yuv_to_rgb(0, 0, 0);
      atomicSub(&h[0 * 16 + (z.r >> 4)], (int)(pad_total * 32u));
      atomicSub(&h[1 * 16 + (z.g >> 4)], (int)(pad_total * 32u));
      atomicSub(&h[2 * 16 + (z.b >> 4)], (int)(pad_total * 32u));
    }
    __syncwarp();
    for (int i = lane; i < 48; i += 32)
      if (h[i]) atomicAdd(&out[(size_t)frame * 48 + i], h[i]);
    __syncwarp();
    g0 += ns;
  }
}

inline bool eligible(const uint8_t* const* lp, const uint8_t* const* cp, int n, size_t pitch, int width, int height) {
  if ((width & 15) || (pitch & 15) || (height & 1)) return false;
  if ((size_t)width * height < 64 * 1024) return false;
  for (int i = 0; i < n; ++i)
    if ((reinterpret_cast<uintptr_t>(lp[i]) | reinterpret_cast<uintptr_t>(cp[i])) & 15) return false;
  return true;
}

// `out` must already be zeroed.
inline int launch(const uint8_t* const* lp, const uint8_t* const* cp, int n, size_t pitch, int width, int height,
                  int32_t* out, cudaStream_t st) {
  for (int i0 = 0; i0 < n; i0 += SCN_MAX_PTRS) {
    const int cnt = (n - i0 < SCN_MAX_PTRS) ? (n - i0) : SCN_MAX_PTRS;
    Params p;
    for (int i = 0; i < cnt; ++i) {
      p.luma.p[i] = lp[i0 + i];
      p.chroma.p[i] = cp[i0 + i];
    }
    p.pitch = pitch;
    p.width = width;
    p.height = height;
    p.units_per_row = (uint32_t)(width / 16);
    p.units_per_frame = p.units_per_row * (uint32_t)(height / 2);
    p.steps_per_frame = (p.units_per_frame + 31) / 32;
    p.total_steps = (uint64_t)cnt * p.steps_per_frame;
    uint64_t ctas = (p.total_steps + kWarps * 8 - 1) / (kWarps * 8);
    if (ctas > (uint64_t)sm_count()) ctas = sm_count();
    if (ctas < 1) ctas = 1;
    {
      LaunchScope ls("nv12_hist_csa_kernel", st);
      nv12_hist_csa_kernel<<<(unsigned)ctas, kThreads, 0, st>>>(p, out + (size_t)i0 * 48);
    }
    int rc = launch_status();
    if (rc) return rc;
  }
  return 0;
}

}  // namespace nvcsa
}  // namespace scn
