// hist_csa.cuh -- bit-sliced streaming histogram for large u8 HWC 3-channel frames (sm_100a).
//
// Why: a 16-bin histogram at HBM speed needs ~23-30 bytes/clk/SM; shared-memory atomics manage
// ~1 and per-byte register counters ~8 (integer pipe: 64 lanes/clk/SM).  Here counting is done
// 32 counters at a time in bit-sliced (vertical) form:
//   1. PRMT as an 8-entry byte LUT turns the high nibbles of a 32-bit word (4 pixels' bytes) into
//      one-hot bytes: word A has bit b of byte k set iff byte k's bin is b (b < 8), word B the
//      same for bins 8..14.  Bin 15 is encoded as 0xFF in A (PRMT's sign-replicate mode) and
//      recovered at the end from the known total:  n15 = (sum(A)+sum(B) - N) / 7.
//      6 PRMT + 2 SHF + 1 LOP3 per 4 bytes.
//   2. The one-hot words are summed with a Harley-Seal carry-save-adder tree (2 LOP3 per CSA,
//      15 CSAs per 16 words) into bit planes: plane p holds bit p of 32 independent counters.
//   3. Per warp span: bit-sliced butterfly add across the 32 lanes (SHFL + 2 LOP3 per plane and
//      round), lane l extracts counter l, bin-15 correction, 48 atomics into the frame's bins.
// Three accumulator sets track the byte->channel phase (a 32-bit word starts at byte offset
// = 0,1,2 mod 3); the static set index is rotated per lane at flush time.
// Main loop cost: ~3.3 integer ops per byte.  Bit-exact.
#pragma once
#include "scn_common.cuh"

namespace scn {
namespace csa {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kPlanes = 10;           // per-lane counters hold up to 1023
constexpr int kHigh = kPlanes - 4;    // planes 4..9
constexpr int kWarpBlock = 6144;      // bytes per warp per block: 12 rounds x 32 lanes x 16 B
constexpr int kMaxBlocks = 63;        // 63 * 16 words per set <= 1023

__device__ __forceinline__ void csa3(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t u = a ^ b;
  h = (a & b) | (u & c);
  l = u ^ c;
}

struct Acc {
  uint32_t p0, p1, p2, p3;     // planes 0..3 (ones, twos, fours, eights)
  uint32_t t0, t1, t2, t3;     // pending inputs of the tree inside a 16-word block
  uint32_t hi[kHigh];          // planes 4..9
  uint32_t pend4, pend5;       // pending weight-16 / weight-32 carries between blocks
};

__device__ __forceinline__ void acc_clear(Acc& a) {
  a.p0 = a.p1 = a.p2 = a.p3 = 0;
  a.t0 = a.t1 = a.t2 = a.t3 = 0;
#pragma unroll
  for (int i = 0; i < kHigh; ++i) a.hi[i] = 0;
  a.pend4 = a.pend5 = 0;
}

// K-th (0..15) word of a block for this accumulator; returns the weight-16 carry when K == 15.
template <int K>
__device__ __forceinline__ void push(Acc& a, uint32_t x, uint32_t& c16) {
  if constexpr ((K & 1) == 0) {
    a.t0 = x;
  } else {
    uint32_t tw;
    csa3(tw, a.p0, a.p0, a.t0, x);
    if constexpr (((K >> 1) & 1) == 0) {
      a.t1 = tw;
    } else {
      uint32_t fo;
      csa3(fo, a.p1, a.p1, a.t1, tw);
      if constexpr (((K >> 2) & 1) == 0) {
        a.t2 = fo;
      } else {
        uint32_t ei;
        csa3(ei, a.p2, a.p2, a.t2, fo);
        if constexpr (((K >> 3) & 1) == 0) {
          a.t3 = ei;
        } else {
          csa3(c16, a.p3, a.p3, a.t3, ei);
        }
      }
    }
  }
}

// add a 1-bit-per-counter word of weight 2^(4+from) into the high planes (ripple half-adders)
template <int FROM>
__device__ __forceinline__ void ripple(Acc& a, uint32_t c) {
#pragma unroll
  for (int q = FROM; q < kHigh; ++q) {
    const uint32_t t = a.hi[q] & c;
    a.hi[q] ^= c;
    c = t;
  }
}

// after block number `blk` (0-based) of a span: fold the weight-16 carry into the high planes
__device__ __forceinline__ void fold_block(Acc& a, uint32_t c16, int blk) {
  if (blk & 1) {
    uint32_t c32;
    csa3(c32, a.hi[0], a.hi[0], a.pend4, c16);
    if (blk & 2) {
      uint32_t c64;
      csa3(c64, a.hi[1], a.hi[1], a.pend5, c32);
      ripple<2>(a, c64);
    } else {
      a.pend5 = c32;
    }
  } else {
    a.pend4 = c16;
  }
}

// span of `nblk` blocks is over: fold the carries still pending
__device__ __forceinline__ void finish_span(Acc& a, int nblk) {
  if (nblk & 2) ripple<1>(a, a.pend5);   // weight 32 -> plane 5
  if (nblk & 1) ripple<0>(a, a.pend4);   // weight 16 -> plane 4
}

constexpr uint32_t kLutLo = 0x08040201u;   // one-hot bytes for index 0..3
constexpr uint32_t kLutHiA = 0x80402010u;  // index 4..7 (A: bin 7 -> 0x80; bin 15 reads it in sign mode -> 0xFF)
constexpr uint32_t kLutHiB = 0x00402010u;  // B: index 7 (bin 15) -> 0

// static phase set of word j of round r inside a warp block, and its position within the set
__host__ __device__ constexpr int set_of(int idx) { return (2 * (idx / 4) + (idx % 4)) % 3; }
__host__ __device__ constexpr int rank_of(int idx) {
  int c = 0;
  for (int i = 0; i < idx; ++i)
    if (set_of(i) == set_of(idx)) ++c;
  return c;
}

template <int IDX>
__device__ __forceinline__ void eat(Acc (&A)[3], Acc (&B)[3], uint32_t (&cA)[3], uint32_t (&cB)[3], uint32_t w) {
  constexpr int S = set_of(IDX);
  constexpr int K = rank_of(IDX);
  const uint32_t a01 = prmt(kLutLo, kLutHiA, w);
  const uint32_t a23 = prmt(kLutLo, kLutHiA, w >> 16);
  const uint32_t a = prmt(a01, a23, 0x7531u);
  const uint32_t wx = w ^ 0x80808080u;
  const uint32_t b01 = prmt(kLutLo, kLutHiB, wx);
  const uint32_t b23 = prmt(kLutLo, kLutHiB, wx >> 16);
  const uint32_t b = prmt(b01, b23, 0x7531u);
  push<K>(A[S], a, cA[S]);
  push<K>(B[S], b, cB[S]);
}

template <int R>
__device__ __forceinline__ void eat_round(Acc (&A)[3], Acc (&B)[3], uint32_t (&cA)[3], uint32_t (&cB)[3],
                                          const uint4 v) {
  eat<R * 4 + 0>(A, B, cA, cB, v.x);
  eat<R * 4 + 1>(A, B, cA, cB, v.y);
  eat<R * 4 + 2>(A, B, cA, cB, v.z);
  eat<R * 4 + 3>(A, B, cA, cB, v.w);
}

__device__ __forceinline__ uint32_t sel3(int rot, uint32_t x0, uint32_t x1, uint32_t x2) {
  return rot == 0 ? x0 : (rot == 1 ? x1 : x2);
}

// cross-lane sum of a bit-sliced counter set: after the call every lane holds the warp total in
// pl[0 .. kPlanes+4]
__device__ __forceinline__ void warp_sum(uint32_t (&pl)[kPlanes + 5]) {
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    uint32_t carry = 0;
#pragma unroll
    for (int p = 0; p < kPlanes + 5; ++p) {
      if (p < kPlanes + d) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, pl[p], 1 << d);
        const uint32_t u = pl[p] ^ o;
        const uint32_t nc = (pl[p] & o) | (u & carry);
        pl[p] = u ^ carry;
        carry = nc;
      } else if (p == kPlanes + d) {
        pl[p] = carry;
      }
    }
  }
}

__device__ __forceinline__ uint32_t extract_lane(const uint32_t (&pl)[kPlanes + 5], int lane) {
  uint32_t v = 0;
#pragma unroll
  for (int p = 0; p < kPlanes + 5; ++p) v |= ((pl[p] >> lane) & 1u) << p;
  return v;
}

__device__ __forceinline__ void planes_of(const Acc& a, uint32_t (&pl)[kPlanes + 5]) {
  pl[0] = a.p0;
  pl[1] = a.p1;
  pl[2] = a.p2;
  pl[3] = a.p3;
#pragma unroll
  for (int q = 0; q < kHigh; ++q) pl[4 + q] = a.hi[q];
#pragma unroll
  for (int q = kPlanes; q < kPlanes + 5; ++q) pl[q] = 0;
}

// Flush one span: add this warp's counts of `nblk` blocks into hist48 (shared or global ints).
__device__ __forceinline__ void flush_span(Acc (&A)[3], Acc (&B)[3], int nblk, int lane, int* hist48) {
  const int rot = lane % 3;  // static set s holds true phase (s + rot) % 3 on this lane
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    finish_span(A[s], nblk);
    finish_span(B[s], nblk);
  }
  const uint32_t words_per_slot = (uint32_t)nblk * 16u * 32u;  // per phase set, whole warp
#pragma unroll
  for (int phase = 0; phase < 3; ++phase) {
    // set holding true phase `phase` on this lane: (phase - rot) mod 3
    uint32_t pa[kPlanes + 5], pb[kPlanes + 5], x0[kPlanes + 5], x1[kPlanes + 5], x2[kPlanes + 5];
    const int pick = (phase + 3 - rot) % 3;
    planes_of(A[0], x0);
    planes_of(A[1], x1);
    planes_of(A[2], x2);
#pragma unroll
    for (int p = 0; p < kPlanes + 5; ++p) pa[p] = sel3(pick, x0[p], x1[p], x2[p]);
    planes_of(B[0], x0);
    planes_of(B[1], x1);
    planes_of(B[2], x2);
#pragma unroll
    for (int p = 0; p < kPlanes + 5; ++p) pb[p] = sel3(pick, x0[p], x1[p], x2[p]);
    warp_sum(pa);
    warp_sum(pb);
    uint32_t ca = extract_lane(pa, lane);  // A counter of (byte slot lane/8, bit lane%8), + n15
    uint32_t cb = extract_lane(pb, lane);  // B counter: bins 8..14 (bit 7 stays 0)
    // n15 of this byte slot: (sum A + sum B - N) / 7 over the 8 lanes of the slot
    uint32_t sum = ca + cb;
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    const uint32_t n15 = (sum - words_per_slot) / 7u;
    ca -= n15;
    const int bit = lane & 7, slot = lane >> 3;
    if (bit == 7) cb = n15;
    const int ch = (phase + slot) % 3;
    if (ca) atomicAdd(&hist48[ch * 16 + bit], (int)ca);
    if (cb) atomicAdd(&hist48[ch * 16 + 8 + bit], (int)cb);
  }
}

struct Params {
  PtrBatch frames;
  int n;                    // frames in this launch
  uint32_t blocks_per_frame;  // whole warp blocks per frame
  uint32_t tail_bytes;        // bytes after the last whole block of a frame
  uint64_t total_blocks;      // n * blocks_per_frame
};

// One warp owns a contiguous range of the global warp-block index space and walks it frame by
// frame; at most kMaxBlocks blocks between flushes.
static __global__ void __launch_bounds__(kThreads, 1)
hist16_csa_kernel(const Params prm, int32_t* __restrict__ out) {
  __shared__ int sh[kWarps][48];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* h = sh[warp];
  const uint64_t gwarp = (uint64_t)blockIdx.x * kWarps + warp;
  const uint64_t nwarps = (uint64_t)gridDim.x * kWarps;
  uint64_t g0 = prm.total_blocks * gwarp / nwarps;
  const uint64_t g1 = prm.total_blocks * (gwarp + 1) / nwarps;

  while (g0 < g1) {
    const uint32_t frame = (uint32_t)(g0 / prm.blocks_per_frame);
    const uint32_t b0 = (uint32_t)(g0 - (uint64_t)frame * prm.blocks_per_frame);
    uint32_t nb = prm.blocks_per_frame - b0;
    if ((uint64_t)nb > g1 - g0) nb = (uint32_t)(g1 - g0);
    if (nb > (uint32_t)kMaxBlocks) nb = kMaxBlocks;
    const uint8_t* __restrict__ src = prm.frames.p[frame] + (size_t)b0 * kWarpBlock + (size_t)lane * 16;

    for (int i = lane; i < 48; i += 32) h[i] = 0;
    __syncwarp();

    Acc A[3], B[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      acc_clear(A[s]);
      acc_clear(B[s]);
    }
    uint4 cur[12];
#pragma unroll
    for (int r = 0; r < 12; ++r) cur[r] = ld_stream_u4(src + r * 512);
    for (uint32_t blk = 0; blk < nb; ++blk) {
      // unconditional prefetch (the last iteration re-reads its own block): loads inside a
      // conditional region make ptxas wait for them at the reconvergence point
      uint4 nxt[12];
      const uint8_t* nsrc = src + (size_t)(blk + 1 < nb ? blk + 1 : blk) * kWarpBlock;
#pragma unroll
      for (int r = 0; r < 12; ++r) nxt[r] = ld_stream_u4(nsrc + r * 512);
      uint32_t cA[3], cB[3];
      eat_round<0>(A, B, cA, cB, cur[0]);
      eat_round<1>(A, B, cA, cB, cur[1]);
      eat_round<2>(A, B, cA, cB, cur[2]);
      eat_round<3>(A, B, cA, cB, cur[3]);
      eat_round<4>(A, B, cA, cB, cur[4]);
      eat_round<5>(A, B, cA, cB, cur[5]);
      eat_round<6>(A, B, cA, cB, cur[6]);
      eat_round<7>(A, B, cA, cB, cur[7]);
      eat_round<8>(A, B, cA, cB, cur[8]);
      eat_round<9>(A, B, cA, cB, cur[9]);
      eat_round<10>(A, B, cA, cB, cur[10]);
      eat_round<11>(A, B, cA, cB, cur[11]);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        fold_block(A[s], cA[s], (int)blk);
        fold_block(B[s], cB[s], (int)blk);
      }
#pragma unroll
      for (int r = 0; r < 12; ++r) cur[r] = nxt[r];
    }
    flush_span(A, B, (int)nb, lane, h);

    // the warp that finishes a frame's last block also counts the frame's tail bytes
    if (b0 + nb == prm.blocks_per_frame && prm.tail_bytes) {
      const uint8_t* t = prm.frames.p[frame] + (size_t)prm.blocks_per_frame * kWarpBlock;
      for (uint32_t o = lane; o < prm.tail_bytes; o += 32) atomicAdd(&h[(o % 3) * 16 + (t[o] >> 4)], 1);
    }
    __syncwarp();
    for (int i = lane; i < 48; i += 32)
      if (h[i]) atomicAdd(&out[(size_t)frame * 48 + i], h[i]);
    __syncwarp();
    g0 += nb;
  }
}

inline bool eligible(const uint8_t* const* ptrs, int n, size_t nbytes) {
  if (nbytes < (size_t)kWarpBlock * 8) return false;
  if (nbytes / kWarpBlock > 0xFFFFFFFFull) return false;
  for (int i = 0; i < n; ++i)
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) return false;
  return true;
}

// `out` must already be zeroed (launch_hist memsets it).
inline int launch(const uint8_t* const* ptrs, int n, size_t nbytes, int32_t* out, cudaStream_t st) {
  for (int i0 = 0; i0 < n; i0 += SCN_MAX_PTRS) {
    const int cnt = (n - i0 < SCN_MAX_PTRS) ? (n - i0) : SCN_MAX_PTRS;
    Params p;
    for (int i = 0; i < cnt; ++i) p.frames.p[i] = ptrs[i0 + i];
    p.n = cnt;
    p.blocks_per_frame = (uint32_t)(nbytes / kWarpBlock);
    p.tail_bytes = (uint32_t)(nbytes - (size_t)p.blocks_per_frame * kWarpBlock);
    p.total_blocks = (uint64_t)cnt * p.blocks_per_frame;
    // persistent: one CTA per SM; with little work use fewer CTAs so every warp gets >= 4 blocks
    uint64_t ctas = (p.total_blocks + kWarps * 4 - 1) / (kWarps * 4);
    if (ctas > (uint64_t)sm_count()) ctas = sm_count();
    if (ctas < 1) ctas = 1;
    {
      LaunchScope ls("hist16_csa_kernel", st);
      hist16_csa_kernel<<<(unsigned)ctas, kThreads, 0, st>>>(p, out + (size_t)i0 * 48);
    }
    int rc = launch_status();
    if (rc) return rc;
  }
  return 0;
}

}  // namespace csa
}  // namespace scn
