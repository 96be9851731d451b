// dwt_fast.cu -- streaming lifting DWT kernels for sm_100a: the forward and inverse transforms of every level whose
// geometry allows it (width a multiple of 8, no horizontal padding on the u8 side), including the level shift, the
// mirror padding in y, the quantiser / de-quantiser, the u8 narrowing and -- for RGB frames -- the colour transforms.
//
// Replaces the per-level kernels of the reference's host loops DWTEngine::DWTForward / DWTReverse
// (DWT/DWTGenerator.cu:753-1124,1268-1424: one kernel + one cudaStreamSynchronize per level, 64x18 register tiles with
// 1.4x / 2.1x read amplification) together with offsetImage (Engines/CodingEngine.cu:581-588),
// removeOffsetAndApplyMaxMin[Lossy] (Engines/DecodingEngine.cu:706-729) and RGBTransform* (CodingEngine.cu:357-405,
// DecodingEngine.cu:599-650).  Arithmetic, boundary rules and the quantiser are the ones of dwt.cu /
// oracle/picsong_oracle.c, bit for bit.
//
// Design: a warp owns a 256-sample-wide strip (8 samples per lane, 16-byte subband accesses; lanes 0 and 31 are halo)
// and streams down a chunk of row pairs.  The vertical lifting state lives in registers per column, the horizontal
// lifting of a finished row pair needs four shuffles; image borders are explicit mirror rules in the lifting steps, so
// halo lanes outside the image never influence a result.  Bandwidth-critical inputs (u8 pixels forward, subband rows
// of the finest inverse level) are staged through a per-warp cp.async ring; 9/7 arithmetic is issued as packed
// f32x2 instructions.  (Levels 1+2 of the forward transform of u8 frames normally go through dwt_fused.cu; the in-warp fusion of
// two levels and the persistent work queue this file once carried were measured, lost, and are gone: DESIGN.md 4.1.)
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "dwt_lift.cuh"

namespace pcs {

namespace {

struct FastFwdArgs {
  const void* src; long long src_fstride; int src_pitch;    // level input: u8 pixels (level 1) or the 32-bit LL plane of the level above
  int H;                                                    // u8 input: original height (rows >= H are mirror padding); else == Hc
  int Wc, Hc;                                               // level input size (multiples of 8 / 2)
  void* out; long long out_fstride; int pitch;              // Mallat plane and its row pitch (the adapted frame width)
  void* ll; long long ll_fstride; int ll_pitch;             // LL destination of the LAST level of this launch (scratch), unless `last`
  int last;                                                 // that level is the last level: LL goes to `out`, quantised
  float q1[4], qs;                                          // quantiser steps of this level: LL, HL, LH, HH
  int nW1, ls1;                                             // row-variant tiling of the reference (9/7 only)
  int nstrips, nchunks, rp;                                 // rp: row pairs per chunk (multiple of 4)
  long long comp_stride;                                    // RGB input: bytes between the R, G and B planes of a frame
  int offset;                                               // u8 input: level shift 2^(bps-1), 0 for signed samples
};


// ---- colour transform fused into the level-1 load (RGB path) -----------------------------------------------------
// Component `comp` (0 = Y, 1 = Cb = B-G, 2 = Cr = R-G) of 8 level-shifted RGB pixels: the reversible transform for 5/3
// (Engines/CodingEngine.cu:357-380), the irreversible one for 9/7 in the operation order of the reference binary
// (t = G*m1; t = fma(R, m0, t); t = fma(B, m2, t), see color.cu), two pixels per packed instruction.
__device__ __forceinline__ void rgb_component(int comp, const int (&r)[8], const int (&g)[8], const int (&b)[8], int (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = comp == 0 ? (r[j] + 2 * g[j] + b[j]) >> 2 : comp == 1 ? b[j] - g[j] : r[j] - g[j];
}
__device__ __forceinline__ void rgb_component(int comp, const float (&r)[8], const float (&g)[8], const float (&b)[8], float (&v)[8]) {
  const float m0 = comp == 0 ? 0.299f : comp == 1 ? -0.168736f : 0.5f;
  const float m1 = comp == 0 ? 0.587f : comp == 1 ? -0.331264f : -0.418688f;
  const float m2 = comp == 0 ? 0.114f : comp == 1 ? 0.5f : -0.081312f;
  const f2 M0 = pk(m0, m0), M1 = pk(m1, m1), M2 = pk(m2, m2);
#pragma unroll
  for (int j = 0; j < 8; j += 2)
    upk(fma2(pk(b[j], b[j + 1]), M2, fma2(pk(r[j], r[j + 1]), M0, mul2(pk(g[j], g[j + 1]), M1))), v[j], v[j + 1]);
}

// One work item = one 256-column strip x one chunk of row pairs of one frame, one level (LL -> scratch or, if last, -> out).
// SRC_U8: the input is the u8 frame (level 1), else the 32-bit LL plane of the level above.
#ifndef PCS_FWD_PF2
#define PCS_FWD_PF2 1
#endif
// 9/7 levels >= 2 (32-bit input): rows prefetched two steps ahead instead of one.  These launches are a single wave of
// long streams and were bound by the bytes in flight (2 KB per warp); with 4 KB per warp at 128 registers (still 4 CTAs
// per SM) the 16 x 4096^2 call goes from 0.433 to 0.41 ms.  The 5/3 variants gain nothing and keep one step.
constexpr bool kPf2On = PCS_FWD_PF2 != 0;
#ifndef PCS_FWD_DEPTH
#define PCS_FWD_DEPTH 6
#endif
constexpr int kFwdDepth = PCS_FWD_DEPTH;
typedef unsigned char FwdRing[kFwdDepth][2][32][8];
#ifndef PCS_FWD_MINB
#define PCS_FWD_MINB 4
#endif
// RGB (u8 input, one level): `frame` is a component PLANE 3f + c of planar RGB frames; the three colour planes of
// frame f are staged side by side (ringp[0..2]) and the colour transform runs in the load.
template <bool LOSSY, typename OT, bool SRC_U8, bool QS1, bool RGB = false>
__device__ __forceinline__ void fwd_item(const FastFwdArgs& a, const int strip, const int chunk, const int frame, FwdRing* const ringp) {
  static_assert(!RGB || SRC_U8, "the fused colour transform reads u8 pixels");
  using T = typename VSel<LOSSY, 8>::T;
  using VP1 = typename VSel<LOSSY, 8>::type;
  using RowT = typename std::conditional<SRC_U8, uint2, Row8<T>>::type;
  constexpr bool kPf2 = kPf2On && LOSSY;
  const int lane = threadIdx.x & 31;
  constexpr int kHalo = 1;                             // halo lanes per side
  constexpr int kEff = (32 - 2 * kHalo) * 8;           // effective columns per strip
  const int Wc = a.Wc, Hc = a.Hc, hw1 = Wc >> 1, hh1 = Hc >> 1;
  const int c0 = strip * kEff - 8 * kHalo + 8 * lane;  // first of the lane's 8 columns
  const bool inimg = c0 >= 0 && c0 < Wc;
  const bool wlane = lane >= kHalo && lane < 32 - kHalo && inimg;
  const bool left = c0 == 0, right = c0 + 8 == Wc;
  // lanes outside the image read inside it; their values are never used (explicit border rules)
  const int cl = min(max(c0, 0), Wc - 8);
  const int sframe = RGB ? frame / 3 : frame, comp = RGB ? frame - 3 * sframe : 0;  // source frame, colour component
  const unsigned char* const src8 = static_cast<const unsigned char*>(a.src) + (size_t)sframe * a.src_fstride + cl;
  const T* const srcT = static_cast<const T*>(a.src) + (size_t)frame * a.src_fstride + cl;
  OT* const out = static_cast<OT*>(a.out) + (size_t)frame * a.out_fstride;
  T* const llbase = a.last ? nullptr : static_cast<T*>(a.ll) + (size_t)frame * a.ll_fstride;
  const float qs = a.qs;

  // row pairs [n0, n1) are written by this chunk
  const int n0 = chunk * a.rp, n1 = min(n0 + a.rp, hh1);
  const int P0 = n0, Plast = n1 - 1;  // first / last pair whose output is needed

  // Row staging.  u8 input (level 1, the bandwidth-critical launch): every lane cp.async-copies its own 8
  // bytes of the rows of the next kDepth steps into a per-warp shared-memory ring and reads them back
  // itself (no barrier needed); the deep prefetch hides the DRAM latency at 16 warps per SM.  32-bit
  // input (deeper levels, small): plain register prefetch one step ahead.
  constexpr int kDepth = kFwdDepth;
  FwdRing& ring = *ringp;
  auto row_index = [&](int y) -> int {  // whole-sample fold at the borders of the level input, then the pad mapping
    int ys = y < 0 ? -y : (y >= Hc ? 2 * (Hc - 1) - y : y);
    if constexpr (SRC_U8) ys = ys < a.H ? ys : 2 * a.H - 1 - ys;
    return ys;
  };
  auto issue_step = [&](int m, int slot) {  // rows 2m+1, 2m+2 (rows past the chunk fold to valid rows and are not used)
    const unsigned char* const r1 = src8 + (size_t)row_index(2 * m + 1) * a.src_pitch;
    const unsigned char* const r2 = src8 + (size_t)row_index(2 * m + 2) * a.src_pitch;
#pragma unroll
    for (int k = 0; k < (RGB ? 3 : 1); ++k) {
      cp_async8(&ringp[k][slot][0][lane][0], r1 + (size_t)k * a.comp_stride);
      cp_async8(&ringp[k][slot][1][lane][0], r2 + (size_t)k * a.comp_stride);
    }
    cp_async_commit();
  };
  // level-1 samples of a staged row (r = 0 / 1: the odd / even row of the step in ring slot `slot`)
  auto ring_row = [&](int slot, int r, T (&v)[8]) {
    if constexpr (RGB) {
      T cr[8], cg[8], cb[8];
      unpack8(*reinterpret_cast<const uint2*>(&ringp[0][slot][r][lane][0]), cr);
      unpack8(*reinterpret_cast<const uint2*>(&ringp[1][slot][r][lane][0]), cg);
      unpack8(*reinterpret_cast<const uint2*>(&ringp[2][slot][r][lane][0]), cb);
      rgb_component(comp, cr, cg, cb, v);
    } else {
      unpack8(*reinterpret_cast<const uint2*>(&ring[slot][r][lane][0]), v, a.offset);
    }
  };
  auto load_row = [&](int y) -> RowT {
    if constexpr (SRC_U8) {
      return __ldg(reinterpret_cast<const uint2*>(src8 + (size_t)row_index(y) * a.src_pitch));
    } else {
      RowT r;
      const auto* p4 = reinterpret_cast<const decltype(r.a)*>(srcT + (size_t)row_index(y) * a.src_pitch);
      r.a = __ldg(p4); r.b = __ldg(p4 + 1);
      return r;
    }
  };

  VP1 v1;
  const int ms = P0 - VP1::kWarm;          // first step: pair P0 is the first valid one
  const int mend = Plast + VP1::kDelay;    // last step
  RowT pa, pb, pc, pd;  // 32-bit input: rows of step m and (PCS_FWD_PF2) of step m + 1, loaded ahead
  if constexpr (SRC_U8) {
#pragma unroll
    for (int d = 0; d < kDepth; ++d) issue_step(ms + d, d);
  } else {
    pa = load_row(2 * ms + 1); pb = load_row(2 * ms + 2);
    if constexpr (kPf2) { pc = load_row(2 * ms + 3); pd = load_row(2 * ms + 4); }
  }
  {
    T r[8];
    if constexpr (RGB) {
      const unsigned char* const p = src8 + (size_t)row_index(2 * ms) * a.src_pitch;
      T cr[8], cg[8], cb[8];
      unpack8(__ldg(reinterpret_cast<const uint2*>(p)), cr);
      unpack8(__ldg(reinterpret_cast<const uint2*>(p + a.comp_stride)), cg);
      unpack8(__ldg(reinterpret_cast<const uint2*>(p + 2 * a.comp_stride)), cb);
      rgb_component(comp, cr, cg, cb, r);
    } else {
      if constexpr (SRC_U8) unpack8(load_row(2 * ms), r, a.offset);
      else unpack8(load_row(2 * ms), r);
    }
    v1.init(r);
  }
  int slot = 0;

#pragma unroll 2
  for (int m = ms; m <= mend; ++m) {
    T xo[8], x2[8], lo[8], hi[8];
    if constexpr (SRC_U8) {
      cp_async_wait<kDepth - 1>();  // the group of step m has landed
      ring_row(slot, 0, xo);
      ring_row(slot, 1, x2);
      issue_step(m + kDepth, slot);
      slot = slot + 1 == kDepth ? 0 : slot + 1;
    } else {
      unpack8(pa, xo); unpack8(pb, x2);
      if constexpr (kPf2) { pa = pc; pb = pd; pc = load_row(2 * m + 5); pd = load_row(2 * m + 6); }
      else { pa = load_row(2 * m + 3); pb = load_row(2 * m + 4); }
    }
    // borders of this level come from the folded loads, so no explicit rule here
    if constexpr (LOSSY) v97_step_packed<8>(v1, xo, x2, lo, hi);
    else v1.step(xo, x2, false, lo, hi);
    const int pr = m - VP1::kDelay;  // row pair leaving the vertical pipeline
    if (pr < P0) continue;           // warm-up (warp-uniform)
    if constexpr (LOSSY) {
      if (pair_variant(pr, a.nW1, a.ls1)) hfwd97_rows<8, true>(lo, hi, left, right);
      else hfwd97_rows<8, false>(lo, hi, left, right);
    } else {
      hfwd53<8>(lo, left, right);
      hfwd53<8>(hi, left, right);
    }
    if (wlane && pr >= n0 && pr < n1) {
      const int oc1 = c0 >> 1;
      emit_pair<LOSSY, QS1, OT, T, 8>(lo, hi, out + (size_t)pr * a.pitch + oc1, out + (size_t)(pr + hh1) * a.pitch + oc1, hw1, a.q1, qs,
                                 a.last != 0, !a.last ? llbase + (size_t)pr * a.ll_pitch + oc1 : nullptr);
    }
  }
}

#ifndef PCS_FWD_MINB_U8
#define PCS_FWD_MINB_U8 4
#endif
template <bool LOSSY, typename OT, bool SRC_U8, bool QS1, bool RGB = false>
__global__ void __launch_bounds__(128, SRC_U8 ? PCS_FWD_MINB_U8 : PCS_FWD_MINB) dwt_forward_fused(const FastFwdArgs a) {
  __shared__ __align__(16) FwdRing s_ring[SRC_U8 ? 4 : 1][RGB ? 3 : 1];
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= a.nstrips * a.nchunks) return;
  fwd_item<LOSSY, OT, SRC_U8, QS1, RGB>(a, item % a.nstrips, item / a.nstrips, blockIdx.y,
                                                       &s_ring[SRC_U8 ? (threadIdx.x >> 5) : 0][0]);
}


// The chunk length is chosen per kernel variant, for the number of CTAs that variant really keeps resident.
template <bool LOSSY, typename OT, bool SRC_U8, bool QS1, bool RGB = false>
void launch_k(FastFwdArgs a, int frames, cudaStream_t st) {
  static const long long slots = resident_ctas(dwt_forward_fused<LOSSY, OT, SRC_U8, QS1, RGB>);
  const int hh1 = a.Hc >> 1;
  a.rp = choose_rp(a.nstrips, hh1, frames, 3, 4, slots);
  a.nchunks = (hh1 + a.rp - 1) / a.rp;
  dim3 grid((a.nstrips * a.nchunks + 3) / 4, frames);
  dwt_forward_fused<LOSSY, OT, SRC_U8, QS1, RGB><<<grid, 128, 0, st>>>(a);
}
template <bool LOSSY, typename OT, bool SRC_U8, bool RGB = false>
void launch(const FastFwdArgs& a, int frames, cudaStream_t st) {
  if (a.qs == 1.0f) launch_k<LOSSY, OT, SRC_U8, true, RGB>(a, frames, st);
  else launch_k<LOSSY, OT, SRC_U8, false, RGB>(a, frames, st);
}



// =================================================================================================
// inverse
// =================================================================================================
struct FastInvArgs {
  long long comp_stride;   // RGB output: bytes between the R, G and B planes of a frame
  const int32_t* q; long long q_fstride; int pitch;        // quantised Mallat plane, row pitch = adapted frame width
  const void* ll; long long ll_fstride; int ll_pitch;      // LL source (scratch, type T) unless `first`
  int first;                                               // coarsest level: LL comes from q (de-quantised)
  int Wc, Hc;                                              // size of this level's OUTPUT
  void* dst; long long dst_fstride; int dst_pitch;         // scratch plane (T) or the u8 frame
  int H;                                                   // u8 output: rows >= H are padding and are not stored
  float r[4], rqs;                                         // reciprocal quantiser steps LL, HL, LH, HH and 1/qs
  int offset;                                              // u8 output: level shift added back before the clamp
  int nstrips, nchunks, rp;
};


// RGB (u8 output only): a CTA is three warps, warp w reconstructs component w (Y, Cb, Cr) of the SAME strip and chunk of
// frame blockIdx.y; the rows of a step are exchanged through shared memory, warp w then forms colour plane w (R, G, B) --
// the inverse colour transform (Engines/DecodingEngine.cu:599-650) fused into the final store.
template <bool LOSSY, bool DST_U8, bool FIRST, bool QS1, bool RGB = false>
#ifndef PCS_INV_MINB
#define PCS_INV_MINB 4
#endif
#ifndef PCS_INV_RING_ALL   // 1: the cp.async ring also stages the inputs of the 9/7 levels >= 2 (32-bit output)
#define PCS_INV_RING_ALL 0
#endif
__global__ void __launch_bounds__(128, PCS_INV_MINB) dwt_inverse_fast_kernel(const FastInvArgs a) {
  using T = typename VSel<LOSSY, 8>::T;
  using VP = typename IVSel<LOSSY, 8>::type;
  using T4 = typename std::conditional<LOSSY, float4, int4>::type;
  static_assert(!RGB || DST_U8, "the fused inverse colour transform writes u8 pixels");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = RGB ? blockIdx.x : blockIdx.x * (blockDim.x >> 5) + warp;
  if (item >= a.nstrips * a.nchunks) return;  // RGB: the same for the three warps of the CTA
  const int frame = RGB ? 3 * blockIdx.y + warp : blockIdx.y;  // component plane this warp reconstructs
  const int strip = item % a.nstrips, chunk = item / a.nstrips;
  const int Wc = a.Wc, Hc = a.Hc, hw = Wc >> 1, hh = Hc >> 1;
  const int c0 = strip * 240 - 8 + 8 * lane;
  const bool inimg = c0 >= 0 && c0 < Wc;
  const bool wlane = lane >= 1 && lane <= 30 && inimg;
  const bool left = c0 == 0, right = c0 + 8 == Wc;
  const int oc = min(max(c0, 0), Wc - 8) >> 1;  // subband column of the lane's 4 low / 4 high samples
  const int32_t* const q = a.q + (size_t)frame * a.q_fstride + oc;
  const T* const lls = FIRST ? nullptr : static_cast<const T*>(a.ll) + (size_t)frame * a.ll_fstride + oc;

  struct Raw { int4 hl, lh, hh; T4 ll; };
  // row pair n: low row n of (LL|HL), high row n of (LH|HH); whole-sample fold of the row index
  auto pair_rows = [&](int n, int& rl, int& rh) {
    if (n < 0) { rl = -n; rh = -n - 1; }
    else if (n >= hh) { rl = Hc - 1 - n; rh = Hc - 2 - n; }
    else { rl = n; rh = n; }
  };
  auto ll_from_q = [&](int4 t) -> T4 {
    if constexpr (LOSSY) {
      const f2 h0 = pk(0.5f * a.r[0], 0.5f * a.r[0]), rqs2 = pk(a.rqs, a.rqs);
      float4 v;
      upk(dequant2<QS1>(t.x, t.y, h0, rqs2), v.x, v.y);
      upk(dequant2<QS1>(t.z, t.w, h0, rqs2), v.z, v.w);
      return v;
    } else return t;
  };
  auto load_pair = [&](int n) -> Raw {
    int rl, rh;
    pair_rows(n, rl, rh);
    const int32_t* row_lo = q + (size_t)rl * a.pitch;
    const int32_t* row_hi = q + (size_t)(rh + hh) * a.pitch;
    Raw r;
    r.hl = __ldg(reinterpret_cast<const int4*>(row_lo + hw));
    r.lh = __ldg(reinterpret_cast<const int4*>(row_hi));
    r.hh = __ldg(reinterpret_cast<const int4*>(row_hi + hw));
    if constexpr (FIRST) r.ll = ll_from_q(__ldg(reinterpret_cast<const int4*>(row_lo)));
    else r.ll = __ldg(reinterpret_cast<const T4*>(lls + (size_t)rl * a.ll_pitch));
    return r;
  };
  // The finest level (u8 output) moves 4 of the 5 bytes per pixel on its READ side: stage its subband rows
  // through a per-warp cp.async ring kDepth steps deep, as the forward level-1 kernel does.
  constexpr int kDepth = 4;
  constexpr bool kRing = DST_U8 || (PCS_INV_RING_ALL && LOSSY);  // deeper 9/7 levels too: see PCS_INV_RING_ALL
  __shared__ __align__(16) unsigned char s_ring[kRing ? (RGB ? 3 : 4) : 1][kRing ? kDepth : 1][4][32][16];
  auto& ring = s_ring[kRing ? warp : 0];
  auto issue_pair = [&](int n, int slot) {
    int rl, rh;
    pair_rows(n, rl, rh);
    const int32_t* row_lo = q + (size_t)rl * a.pitch;
    const int32_t* row_hi = q + (size_t)(rh + hh) * a.pitch;
    cp_async16(&ring[slot][0][lane][0], row_lo + hw);
    cp_async16(&ring[slot][1][lane][0], row_hi);
    cp_async16(&ring[slot][2][lane][0], row_hi + hw);
    if constexpr (FIRST) cp_async16(&ring[slot][3][lane][0], row_lo);
    else cp_async16(&ring[slot][3][lane][0], lls + (size_t)rl * a.ll_pitch);
    cp_async_commit();
  };
  auto read_pair = [&](int slot) -> Raw {
    Raw r;
    r.hl = *reinterpret_cast<const int4*>(&ring[slot][0][lane][0]);
    r.lh = *reinterpret_cast<const int4*>(&ring[slot][1][lane][0]);
    r.hh = *reinterpret_cast<const int4*>(&ring[slot][2][lane][0]);
    if constexpr (FIRST) r.ll = ll_from_q(*reinterpret_cast<const int4*>(&ring[slot][3][lane][0]));
    else r.ll = *reinterpret_cast<const T4*>(&ring[slot][3][lane][0]);
    return r;
  };
  auto expand = [&](const Raw& r, T (&lo)[8], T (&hi)[8]) {
    lo[0] = r.ll.x; lo[2] = r.ll.y; lo[4] = r.ll.z; lo[6] = r.ll.w;
    if constexpr (LOSSY) {
      // odd columns: HL (low row) and HH (high row) of a column form the pair the horizontal pass works on
      const f2 h13 = pk(0.5f * a.r[1], 0.5f * a.r[3]), h22 = pk(0.5f * a.r[2], 0.5f * a.r[2]), rqs2 = pk(a.rqs, a.rqs);
      upk(dequant2<QS1>(r.hl.x, r.hh.x, h13, rqs2), lo[1], hi[1]);
      upk(dequant2<QS1>(r.hl.y, r.hh.y, h13, rqs2), lo[3], hi[3]);
      upk(dequant2<QS1>(r.hl.z, r.hh.z, h13, rqs2), lo[5], hi[5]);
      upk(dequant2<QS1>(r.hl.w, r.hh.w, h13, rqs2), lo[7], hi[7]);
      upk(dequant2<QS1>(r.lh.x, r.lh.y, h22, rqs2), hi[0], hi[2]);
      upk(dequant2<QS1>(r.lh.z, r.lh.w, h22, rqs2), hi[4], hi[6]);
    } else {
      lo[1] = r.hl.x; lo[3] = r.hl.y; lo[5] = r.hl.z; lo[7] = r.hl.w;
      hi[0] = r.lh.x; hi[2] = r.lh.y; hi[4] = r.lh.z; hi[6] = r.lh.w;
      hi[1] = r.hh.x; hi[3] = r.hh.y; hi[5] = r.hh.z; hi[7] = r.hh.w;
    }
  };
  auto store_row = [&](int y, const T (&v)[8]) {
    if constexpr (DST_U8) {
      if (y >= a.H) return;
      unsigned b[8];
      if constexpr (LOSSY) {  // clamp(rn((v + offset) + 0.01)), DecodingEngine.cu:714
        const float foff = (float)a.offset;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float t0, t1;
          upk(add2(add2(pk(v[j], v[j + 1]), pk(foff, foff)), pk(0.01f, 0.01f)), t0, t1);
          b[j] = sat_u8(t0); b[j + 1] = sat_u8(t1);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = (unsigned)min(max(v[j] + a.offset, 0), 255);  // :728
      }
      uint2 w;
      w.x = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
      w.y = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
      __stcs(reinterpret_cast<uint2*>(static_cast<unsigned char*>(a.dst) + (size_t)frame * a.dst_fstride + (size_t)y * a.dst_pitch + c0), w);
    } else {
      T* d = static_cast<T*>(a.dst) + (size_t)frame * a.dst_fstride + (size_t)y * a.dst_pitch + c0;
      T4 t0, t1;
      t0.x = v[0]; t0.y = v[1]; t0.z = v[2]; t0.w = v[3]; t1.x = v[4]; t1.y = v[5]; t1.z = v[6]; t1.w = v[7];
      st4_keep<T>(d, t0.x, t0.y, t0.z, t0.w); st4_keep<T>(d + 4, t1.x, t1.y, t1.z, t1.w);  // read back by the next finer level
    }
  };

  // RGB: exchange of a step's two rows between the three component warps (double buffered: one barrier per step)
  __shared__ __align__(16) T s_x[RGB ? 2 : 1][RGB ? 2 : 1][RGB ? 3 : 1][32][8];
  auto colour_row = [&](int y, const T (&cy)[8], const T (&cb)[8], const T (&cr)[8]) {  // colour plane `warp` of row y
    if (y >= a.H) return;
    unsigned b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int v;
      if constexpr (LOSSY) {  // DecodingEngine.cu:620-640 in the operation order of color.cu
        const float t = warp == 0 ? __fmaf_rn(cr[j], 1.402f, cy[j])
                      : warp == 1 ? __fmaf_rn(cr[j], -0.714136f, __fmaf_rn(cb[j], -0.344136f, cy[j]))
                                  : __fmaf_rn(cb[j], 1.772f, cy[j]);
        v = __float2int_rn(__fadd_rn(t, 0.01f));
      } else {
        const int g = cy[j] - ((cb[j] + cr[j]) >> 2);
        v = warp == 0 ? cr[j] + g : warp == 1 ? g : cb[j] + g;
      }
      b[j] = (unsigned)min(max(v + 128, 0), 255);
    }
    uint2 w;
    w.x = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
    w.y = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
    __stcs(reinterpret_cast<uint2*>(static_cast<unsigned char*>(a.dst) + (size_t)blockIdx.y * a.dst_fstride + (size_t)warp * a.comp_stride +
                                    (size_t)y * a.dst_pitch + c0), w);
  };

  const int n0 = chunk * a.rp, n1 = min(n0 + a.rp, hh);
  VP vp; vp.init();
  const int ms = n0 - VP::kLead, mend = n1 + VP::kTail;
  Raw cur;
  int slot = 0;
  if constexpr (kRing) {
#pragma unroll
    for (int d = 0; d < kDepth; ++d) issue_pair(ms + d, d);
  } else {
    cur = load_pair(ms);
  }
#pragma unroll 2
  for (int m = ms; m <= mend; ++m) {
    T lo[8], hi[8], re[8], ro[8];
    if constexpr (kRing) {
      cp_async_wait<kDepth - 1>();  // pair m has landed
      cur = read_pair(slot);
      issue_pair(m + kDepth, slot);  // pairs past the chunk fold to valid rows and are not used
      slot = slot + 1 == kDepth ? 0 : slot + 1;
      expand(cur, lo, hi);
    } else {
      expand(cur, lo, hi);
      cur = load_pair(m + 1);  // prefetch
    }
    if constexpr (LOSSY) { hinv97_rows<8>(lo, hi, left, right); iv97_step_packed<8>(vp, lo, hi, re, ro); }
    else { hinv53<8>(lo, left, right); hinv53<8>(hi, left, right); vp.step(lo, hi, re, ro); }
    const int yo = VP::odd_row(m), ye = VP::even_row(m);
    if constexpr (RGB) {
      const int buf = m & 1;
      T4* const mine_o = reinterpret_cast<T4*>(&s_x[buf][0][warp][lane][0]);
      T4* const mine_e = reinterpret_cast<T4*>(&s_x[buf][1][warp][lane][0]);
      T4 t;
      t.x = ro[0]; t.y = ro[1]; t.z = ro[2]; t.w = ro[3]; mine_o[0] = t;
      t.x = ro[4]; t.y = ro[5]; t.z = ro[6]; t.w = ro[7]; mine_o[1] = t;
      t.x = re[0]; t.y = re[1]; t.z = re[2]; t.w = re[3]; mine_e[0] = t;
      t.x = re[4]; t.y = re[5]; t.z = re[6]; t.w = re[7]; mine_e[1] = t;
      __syncthreads();  // every warp of the CTA runs the same steps (same strip and chunk), so this is uniform
      if (wlane) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int y = r ? ye : yo;
          if (y < 2 * n0 || y >= 2 * n1) continue;
          T cy[8], cb[8], cr[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { cy[j] = s_x[buf][r][0][lane][j]; cb[j] = s_x[buf][r][1][lane][j]; cr[j] = s_x[buf][r][2][lane][j]; }
          colour_row(y, cy, cb, cr);
        }
      }
    } else if (wlane) {
      if (yo >= 2 * n0 && yo < 2 * n1) store_row(yo, ro);
      if (ye >= 2 * n0 && ye < 2 * n1) store_row(ye, re);
    }
  }
}

// The chunk length is chosen per kernel variant: a grid sized for 4 CTAs per SM runs 1.3 waves on a variant that
// only fits 3 (the 9/7 inverse needs 130-138 registers), i.e. a second wave that is 30 % full.
template <bool LOSSY, bool DST_U8, bool FIRST, bool QS1>
void launch_inv_k(FastInvArgs a, int frames, cudaStream_t st) {
  static const long long slots = resident_ctas(dwt_inverse_fast_kernel<LOSSY, DST_U8, FIRST, QS1>);
  const int hh = a.Hc >> 1;
  a.rp = choose_rp(a.nstrips, hh, frames, 4, 2, slots);
  a.nchunks = (hh + a.rp - 1) / a.rp;
  dim3 grid((a.nstrips * a.nchunks + 3) / 4, frames);
  dwt_inverse_fast_kernel<LOSSY, DST_U8, FIRST, QS1><<<grid, 128, 0, st>>>(a);
}
// RGB: one item (three component warps) per CTA of 96 threads, grid.y counts frames
template <bool LOSSY, bool FIRST, bool QS1>
void launch_inv_rgb_k(FastInvArgs a, int frames, cudaStream_t st) {
  static const long long slots = [] {
    int dev = 0, sms = 148, per_sm = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dwt_inverse_fast_kernel<LOSSY, true, FIRST, QS1, true>, 96, 0);
    return (long long)sms * std::max(per_sm, 1);
  }();
  const int hh = a.Hc >> 1;
  // choose_rp counts CTAs of four items: present a CTA of one item as four times as many slots
  a.rp = choose_rp(a.nstrips, hh, frames, 4, 2, slots / 4 > 0 ? slots / 4 : 1);
  a.nchunks = (hh + a.rp - 1) / a.rp;
  dim3 grid(a.nstrips * a.nchunks, frames);
  dwt_inverse_fast_kernel<LOSSY, true, FIRST, QS1, true><<<grid, 96, 0, st>>>(a);
}

template <bool LOSSY, bool DST_U8, bool FIRST>
void launch_inv(const FastInvArgs& a, int frames, bool qs1, cudaStream_t st) {
  if (qs1) launch_inv_k<LOSSY, DST_U8, FIRST, true>(a, frames, st);
  else launch_inv_k<LOSSY, DST_U8, FIRST, false>(a, frames, st);
}

}  // namespace

// Can the streaming kernel take level `level` (1-based) of this geometry?  src = its input.
bool dwt_forward_fast_ok(const void* src, const void* d_coeff, const void* ll, bool src_u8, int src_pitch, long long src_fstride, int H,
                         int Wc, int Hc) {
  if (Wc % 8 || Hc % 2 || Wc < 64 || Hc < 16) return false;
  if (reinterpret_cast<uintptr_t>(d_coeff) & 15) return false;
  if (ll && (reinterpret_cast<uintptr_t>(ll) & 15)) return false;
  if (src_u8) {
    if ((reinterpret_cast<uintptr_t>(src) & 7) || src_pitch % 8 || src_fstride % 8 || src_pitch != Wc || H > Hc) return false;
  } else {
    if ((reinterpret_cast<uintptr_t>(src) & 15) || src_pitch % 4 || src_fstride % 4) return false;
  }
  return true;
}

// Level `level` (1-based) of the forward transform in one launch.
int dwt_forward_fast(const void* src, bool src_u8, long long src_fstride, int src_pitch, int H, int Wc, int Hc, void* d_coeff,
                     long long out_fstride, int pitch, void* d_ll, long long ll_fstride, int ll_pitch, int level, int wl,
                     bool lossy, float qs, bool out_int, int frames, cudaStream_t st, bool rgb, long long comp_stride, int offset) {
  // rgb: src = planar u8 RGB frames (src_fstride per FRAME, comp_stride between colour planes), `frames` counts
  // component planes (3 per frame) and the colour transform is fused into the load
  FastFwdArgs a{};
  a.comp_stride = comp_stride; a.offset = offset;
  a.src = src; a.src_fstride = src_fstride; a.src_pitch = src_pitch; a.H = src_u8 ? H : Hc; a.Wc = Wc; a.Hc = Hc;
  a.out = d_coeff; a.out_fstride = out_fstride; a.pitch = pitch;
  a.ll = d_ll; a.ll_fstride = ll_fstride; a.ll_pitch = ll_pitch;
  a.last = (level == wl);
  for (int i = 0; i < 4; ++i) a.q1[i] = kQ[level - 1][i];
  a.qs = qs;
  variant_params(Hc, &a.nW1, &a.ls1);
  a.nstrips = (Wc + 239) / 240;  // 30 working lanes x 8 columns per strip
  if (src_u8) {
    if (rgb) {
      if (!lossy) launch<false, int, true, true>(a, frames, st);
      else if (out_int) launch<true, int, true, true>(a, frames, st);
      else launch<true, float, true, true>(a, frames, st);
    } else {
      if (!lossy) launch<false, int, true>(a, frames, st);
      else if (out_int) launch<true, int, true>(a, frames, st);
      else launch<true, float, true>(a, frames, st);
    }
  } else {
    if (!lossy) launch<false, int, false>(a, frames, st);
    else if (out_int) launch<true, int, false>(a, frames, st);
    else launch<true, float, false>(a, frames, st);
  }
  ++g_launches;
  return (int)cudaGetLastError();
}


// Can the streaming inverse kernel produce the Wc x Hc output of one level?
bool dwt_inverse_fast_ok(const void* q, const void* ll, const void* dst, bool dst_u8, int dst_pitch, long long dst_fstride, int ll_pitch,
                         long long ll_fstride, int Wc, int Hc) {
  if (Wc % 8 || Hc % 2 || Wc < 64 || Hc < 16) return false;
  if (reinterpret_cast<uintptr_t>(q) & 15) return false;
  if (ll && ((reinterpret_cast<uintptr_t>(ll) & 15) || ll_pitch % 4 || ll_fstride % 4)) return false;
  if (dst_u8) return !(reinterpret_cast<uintptr_t>(dst) & 7) && dst_pitch == Wc && dst_fstride % 8 == 0;
  return !(reinterpret_cast<uintptr_t>(dst) & 15) && dst_pitch % 4 == 0 && dst_fstride % 4 == 0;
}

int dwt_inverse_fast(const int32_t* d_q, long long q_fstride, int pitch, const void* ll, long long ll_fstride, int ll_pitch, bool first,
                     int Wc, int Hc, void* dst, bool dst_u8, long long dst_fstride, int dst_pitch, int H, int level, bool lossy, float qs,
                     int frames, cudaStream_t st, bool rgb, long long comp_stride, int offset) {
  // rgb (dst_u8 only): dst = planar u8 RGB frames (dst_fstride per FRAME, comp_stride between colour planes), `frames`
  // counts frames, d_q / ll hold three component planes per frame
  FastInvArgs a{};
  a.comp_stride = comp_stride; a.offset = offset;
  a.q = d_q; a.q_fstride = q_fstride; a.pitch = pitch;
  a.ll = ll; a.ll_fstride = ll_fstride; a.ll_pitch = ll_pitch; a.first = first;
  a.Wc = Wc; a.Hc = Hc; a.dst = dst; a.dst_fstride = dst_fstride; a.dst_pitch = dst_pitch; a.H = dst_u8 ? H : Hc;
  for (int i = 0; i < 4; ++i) a.r[i] = 1.0f / kQ[level - 1][i];
  a.rqs = 1.0f / qs;
  a.nstrips = (Wc + 239) / 240;
  const bool qs1 = a.rqs == 1.0f;
  if (rgb && dst_u8) {
    if (!lossy) { if (first) launch_inv_rgb_k<false, true, true>(a, frames, st); else launch_inv_rgb_k<false, false, true>(a, frames, st); }
    else if (qs1) { if (first) launch_inv_rgb_k<true, true, true>(a, frames, st); else launch_inv_rgb_k<true, false, true>(a, frames, st); }
    else { if (first) launch_inv_rgb_k<true, true, false>(a, frames, st); else launch_inv_rgb_k<true, false, false>(a, frames, st); }
    ++g_launches;
    return (int)cudaGetLastError();
  }
  if (!lossy) {
    if (dst_u8) { if (first) launch_inv<false, true, true>(a, frames, qs1, st); else launch_inv<false, true, false>(a, frames, qs1, st); }
    else { if (first) launch_inv<false, false, true>(a, frames, qs1, st); else launch_inv<false, false, false>(a, frames, qs1, st); }
  } else {
    if (dst_u8) { if (first) launch_inv<true, true, true>(a, frames, qs1, st); else launch_inv<true, true, false>(a, frames, qs1, st); }
    else { if (first) launch_inv<true, false, true>(a, frames, qs1, st); else launch_inv<true, false, false>(a, frames, qs1, st); }
  }
  ++g_launches;
  return (int)cudaGetLastError();
}

}  // namespace pcs
