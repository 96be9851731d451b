// dwt.cu -- streaming lifting DWT for sm_100a (forward + inverse, reversible 5/3 and irreversible 9/7).
//
// Replaces the reference's DWT stage: DWTEngine<T,Y>::DWTForward / ::DWTReverse and their kernels
// (DWT/DWTGenerator.cu:753-1124, :1268-1424 of 13Karl/CUDA-Image-and-Video-codec), plus the pixel
// pre/post kernels offsetImage (Engines/CodingEngine.cu:581-588) and
// removeOffsetAndApplyMaxMin[Lossy] (Engines/DecodingEngine.cu:706-729) and the host-side mirror
// padding / u8 narrowing (IO/IOManager.ipp:72-112, :332-335), which are fused into the level-1 load
// and store here.
//
// Design (not the reference's 64x18 register tile with 1.4-2x overlap re-reads):
//   * one warp owns a 128-column strip (120 effective + one 4-sample halo lane on each side) and
//     STREAMS down a chunk of rows, keeping only the lifting pipeline state in registers
//     (2 rows for 5/3, 4 for 9/7): every input row is read once per strip, no vertical overlap
//     except a 2..4 row warm-up per chunk;
//   * horizontal lifting is done on the 4 samples a lane holds plus one shuffle per lifting step;
//   * loads are 4-byte (u8 level) / 16-byte (int32/f32 levels) per lane, stores 8 bytes per lane
//     per subband row, all coalesced; no shared memory;
//   * image borders use whole-sample symmetric extension by index folding at load time, which is
//     bit-identical to the reference's mirrored lifting steps (SURVEY.md Annex A.2/A.3).
// 9/7 arithmetic uses explicit __fmaf_rn/__fadd_rn/__fmul_rn so that the operation shapes are
// exactly those of the reference binary, including its one cross-statement FMA contraction
// (vertical scale fused into horizontal step one on odd columns; see oracle/picsong_oracle.c).
#include <algorithm>
#include <cstdlib>

#include "pcs_common.cuh"

namespace pcs {

namespace {

constexpr float kA = -1.586134342059924f, kB = -0.052980118572961f, kG = 0.882911075530934f,
                kD = 0.443506852043971f, kK = 1.230174104914001f, kIK = 0.812893066f;
constexpr int kStrip = 120;  // effective columns per warp strip (lanes 1..30 x 4 samples)

struct FwdArgs {
  const void* src; long long src_fstride; int src_pitch;  // level input
  int W, H;                                              // original dims (u8 level: pad mapping)
  int Wc, Hc;                                            // level input size
  void* out; long long out_fstride; int Wa;              // Mallat plane
  void* ll; long long ll_fstride; int ll_pitch;          // LL destination
  int last, src_vec, out_vec;
  float q_ll, q_hl, q_lh, q_hh, qs;
  int rp, nstrips, nchunks;
  int offset;                                            // u8 level: level shift 2^(bps-1), 0 for signed samples
};

struct InvArgs {
  const int32_t* q; long long q_fstride; int Wa;         // quantised Mallat plane
  const void* ll; long long ll_fstride; int ll_pitch;    // LL source (scratch) when !first
  int first;                                             // coarsest level: LL from q
  int Wc, Hc;                                            // level OUTPUT size
  void* dst; long long dst_fstride; int dst_pitch;       // scratch (T) or u8 pixels
  int W, H, flat_crop;                                   // final crop
  int in_vec, dst_vec;
  float r_ll, r_hl, r_lh, r_hh, rqs;                     // reciprocal quantiser steps
  int rp, nstrips, nchunks;
  int offset;                                            // u8 output: level shift added back before the clamp
};

template <typename T> struct Vec4;
template <> struct Vec4<int> { using type = int4; };
template <> struct Vec4<float> { using type = float4; };
template <typename T> struct Vec2;
template <> struct Vec2<int> { using type = int2; };
template <> struct Vec2<float> { using type = float2; };

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------

template <typename T, bool SRC_U8>
struct RowLoader {
  const unsigned char* base8; const T* baseT;
  int pitch, Hc, H, c0; int col[4]; bool vec; int offset;
  __device__ __forceinline__ void load(int y, T v[4]) const {
    int ys = fold_ws(y, Hc);
    if (SRC_U8) {
      ys = pad_src(ys, H);
      const unsigned char* r = base8 + (size_t)ys * pitch;
      if (vec) {
        uchar4 p = __ldg(reinterpret_cast<const uchar4*>(r + c0));
        v[0] = (T)((int)p.x - offset); v[1] = (T)((int)p.y - offset); v[2] = (T)((int)p.z - offset); v[3] = (T)((int)p.w - offset);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (T)((int)__ldg(r + col[j]) - offset);
      }
    } else {
      const T* r = baseT + (size_t)ys * pitch;
      if (vec) {
        typename Vec4<T>::type p = __ldg(reinterpret_cast<const typename Vec4<T>::type*>(r + c0));
        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldg(r + col[j]);
      }
    }
  }
};

// 5/3 horizontal analysis of 4 samples (DWTGenerator.cu:72-76,279-292): in x = e0,o0,e1,o1; out the same
// slots hold s0,d0,s1,d1.
__device__ __forceinline__ void hfwd53(int x[4]) {
  const int nx0 = __shfl_down_sync(0xffffffffu, x[0], 1);
  const int d0 = x[1] - ((x[0] + x[2]) >> 1);
  const int d1 = x[3] - ((x[2] + nx0) >> 1);
  const int pd1 = __shfl_up_sync(0xffffffffu, d1, 1);
  x[0] = x[0] + ((pd1 + d0 + 2) >> 2);
  x[2] = x[2] + ((d0 + d1 + 2) >> 2);
  x[1] = d0; x[3] = d1;
}

// 9/7 horizontal analysis (DWTGenerator.cu:91-104,311-323) of one row whose vertical-pass scale S is
// still pending: even columns are scaled first, odd columns get S fused into step one.  `fused_sum`
// selects the second contraction the reference binary contains for its tile rows 12..17
// (sum = fma(v, S, right) instead of v*S + right; see fwd97_row_variant).
// `left_edge`: the lane holding image column 0 takes o[-1] := o[0] (the reference's clamped shuffle,
// DWTGenerator.cu:317,319) instead of the halo lane's recomputation, which is not bitwise symmetric
// when the sum is fused.  Likewise at the right edge e[Wc] := e[Wc-2]: `right1` = the lane's second
// pair is the image's last column pair, `right0` = its first pair is (then the second pair is mirror data).
__device__ __forceinline__ void hfwd97(float x[4], float S, bool fused_sum, bool left_edge, bool right0, bool right1) {
  float e0 = __fmul_rn(x[0], S), e1 = __fmul_rn(x[2], S);
  float n0 = __shfl_down_sync(0xffffffffu, e0, 1);
  const float s0 = fused_sum ? __fmaf_rn(x[0], S, e1) : __fadd_rn(e0, e1);
  const float s1 = fused_sum ? __fmaf_rn(x[2], S, n0) : __fadd_rn(e1, n0);
  float o0 = __fmaf_rn(x[1], S, __fmul_rn(s0, kA));
  float o1 = __fmaf_rn(x[3], S, __fmul_rn(s1, kA));
  float p1 = __shfl_up_sync(0xffffffffu, o1, 1);
  if (left_edge) p1 = o0;
  e0 = __fmaf_rn(__fadd_rn(p1, o0), kB, e0);
  e1 = __fmaf_rn(__fadd_rn(o0, o1), kB, e1);
  n0 = __shfl_down_sync(0xffffffffu, e0, 1);
  if (right0) e1 = e0;
  if (right1) n0 = e1;
  o0 = __fmaf_rn(__fadd_rn(e0, e1), kG, o0);
  o1 = __fmaf_rn(__fadd_rn(e1, n0), kG, o1);
  p1 = __shfl_up_sync(0xffffffffu, o1, 1);
  if (left_edge) p1 = o0;
  e0 = __fmul_rn(__fmaf_rn(__fadd_rn(p1, o0), kD, e0), kIK);
  e1 = __fmul_rn(__fmaf_rn(__fadd_rn(o0, o1), kD, e1), kIK);
  x[0] = e0; x[2] = e1;
  x[1] = __fmul_rn(o0, kK); x[3] = __fmul_rn(o1, kK);
}

// Which image rows the reference computes with the fused neighbour sum: its kernel is unrolled over a
// 64x18 register tile and nvcc made that contraction for tile rows 12..17 only.  Tile-local row index
// under the reference's vertical tiling (DWTGenerator.cu:699-744,1281-1285): a tile every 10 rows; the
// top tile writes local rows 0..13, middle tiles 4..13, the bottom tile (aligned to the image bottom)
// 4..17.  Rows written by both a middle and the bottom tile are a write race in the reference; the
// bottom tile is modelled (oracle/picsong_oracle.c fwd97_row_variant).
__device__ __forceinline__ bool fwd97_row_variant(int y, int Hc) {
  const int nWy = (Hc - 8 + 9) / 10;
  if (nWy <= 1) return y >= 12;
  const int last_start = (nWy - 1) * 10 - ((10 - ((Hc - 8) % 10)) % 10);
  if (y >= last_start + 4) return (y - last_start) >= 12;
  if (y < 14) return y >= 12;
  return (y - 4) % 10 >= 8;
}

template <typename OT> __device__ __forceinline__ OT cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ int cvt_out<int>(float v) { return __float2int_rz(v); }  // BPCEngine.cu:49

template <typename OT>
__device__ __forceinline__ void store2(OT* p, OT a, OT b, bool va, bool vb, bool vec) {
  if (vec && va && vb) {
    typename Vec2<OT>::type t; t.x = a; t.y = b;
    *reinterpret_cast<typename Vec2<OT>::type*>(p) = t;
  } else {
    if (va) p[0] = a;
    if (vb) p[1] = b;
  }
}

// One level of the forward transform.  T = lifting type (int / float), OT = subband output type
// (int for 5/3; float or int (truncated quantisation index) for 9/7).
template <typename T, typename OT, bool LOSSY, bool SRC_U8>
__global__ void __launch_bounds__(128) dwt_forward_level(const FwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= a.nstrips * a.nchunks) return;
  const int frame = blockIdx.y;
  const int strip = item % a.nstrips, chunk = item / a.nstrips;
  const int Wc = a.Wc, Hc = a.Hc, half_w = Wc >> 1, half_h = Hc >> 1;

  RowLoader<T, SRC_U8> ld;
  ld.pitch = a.src_pitch; ld.Hc = Hc; ld.H = a.H;
  ld.c0 = strip * kStrip - 4 + 4 * lane;
  ld.base8 = SRC_U8 ? static_cast<const unsigned char*>(a.src) + (size_t)frame * a.src_fstride : nullptr;
  ld.baseT = SRC_U8 ? nullptr : static_cast<const T*>(a.src) + (size_t)frame * a.src_fstride;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = fold_ws(ld.c0 + j, Wc);
    ld.col[j] = SRC_U8 ? pad_src(c, a.W) : c;
  }
  ld.vec = a.src_vec && ld.c0 >= 0 && ld.c0 + 3 < (SRC_U8 ? min(Wc, a.W) : Wc);
  ld.offset = a.offset;

  const bool wlane = lane >= 1 && lane <= 30;
  const bool v0 = wlane && ld.c0 < Wc, v1 = wlane && ld.c0 + 2 < Wc;
  const int oc = ld.c0 >> 1;  // column inside each subband
  OT* const out = static_cast<OT*>(a.out) + (size_t)frame * a.out_fstride;
  OT* const llb = a.last ? out : nullptr;
  T* const lls = a.last ? nullptr : static_cast<T*>(a.ll) + (size_t)frame * a.ll_fstride;
  const bool ovec = a.out_vec != 0;

  const int n0 = chunk * a.rp, n1 = min(n0 + a.rp, half_h);

  // Writes the two subband rows produced for row pair n: lo = vertical-low row (LL|HL), hi = (LH|HH).
  auto emit = [&](int n, T lo[4], T hi[4]) {
    if (LOSSY) {
      const float* l = reinterpret_cast<const float*>(lo);
      const float* h = reinterpret_cast<const float*>(hi);
      OT* row_lo = out + (size_t)n * a.Wa;
      OT* row_hi = out + (size_t)(n + half_h) * a.Wa;
      if (a.last) {
        store2<OT>(row_lo + oc, cvt_out<OT>(__fmul_rn(__fmul_rn(l[0], a.q_ll), a.qs)),
                   cvt_out<OT>(__fmul_rn(__fmul_rn(l[2], a.q_ll), a.qs)), v0, v1, ovec);
      } else {
        store2<float>(reinterpret_cast<float*>(lls) + (size_t)n * a.ll_pitch + oc, l[0], l[2], v0, v1, ovec);
      }
      store2<OT>(row_lo + half_w + oc, cvt_out<OT>(__fmul_rn(__fmul_rn(l[1], a.q_hl), a.qs)),
                 cvt_out<OT>(__fmul_rn(__fmul_rn(l[3], a.q_hl), a.qs)), v0, v1, ovec);
      store2<OT>(row_hi + oc, cvt_out<OT>(__fmul_rn(__fmul_rn(h[0], a.q_lh), a.qs)),
                 cvt_out<OT>(__fmul_rn(__fmul_rn(h[2], a.q_lh), a.qs)), v0, v1, ovec);
      store2<OT>(row_hi + half_w + oc, cvt_out<OT>(__fmul_rn(__fmul_rn(h[1], a.q_hh), a.qs)),
                 cvt_out<OT>(__fmul_rn(__fmul_rn(h[3], a.q_hh), a.qs)), v0, v1, ovec);
    } else {
      const int* l = reinterpret_cast<const int*>(lo);
      const int* h = reinterpret_cast<const int*>(hi);
      int* row_lo = reinterpret_cast<int*>(out) + (size_t)n * a.Wa;
      int* row_hi = reinterpret_cast<int*>(out) + (size_t)(n + half_h) * a.Wa;
      if (a.last) store2<int>(row_lo + oc, l[0], l[2], v0, v1, ovec);
      else store2<int>(reinterpret_cast<int*>(lls) + (size_t)n * a.ll_pitch + oc, l[0], l[2], v0, v1, ovec);
      store2<int>(row_lo + half_w + oc, l[1], l[3], v0, v1, ovec);
      store2<int>(row_hi + oc, h[0], h[2], v0, v1, ovec);
      store2<int>(row_hi + half_w + oc, h[1], h[3], v0, v1, ovec);
    }
    (void)llb;
  };

  if (!LOSSY) {
    // d[n] = x[2n+1] - ((x[2n]+x[2n+2])>>1);  s[n] = x[2n] + ((d[n-1]+d[n]+2)>>2)   (DWTGenerator.cu:72-76,137-157)
    int xe[4], dp[4], t0[4], t1[4];
    ld.load(2 * n0 - 2, reinterpret_cast<T*>(t0));
    ld.load(2 * n0 - 1, reinterpret_cast<T*>(t1));
    ld.load(2 * n0, reinterpret_cast<T*>(xe));
#pragma unroll
    for (int j = 0; j < 4; ++j) dp[j] = t1[j] - ((t0[j] + xe[j]) >> 1);
    int xo[4], x2[4];
    ld.load(2 * n0 + 1, reinterpret_cast<T*>(xo));
    ld.load(2 * n0 + 2, reinterpret_cast<T*>(x2));
    for (int n = n0; n < n1; ++n) {
      int no[4], n2[4];
      if (n + 1 < n1) {  // prefetch the next row pair while this one is filtered
        ld.load(2 * n + 3, reinterpret_cast<T*>(no));
        ld.load(2 * n + 4, reinterpret_cast<T*>(n2));
      }
      int s[4], d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        d[j] = xo[j] - ((xe[j] + x2[j]) >> 1);
        s[j] = xe[j] + ((dp[j] + d[j] + 2) >> 2);
        dp[j] = d[j]; xe[j] = x2[j];
      }
      hfwd53(s);
      hfwd53(d);
      emit(n, reinterpret_cast<T*>(s), reinterpret_cast<T*>(d));
#pragma unroll
      for (int j = 0; j < 4; ++j) { xo[j] = no[j]; x2[j] = n2[j]; }
    }
  } else {
    // 9/7 pipeline (DWTGenerator.cu:91-104,184-227): at step m rows 2m+1, 2m+2 arrive and row pair m-1 leaves.
    float xe[4], o1p[4] = {0, 0, 0, 0}, e1p[4] = {0, 0, 0, 0}, o2pp[4] = {0, 0, 0, 0};
    ld.load(2 * (n0 - 2), reinterpret_cast<T*>(xe));
    float xo[4], x2[4];
    ld.load(2 * (n0 - 2) + 1, reinterpret_cast<T*>(xo));
    ld.load(2 * (n0 - 2) + 2, reinterpret_cast<T*>(x2));
    for (int m = n0 - 2; m <= n1; ++m) {
      float no[4], n2[4];
      if (m < n1) {
        ld.load(2 * m + 3, reinterpret_cast<T*>(no));
        ld.load(2 * m + 4, reinterpret_cast<T*>(n2));
      }
      float lo[4], hi[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float o1 = __fmaf_rn(__fadd_rn(xe[j], x2[j]), kA, xo[j]);
        const float e1 = __fmaf_rn(__fadd_rn(o1p[j], o1), kB, xe[j]);
        const float o2 = __fmaf_rn(__fadd_rn(e1p[j], e1), kG, o1p[j]);   // o2[m-1]
        const float e2 = __fmaf_rn(__fadd_rn(o2pp[j], o2), kD, e1p[j]);  // e2[m-1], scale 1/K pending
        lo[j] = e2; hi[j] = o2;
        xe[j] = x2[j]; o1p[j] = o1; e1p[j] = e1; o2pp[j] = o2;
      }
      if (m - 1 >= n0) {  // warp-uniform
        const bool r0 = ld.c0 + 1 == Wc - 1, r1 = ld.c0 + 3 == Wc - 1;
        hfwd97(lo, kIK, fwd97_row_variant(2 * (m - 1), Hc), ld.c0 == 0, r0, r1);
        hfwd97(hi, kK, fwd97_row_variant(2 * (m - 1) + 1, Hc), ld.c0 == 0, r0, r1);
        emit(m - 1, reinterpret_cast<T*>(lo), reinterpret_cast<T*>(hi));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { xo[j] = no[j]; x2[j] = n2[j]; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// inverse
// ---------------------------------------------------------------------------------------------

// 5/3 horizontal synthesis (DWTGenerator.cu:79-85,295-308): x = s0,d0,s1,d1 -> e0,o0,e1,o1
__device__ __forceinline__ void hinv53(int x[4]) {
  const int pd1 = __shfl_up_sync(0xffffffffu, x[3], 1);
  const int e0 = x[0] - ((pd1 + x[1] + 2) >> 2);
  const int e1 = x[2] - ((x[1] + x[3] + 2) >> 2);
  const int ne0 = __shfl_down_sync(0xffffffffu, e0, 1);
  x[1] = x[1] + ((e0 + e1) >> 1);
  x[3] = x[3] + ((e1 + ne0) >> 1);
  x[0] = e0; x[2] = e1;
}

// 9/7 horizontal synthesis (DWTGenerator.cu:110-122,326-339).  The reference's fast-math divisions by
// K and 1/K are multiplications by the rounded reciprocals here (and in the oracle).
__device__ __forceinline__ void hinv97(float x[4], float rK, float rIK) {
  float o0 = __fmul_rn(x[1], rK), o1 = __fmul_rn(x[3], rK);
  float p1 = __shfl_up_sync(0xffffffffu, o1, 1);
  float e0 = __fmaf_rn(-__fadd_rn(p1, o0), kD, __fmul_rn(x[0], rIK));
  float e1 = __fmaf_rn(-__fadd_rn(o0, o1), kD, __fmul_rn(x[2], rIK));
  float n0 = __shfl_down_sync(0xffffffffu, e0, 1);
  o0 = __fmaf_rn(-__fadd_rn(e0, e1), kG, o0);
  o1 = __fmaf_rn(-__fadd_rn(e1, n0), kG, o1);
  p1 = __shfl_up_sync(0xffffffffu, o1, 1);
  e0 = __fmaf_rn(-__fadd_rn(p1, o0), kB, e0);
  e1 = __fmaf_rn(-__fadd_rn(o0, o1), kB, e1);
  n0 = __shfl_down_sync(0xffffffffu, e0, 1);
  o0 = __fmaf_rn(-__fadd_rn(e0, e1), kA, o0);
  o1 = __fmaf_rn(-__fadd_rn(e1, n0), kA, o1);
  x[0] = e0; x[1] = o0; x[2] = e1; x[3] = o1;
}

// De-quantiser on subband read (DWTGenerator.cu:513-553): 0 -> 0, else (|q|+0.5)*sgn / Q / qs.
__device__ __forceinline__ float dequant(int q, float rq, float rqs) {
  if (q == 0) return 0.0f;
  const float mag = __fadd_rn(fabsf((float)q), 0.5f);
  return __fmul_rn(__fmul_rn(q < 0 ? -mag : mag, rq), rqs);
}

template <typename T, bool LOSSY, bool DST_U8>
__global__ void __launch_bounds__(128) dwt_inverse_level(const InvArgs a) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= a.nstrips * a.nchunks) return;
  const int frame = blockIdx.y;
  const int strip = item % a.nstrips, chunk = item / a.nstrips;
  const int Wc = a.Wc, Hc = a.Hc, half_w = Wc >> 1, half_h = Hc >> 1;
  const int c0 = strip * kStrip - 4 + 4 * lane;
  // subband columns of the lane's 4 samples (even samples: low band, odd: high band)
  int lc[2], hc[2];
  lc[0] = fold_ws(c0, Wc) >> 1; hc[0] = fold_ws(c0 + 1, Wc) >> 1;
  lc[1] = fold_ws(c0 + 2, Wc) >> 1; hc[1] = fold_ws(c0 + 3, Wc) >> 1;
  const bool vec = a.in_vec && c0 >= 0 && c0 + 3 < Wc;
  const int32_t* const q = a.q + (size_t)frame * a.q_fstride;
  const T* const lls = a.first ? nullptr : static_cast<const T*>(a.ll) + (size_t)frame * a.ll_fstride;

  // loads the horizontally interleaved row pair n: lo = (LL|HL) row, hi = (LH|HH) row, de-quantised
  auto load_pair = [&](int n, T lo[4], T hi[4]) {
    const int rl = fold_ws(2 * n, Hc) >> 1, rh = fold_ws(2 * n + 1, Hc) >> 1;
    const int32_t* row_lo = q + (size_t)rl * a.Wa;
    const int32_t* row_hi = q + (size_t)(rh + half_h) * a.Wa;
    int hl[2], lh[2], hh[2];
    if (vec) {
      int2 t = __ldg(reinterpret_cast<const int2*>(row_lo + half_w + lc[0])); hl[0] = t.x; hl[1] = t.y;
      t = __ldg(reinterpret_cast<const int2*>(row_hi + lc[0])); lh[0] = t.x; lh[1] = t.y;
      t = __ldg(reinterpret_cast<const int2*>(row_hi + half_w + lc[0])); hh[0] = t.x; hh[1] = t.y;
    } else {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        hl[k] = __ldg(row_lo + half_w + hc[k]);
        lh[k] = __ldg(row_hi + lc[k]);
        hh[k] = __ldg(row_hi + half_w + hc[k]);
      }
    }
    T ll[2];
    if (a.first) {
      int t0, t1;
      if (vec) { int2 t = __ldg(reinterpret_cast<const int2*>(row_lo + lc[0])); t0 = t.x; t1 = t.y; }
      else { t0 = __ldg(row_lo + lc[0]); t1 = __ldg(row_lo + lc[1]); }
      if (LOSSY) { ll[0] = (T)dequant(t0, a.r_ll, a.rqs); ll[1] = (T)dequant(t1, a.r_ll, a.rqs); }
      else { ll[0] = (T)t0; ll[1] = (T)t1; }
    } else {
      const T* r = lls + (size_t)rl * a.ll_pitch;
      if (vec) {
        typename Vec2<T>::type t = __ldg(reinterpret_cast<const typename Vec2<T>::type*>(r + lc[0]));
        ll[0] = t.x; ll[1] = t.y;
      } else { ll[0] = __ldg(r + lc[0]); ll[1] = __ldg(r + lc[1]); }
    }
    lo[0] = ll[0]; lo[2] = ll[1];
    if (LOSSY) {
      lo[1] = (T)dequant(hl[0], a.r_hl, a.rqs); lo[3] = (T)dequant(hl[1], a.r_hl, a.rqs);
      hi[0] = (T)dequant(lh[0], a.r_lh, a.rqs); hi[2] = (T)dequant(lh[1], a.r_lh, a.rqs);
      hi[1] = (T)dequant(hh[0], a.r_hh, a.rqs); hi[3] = (T)dequant(hh[1], a.r_hh, a.rqs);
    } else {
      lo[1] = (T)hl[0]; lo[3] = (T)hl[1];
      hi[0] = (T)lh[0]; hi[2] = (T)lh[1];
      hi[1] = (T)hh[0]; hi[3] = (T)hh[1];
    }
  };

  const bool wlane = lane >= 1 && lane <= 30;
  auto store_row = [&](int y, const T v[4]) {
    if (!wlane) return;
    if (DST_U8) {
      unsigned char px[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int r;
        if (LOSSY) r = __float2int_rn(__fadd_rn(__fadd_rn((float)v[j], (float)a.offset), 0.01f));  // DecodingEngine.cu:714
        else r = (int)v[j] + a.offset;                                                          // :728
        px[j] = (unsigned char)min(max(r, 0), 255);
      }
      unsigned char* d = static_cast<unsigned char*>(a.dst) + (size_t)frame * a.dst_fstride;
      if (a.dst_vec) {  // W == W_a: rows are contiguous, both crop modes coincide
        if (y < a.H && c0 < Wc) *reinterpret_cast<uchar4*>(d + (size_t)y * a.W + c0) = make_uchar4(px[0], px[1], px[2], px[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int x = c0 + j;
          if (x >= Wc) continue;
          if (a.flat_crop) {  // reference quirk: first W*H elements of the W_a-stride plane (IOManager.ipp:332-335)
            const long long idx = (long long)y * Wc + x;
            if (idx < (long long)a.W * a.H) d[idx] = px[j];
          } else if (x < a.W && y < a.H) {
            d[(size_t)y * a.W + x] = px[j];
          }
        }
      }
    } else {
      T* d = static_cast<T*>(a.dst) + (size_t)frame * a.dst_fstride + (size_t)y * a.dst_pitch;
      if (a.dst_vec && c0 + 3 < Wc) {
        typename Vec4<T>::type t; t.x = v[0]; t.y = v[1]; t.z = v[2]; t.w = v[3];
        *reinterpret_cast<typename Vec4<T>::type*>(d + c0) = t;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (c0 + j < Wc) d[c0 + j] = v[j];
      }
    }
  };

  const int n0 = chunk * a.rp, n1 = min(n0 + a.rp, half_h);
  if (!LOSSY) {
    // x[2n] = s[n] - ((d[n-1]+d[n]+2)>>2);  x[2n+1] = d[n] + ((x[2n]+x[2n+2])>>1)   (DWTGenerator.cu:79-85,160-181)
    int s[4], d[4], dp[4], xp[4] = {0, 0, 0, 0}, tmp[4];
    load_pair(n0 - 1, reinterpret_cast<T*>(tmp), reinterpret_cast<T*>(dp));
    hinv53(dp);
    load_pair(n0, reinterpret_cast<T*>(s), reinterpret_cast<T*>(d));
    for (int n = n0; n <= n1; ++n) {
      int ns[4], nd[4];
      if (n < n1) load_pair(n + 1, reinterpret_cast<T*>(ns), reinterpret_cast<T*>(nd));
      hinv53(s);
      hinv53(d);
      int xe[4], xo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xe[j] = s[j] - ((dp[j] + d[j] + 2) >> 2);
        xo[j] = dp[j] + ((xp[j] + xe[j]) >> 1);
        xp[j] = xe[j]; dp[j] = d[j];
      }
      if (n > n0) store_row(2 * n - 1, reinterpret_cast<const T*>(xo));
      if (n < n1) store_row(2 * n, reinterpret_cast<const T*>(xe));
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[j] = ns[j]; d[j] = nd[j]; }
    }
  } else {
    // 9/7 synthesis pipeline (DWTGenerator.cu:110-122,230-272): at step m pair m arrives; rows 2(m-1) and 2(m-2)+1 leave.
    const float rK = 1.0f / kK, rIK = 1.0f / kIK;
    float opp[4] = {0, 0, 0, 0}, e1p[4] = {0, 0, 0, 0}, o1pp[4] = {0, 0, 0, 0}, e0pp[4] = {0, 0, 0, 0};
    float lo[4], hi[4];
    load_pair(n0 - 2, reinterpret_cast<T*>(lo), reinterpret_cast<T*>(hi));
    for (int m = n0 - 2; m <= n1 + 1; ++m) {
      float nlo[4], nhi[4];
      if (m <= n1) load_pair(m + 1, reinterpret_cast<T*>(nlo), reinterpret_cast<T*>(nhi));
      hinv97(lo, rK, rIK);
      hinv97(hi, rK, rIK);
      float r_even[4], r_odd[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float op = __fmul_rn(hi[j], rK);
        const float e1 = __fmaf_rn(-__fadd_rn(opp[j], op), kD, __fmul_rn(lo[j], rIK));
        const float o1p = __fmaf_rn(-__fadd_rn(e1p[j], e1), kG, opp[j]);      // o1[m-1]
        const float e0p = __fmaf_rn(-__fadd_rn(o1pp[j], o1p), kB, e1p[j]);    // x[2(m-1)]
        const float xo = __fmaf_rn(-__fadd_rn(e0pp[j], e0p), kA, o1pp[j]);    // x[2(m-2)+1]
        r_even[j] = e0p; r_odd[j] = xo;
        opp[j] = op; e1p[j] = e1; o1pp[j] = o1p; e0pp[j] = e0p;
      }
      if (m - 2 >= n0 && m - 2 < n1) store_row(2 * (m - 2) + 1, reinterpret_cast<const T*>(r_odd));
      if (m - 1 >= n0 && m - 1 < n1) store_row(2 * (m - 1), reinterpret_cast<const T*>(r_even));
#pragma unroll
      for (int j = 0; j < 4; ++j) { lo[j] = nlo[j]; hi[j] = nhi[j]; }
    }
  }
}

// DWT/DWTGenerator.cuh:168-179 quantiser steps, [level-1][LL,HL,LH,HH]
const float kQStep[10][4] = {
    {1.965908f, 1.0112865f, 1.0112865f, 0.52021784f}, {4.1224113f, 1.9968134f, 1.9968134f, 0.96721643f},
    {8.416739f, 4.1833673f, 4.1833673f, 2.0792568f},  {16.935543f, 8.534108f, 8.534108f, 4.3004827f},
    {33.924816f, 17.166693f, 17.166693f, 8.686718f},  {67.87687f, 34.385098f, 34.385098f, 17.41882f},
    {135.76744f, 68.7964f, 68.7964f, 34.860676f},     {271.5416f, 137.60588f, 137.60588f, 69.73287f},
    {543.0866f, 275.21814f, 275.21814f, 139.47136f},  {1086.1624f, 550.43286f, 550.43286f, 278.94202f}};

inline int rows_per_warp(int half_h, int nstrips, int frames) {
  // enough warps to fill 148 SMs several times over, but chunks long enough to amortise warm-up rows
  int rp = 32;
  while (rp > 8 && (long long)nstrips * ((half_h + rp - 1) / rp) * frames < 148LL * 24) rp >>= 1;
  return rp;
}

}  // namespace

// Offsets (in elements) of LL_l, l = 1..wl-1, inside one frame's scratch area.
size_t dwt_scratch_elems(int Wa, int Ha, int wl) {
  size_t n = 0;
  for (int l = 1; l < wl; ++l) n += (size_t)(Wa >> l) * (Ha >> l);
  return n;
}
static size_t ll_offset(int Wa, int Ha, int l) {
  size_t n = 0;
  for (int k = 1; k < l; ++k) n += (size_t)(Wa >> k) * (Ha >> k);
  return n;
}

bool dwt_forward_fast_ok(const void*, const void*, const void*, bool, int, long long, int, int, int);
int dwt_forward_fast(const void*, bool, long long, int, int, int, int, void*, long long, int, void*, long long, int, int, int, bool, float,
                     bool, int, cudaStream_t, bool rgb, long long comp_stride, int offset);

bool dwt_inverse_fast_ok(const void*, const void*, const void*, bool, int, long long, int, long long, int, int);
int dwt_inverse_fast(const int32_t*, long long, int, const void*, long long, int, bool, int, int, void*, bool, long long, int, int, int, bool,
                     float, int, cudaStream_t, bool rgb, long long comp_stride, int offset);

bool dwt_forward_l12_ok(const void*, const void*, const void*, int, int, int, int, int);
int dwt_forward_l12(const uint8_t*, int, int, int, int, void*, long long, void*, long long, int, bool, float, bool, int, cudaStream_t, int offset);

// PCS_DWT_FUSE = 0: per-level kernels of this file only; 1: streaming kernels of dwt_fast.cu, one level per launch;
// 2 (default): levels 1+2 of u8 frames in one launch (dwt_fused.cu), the deeper levels through the streaming kernels
static int fuse_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("PCS_DWT_FUSE");
    mode = e ? atoi(e) : 2;
    if (mode < 0 || mode > 2) mode = 2;
  }
  return mode;
}

// Host loop over levels (replaces DWTEngine::DWTForward, DWT/DWTGenerator.cu:1268-1342; no host
// synchronisation between levels).  out_int: 9/7 only, write (int)c instead of c.
// tplane != nullptr (RGB path): level 1 reads colour-transformed 32-bit planes [frames][Ha][Wa] instead of u8 pixels.
// rgb_px (RGB path, fused): d_px holds planar u8 RGB frames, `frames` counts component planes, level 1 applies the
// colour transform in its load (streaming kernel only: the caller checks dwt_forward_rgb_ok first).
static int dwt_forward_levels(const uint8_t* d_px, void* d_coeff, void* d_scratch, int W, int H, int Wa, int Ha, int wl,
                              bool lossy, float qs, bool out_int, int frames, int lfirst, int llast, cudaStream_t st,
                              const void* tplane = nullptr, bool rgb_px = false, int offset = 128) {
  const size_t A = (size_t)Wa * Ha, S = dwt_scratch_elems(Wa, Ha, wl);
  int first = lfirst;
  // levels 1+2 of u8 frames in one launch: the LL1 plane stays on chip (dwt_fused.cu)
  if (fuse_mode() == 2 && first == 1 && llast >= 2 && !tplane && !rgb_px && S % 4 == 0) {
    void* ll2 = wl == 2 ? nullptr : static_cast<char*>(d_scratch) + ll_offset(Wa, Ha, 2) * 4;
    if (dwt_forward_l12_ok(d_px, d_coeff, ll2, W, H, Wa, Ha, wl)) {
      int rc = dwt_forward_l12(d_px, W, H, Wa, Ha, d_coeff, (long long)A, ll2, (long long)S, wl, lossy, qs, out_int, frames, st, offset);
      if (rc) return rc;
      first = 3;
    }
  }
  // (remaining) leading levels through the streaming kernel of dwt_fast.cu while the geometry allows it
  while (fuse_mode() > 0 && first <= llast) {
    const int l = first, Wc = Wa >> (l - 1), Hc = Ha >> (l - 1);
    const bool u8 = l == 1 && !tplane;
    const void* src = u8 ? static_cast<const void*>(d_px) : l == 1 ? tplane : static_cast<const char*>(d_scratch) + ll_offset(Wa, Ha, l - 1) * 4;
    const long long sfs = u8 ? (long long)W * H * (rgb_px ? 3 : 1) : l == 1 ? (long long)A : (long long)S;
    const int sp = u8 ? W : Wc;
    void* ll = (l == wl) ? nullptr : static_cast<char*>(d_scratch) + ll_offset(Wa, Ha, l) * 4;
    if (!dwt_forward_fast_ok(src, d_coeff, ll, u8, sp, sfs, H, Wc, Hc)) break;
    if (S % 4) break;
    int rc = dwt_forward_fast(src, u8, sfs, sp, H, Wc, Hc, d_coeff, (long long)A, Wa, ll, (long long)S, Wa >> l, l, wl, lossy,
                              qs, out_int, frames, st, u8 && rgb_px, (long long)W * H, offset);
    if (rc) return rc;
    first = l + 1;
  }
  for (int l = first; l <= llast; ++l) {
    FwdArgs a{};
    a.Wc = Wa >> (l - 1); a.Hc = Ha >> (l - 1);
    a.W = W; a.H = H; a.Wa = Wa;
    const bool px_in = l == 1 && !tplane;
    if (px_in) {
      a.src = d_px; a.src_fstride = (long long)W * H; a.src_pitch = W;
      a.src_vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_px) & 3) == 0) && (((long long)W * H) % 4 == 0);
    } else if (l == 1) {
      a.src = tplane; a.src_fstride = (long long)A; a.src_pitch = Wa;
      a.src_vec = (reinterpret_cast<uintptr_t>(tplane) & 15) == 0;
    } else {
      a.src = static_cast<char*>(d_scratch) + ll_offset(Wa, Ha, l - 1) * 4; a.src_fstride = (long long)S;
      a.src_pitch = a.Wc; a.src_vec = (a.Wc % 4 == 0) && (ll_offset(Wa, Ha, l - 1) % 4 == 0) && (S % 4 == 0);
    }
    a.out = d_coeff; a.out_fstride = (long long)A;
    a.last = (l == wl);
    a.ll = a.last ? d_coeff : static_cast<char*>(d_scratch) + ll_offset(Wa, Ha, l) * 4;
    a.ll_fstride = a.last ? (long long)A : (long long)S; a.ll_pitch = a.last ? Wa : (a.Wc >> 1);
    a.out_vec = (a.Wc % 4 == 0) && (S % 2 == 0) && (ll_offset(Wa, Ha, l) % 2 == 0);
    a.q_ll = kQStep[l - 1][0]; a.q_hl = kQStep[l - 1][1]; a.q_lh = kQStep[l - 1][2]; a.q_hh = kQStep[l - 1][3];
    a.qs = qs; a.offset = offset;
    a.nstrips = (a.Wc + kStrip - 1) / kStrip;
    a.rp = rows_per_warp(a.Hc >> 1, a.nstrips, frames);
    a.nchunks = ((a.Hc >> 1) + a.rp - 1) / a.rp;
    dim3 grid((a.nstrips * a.nchunks + 3) / 4, frames), block(128);
    if (!lossy) {
      if (px_in) dwt_forward_level<int, int, false, true><<<grid, block, 0, st>>>(a);
      else dwt_forward_level<int, int, false, false><<<grid, block, 0, st>>>(a);
    } else if (out_int) {
      if (px_in) dwt_forward_level<float, int, true, true><<<grid, block, 0, st>>>(a);
      else dwt_forward_level<float, int, true, false><<<grid, block, 0, st>>>(a);
    } else {
      if (px_in) dwt_forward_level<float, float, true, true><<<grid, block, 0, st>>>(a);
      else dwt_forward_level<float, float, true, false><<<grid, block, 0, st>>>(a);
    }
    ++g_launches;
  }
  return (int)cudaGetLastError();
}

// Replaces DWTEngine::DWTReverse (DWT/DWTGenerator.cu:1349-1424) + removeOffsetAndApplyMaxMin[Lossy].
// tplane != nullptr (RGB path): level 1 writes 32-bit planes [frames][Ha][Wa] for the inverse colour transform
// instead of clamped u8 pixels.
// rgb_px (RGB path, fused): `frames` counts component planes, level 1 writes planar u8 RGB frames to d_px with the
// inverse colour transform in its store (streaming kernel only: the caller checks dwt_inverse_rgb_ok first).
static int dwt_inverse_levels(const int32_t* d_q, uint8_t* d_px, void* d_scratch, int W, int H, int Wa, int Ha, int wl, bool lossy,
                              float qs, bool flat_crop, int frames, int lhigh, int llow, cudaStream_t st, void* tplane = nullptr,
                              bool rgb_px = false, int offset = 128) {
  const size_t A = (size_t)Wa * Ha, S = dwt_scratch_elems(Wa, Ha, wl);
  for (int l = lhigh; l >= llow; --l) {
    InvArgs a{};
    a.q = d_q; a.q_fstride = (long long)A; a.Wa = Wa;
    a.Wc = Wa >> (l - 1); a.Hc = Ha >> (l - 1);
    a.first = (l == wl);
    if (!a.first) {
      a.ll = static_cast<char*>(d_scratch) + ll_offset(Wa, Ha, l) * 4; a.ll_fstride = (long long)S; a.ll_pitch = a.Wc >> 1;
    }
    a.in_vec = (a.Wc % 4 == 0) && (S % 2 == 0) && (ll_offset(Wa, Ha, l) % 2 == 0);
    a.W = W; a.H = H; a.flat_crop = flat_crop ? 1 : 0;
    const bool px_out = l == 1 && !tplane;
    if (px_out) {
      a.dst = d_px; a.dst_fstride = (long long)W * H * (rgb_px ? 3 : 1); a.dst_pitch = W;
      a.dst_vec = (W == Wa) && ((reinterpret_cast<uintptr_t>(d_px) & 3) == 0) && (((long long)W * H) % 4 == 0);
    } else if (l == 1) {
      a.dst = tplane; a.dst_fstride = (long long)A; a.dst_pitch = Wa;
      a.dst_vec = (reinterpret_cast<uintptr_t>(tplane) & 15) == 0;
    } else {
      a.dst = static_cast<char*>(d_scratch) + ll_offset(Wa, Ha, l - 1) * 4; a.dst_fstride = (long long)S; a.dst_pitch = a.Wc;
      a.dst_vec = (a.Wc % 4 == 0) && (S % 4 == 0) && (ll_offset(Wa, Ha, l - 1) % 4 == 0);
    }
    a.r_ll = 1.0f / kQStep[l - 1][0]; a.r_hl = 1.0f / kQStep[l - 1][1];
    a.r_lh = 1.0f / kQStep[l - 1][2]; a.r_hh = 1.0f / kQStep[l - 1][3];
    a.rqs = 1.0f / qs; a.offset = offset;
    if (fuse_mode() > 0 && S % 4 == 0 && !(px_out && flat_crop && W != Wa) &&
        dwt_inverse_fast_ok(d_q, a.first ? nullptr : a.ll, a.dst, px_out, a.dst_pitch, a.dst_fstride, a.ll_pitch, a.ll_fstride, a.Wc, a.Hc)) {
      const bool rgb = px_out && rgb_px;
      int rc = dwt_inverse_fast(d_q, (long long)A, Wa, a.first ? nullptr : a.ll, a.ll_fstride, a.ll_pitch, a.first != 0, a.Wc, a.Hc, a.dst, px_out,
                                a.dst_fstride, a.dst_pitch, H, l, lossy, qs, rgb ? frames / 3 : frames, st, rgb, (long long)W * H, offset);
      if (rc) return rc;
      continue;
    }
    a.nstrips = (a.Wc + kStrip - 1) / kStrip;
    a.rp = rows_per_warp(a.Hc >> 1, a.nstrips, frames);
    a.nchunks = ((a.Hc >> 1) + a.rp - 1) / a.rp;
    dim3 grid((a.nstrips * a.nchunks + 3) / 4, frames), block(128);
    if (!lossy) {
      if (px_out) dwt_inverse_level<int, false, true><<<grid, block, 0, st>>>(a);
      else dwt_inverse_level<int, false, false><<<grid, block, 0, st>>>(a);
    } else {
      if (px_out) dwt_inverse_level<float, true, true><<<grid, block, 0, st>>>(a);
      else dwt_inverse_level<float, true, false><<<grid, block, 0, st>>>(a);
    }
    ++g_launches;
  }
  return (int)cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------
// Batch drivers: all frames of a call go through each launch together (one launch per level, or one launch for
// levels 1+2 where the fused kernel of dwt_fused.cu applies); nothing synchronises the host.
// ---------------------------------------------------------------------------------------------
int dwt_forward(const uint8_t* d_px, void* d_coeff, void* d_scratch, int W, int H, int Wa, int Ha, int wl, bool lossy, float qs,
                bool out_int, int frames, cudaStream_t st, int offset) {
  return dwt_forward_levels(d_px, d_coeff, d_scratch, W, H, Wa, Ha, wl, lossy, qs, out_int, frames, 1, wl, st, nullptr, false, offset);
}

int dwt_inverse(const int32_t* d_q, uint8_t* d_px, void* d_scratch, int W, int H, int Wa, int Ha, int wl, bool lossy, float qs,
                bool flat_crop, int frames, cudaStream_t st, int offset) {
  return dwt_inverse_levels(d_q, d_px, d_scratch, W, H, Wa, Ha, wl, lossy, qs, flat_crop, frames, wl, 1, st, nullptr, false, offset);
}


// RGB path, forward, colour transform fused into the level-1 load: possible when the streaming kernel takes level 1
// of this geometry (no horizontal padding, 8-byte aligned colour planes).
bool dwt_forward_rgb_ok(const uint8_t* d_px, const void* d_coeff, const void* d_scratch, int W, int H, int Wa, int Ha, int wl) {
  const size_t S = dwt_scratch_elems(Wa, Ha, wl);
  const void* ll = wl == 1 ? nullptr : static_cast<const char*>(d_scratch) + ll_offset(Wa, Ha, 1) * 4;
  return fuse_mode() > 0 && S % 4 == 0 && ((long long)W * H) % 8 == 0 &&
         dwt_forward_fast_ok(d_px, d_coeff, ll, true, W, (long long)W * H * 3, H, Wa, Ha);
}
int dwt_forward_rgb(const uint8_t* d_px, void* d_coeff, void* d_scratch, int W, int H, int Wa, int Ha, int wl, bool lossy, float qs,
                    bool out_int, int planes, cudaStream_t st) {
  return dwt_forward_levels(d_px, d_coeff, d_scratch, W, H, Wa, Ha, wl, lossy, qs, out_int, planes, 1, wl, st, nullptr, true);
}

// RGB path: the same transforms on colour-transformed 32-bit planes (`planes` = frames x 3 components).
int dwt_forward_planes(const void* d_tplanes, void* d_coeff, void* d_scratch, int Wa, int Ha, int wl, bool lossy, float qs, bool out_int,
                       int planes, cudaStream_t st) {
  return dwt_forward_levels(nullptr, d_coeff, d_scratch, Wa, Ha, Wa, Ha, wl, lossy, qs, out_int, planes, 1, wl, st, d_tplanes);
}
// RGB path, inverse, colour transform fused into the level-1 store: possible when the streaming kernel produces level 1
// of this geometry (no horizontal crop, 8-byte aligned colour planes).
bool dwt_inverse_rgb_ok(const int32_t* d_q, const uint8_t* d_px, const void* d_scratch, int W, int H, int Wa, int Ha, int wl) {
  const size_t S = dwt_scratch_elems(Wa, Ha, wl);
  const void* ll = wl == 1 ? nullptr : static_cast<const char*>(d_scratch) + ll_offset(Wa, Ha, 1) * 4;
  return fuse_mode() > 0 && S % 4 == 0 && W == Wa && ((long long)W * H) % 8 == 0 &&
         dwt_inverse_fast_ok(d_q, ll, d_px, true, W, (long long)W * H * 3, Wa >> 1, (long long)S, Wa, Ha);
}
int dwt_inverse_rgb(const int32_t* d_q, uint8_t* d_px, void* d_scratch, int W, int H, int Wa, int Ha, int wl, bool lossy, float qs,
                    int planes, cudaStream_t st) {
  return dwt_inverse_levels(d_q, d_px, d_scratch, W, H, Wa, Ha, wl, lossy, qs, false, planes, wl, 1, st, nullptr, true);
}
int dwt_inverse_planes(const int32_t* d_q, void* d_tplanes, void* d_scratch, int Wa, int Ha, int wl, bool lossy, float qs, int planes,
                       cudaStream_t st) {
  return dwt_inverse_levels(d_q, nullptr, d_scratch, Wa, Ha, Wa, Ha, wl, lossy, qs, false, planes, wl, 1, st, d_tplanes);
}

}  // namespace pcs
