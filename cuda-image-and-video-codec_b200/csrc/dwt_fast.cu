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
// f32x2 instructions.  Also in this file, both off by default and kept as measured experiments: levels
// 1+2 fused in one warp (LEVELS = 2) and all levels as one dependency-ordered persistent work queue.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <queue>
#include <type_traits>
#include <vector>

#include "pcs_common.cuh"

namespace pcs {

namespace {

constexpr float kA = -1.586134342059924f, kB = -0.052980118572961f, kG = 0.882911075530934f,
                kD = 0.443506852043971f, kK = 1.230174104914001f, kIK = 0.812893066f;
constexpr unsigned kFullMask = 0xffffffffu;

// ---- packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE-rn fp32 operations per issued instruction,
// bit-identical to the scalar __fadd_rn / __fmul_rn / __fmaf_rn with -ftz) --------------------------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 shfl_down2(f2 v) {
  float a, b; upk(v, a, b);
  return pk(__shfl_down_sync(kFullMask, a, 1), __shfl_down_sync(kFullMask, b, 1));
}
__device__ __forceinline__ f2 shfl_up2(f2 v) {
  float a, b; upk(v, a, b);
  return pk(__shfl_up_sync(kFullMask, a, 1), __shfl_up_sync(kFullMask, b, 1));
}

struct FastFwdArgs {
  const void* src; long long src_fstride; int src_pitch;    // level input: u8 pixels (level 1) or the 32-bit LL plane of the level above
  int H;                                                    // u8 input: original height (rows >= H are mirror padding); else == Hc
  int Wc, Hc;                                               // level input size (multiples of 8 / 2)
  void* out; long long out_fstride; int pitch;              // Mallat plane and its row pitch (the adapted frame width)
  void* ll; long long ll_fstride; int ll_pitch;             // LL destination of the LAST level of this launch (scratch), unless `last`
  int last;                                                 // that level is the last level: LL goes to `out`, quantised
  float q1[4], q2[4], qs;                                   // quantiser steps of the first / second level of this launch: LL, HL, LH, HH
  int nW1, ls1, nW2, ls2;                                   // row-variant tiling of the reference, per level (9/7 only)
  int nstrips, nchunks, rp;                                 // rp: row pairs per chunk (multiple of 4)
  long long comp_stride;                                    // RGB input: bytes between the R, G and B planes of a frame
};

// Which row PAIR the reference computes with the fused neighbour sum (dwt.cu fwd97_row_variant restated
// per pair: both rows of a pair always agree because tile seams are even).
__device__ __forceinline__ bool pair_variant(int n, int nW, int ls) {
  if (nW <= 1) return n >= 6;
  if (2 * n >= ls + 4) return 2 * n - ls >= 12;
  if (n < 7) return n >= 6;
  return (n - 2) % 5 == 4;
}

template <typename OT> __device__ __forceinline__ OT cvt(float v);
template <> __device__ __forceinline__ float cvt<float>(float v) { return v; }
template <> __device__ __forceinline__ int cvt<int>(float v) { return __float2int_rz(v); }  // BPCEngine.cu:49
// c * Q * qs, two rounded multiplies left to right (DWTGenerator.cu:403-419); QS1: qs == 1.0f, x * 1.0f == x exactly
template <typename OT, bool QS1> __device__ __forceinline__ OT quant(float v, float q, float qs) {
  const float t = __fmul_rn(v, q);
  return cvt<OT>(QS1 ? t : __fmul_rn(t, qs));
}

// ---- horizontal analysis, N samples per lane (N = 8: level 1, N = 4: level 2) ----------------------
// 5/3 (DWTGenerator.cu:72-76,279-292).  left: the lane holds image column 0 (d[-1] := d[0]);
// right: the lane's last pair is the image's last pair (x[Wc] := x[Wc-2]).
template <int N>
__device__ __forceinline__ void hfwd53(int (&x)[N], bool left, bool right) {
  constexpr int P = N / 2;
  int nx = __shfl_down_sync(kFullMask, x[0], 1);
  if (right) nx = x[N - 2];
  int d[P];
#pragma unroll
  for (int k = 0; k < P; ++k) d[k] = x[2 * k + 1] - ((x[2 * k] + (k + 1 < P ? x[2 * k + 2] : nx)) >> 1);
  int pd = __shfl_up_sync(kFullMask, d[P - 1], 1);
  if (left) pd = d[0];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    x[2 * k] = x[2 * k] + (((k ? d[k - 1] : pd) + d[k] + 2) >> 2);
    x[2 * k + 1] = d[k];
  }
}

// 9/7 (DWTGenerator.cu:91-104,311-323) of a row whose vertical-pass scale S is still pending; see
// dwt.cu hfwd97 for the operation shapes (they are the reference binary's, FMA contractions included).
template <int N, bool FUSED_SUM>
__device__ __forceinline__ void hfwd97(float (&x)[N], float S, bool left, bool right) {
  constexpr int P = N / 2;
  float e[P], o[P];
#pragma unroll
  for (int k = 0; k < P; ++k) e[k] = __fmul_rn(x[2 * k], S);
  float n0 = __shfl_down_sync(kFullMask, e[0], 1);
  if (right) n0 = e[P - 1];
  if constexpr (FUSED_SUM) {
#pragma unroll
    for (int k = 0; k < P; ++k) o[k] = __fmaf_rn(x[2 * k], S, k + 1 < P ? e[k + 1] : n0);
  } else {
#pragma unroll
    for (int k = 0; k < P; ++k) o[k] = __fadd_rn(e[k], k + 1 < P ? e[k + 1] : n0);
  }
#pragma unroll
  for (int k = 0; k < P; ++k) o[k] = __fmaf_rn(x[2 * k + 1], S, __fmul_rn(o[k], kA));
  float p1 = __shfl_up_sync(kFullMask, o[P - 1], 1);
  if (left) p1 = o[0];
#pragma unroll
  for (int k = P - 1; k >= 0; --k) e[k] = __fmaf_rn(__fadd_rn(k ? o[k - 1] : p1, o[k]), kB, e[k]);
  n0 = __shfl_down_sync(kFullMask, e[0], 1);
  if (right) n0 = e[P - 1];
#pragma unroll
  for (int k = 0; k < P; ++k) o[k] = __fmaf_rn(__fadd_rn(e[k], k + 1 < P ? e[k + 1] : n0), kG, o[k]);
  p1 = __shfl_up_sync(kFullMask, o[P - 1], 1);
  if (left) p1 = o[0];
#pragma unroll
  for (int k = P - 1; k >= 0; --k) e[k] = __fmul_rn(__fmaf_rn(__fadd_rn(k ? o[k - 1] : p1, o[k]), kD, e[k]), kIK);
#pragma unroll
  for (int k = 0; k < P; ++k) { x[2 * k] = e[k]; x[2 * k + 1] = __fmul_rn(o[k], kK); }
}

// The same on BOTH rows of a row pair at once: sample j of the low row (pending scale 1/K) and of the high row
// (pending scale K) travel as one packed pair, so every lifting operation is issued once for the two rows.
template <int N, bool FUSED_SUM>
__device__ __forceinline__ void hfwd97_rows(float (&lo)[N], float (&hi)[N], bool left, bool right) {
  constexpr int P = N / 2;
  const f2 S = pk(kIK, kK), A2 = pk(kA, kA), B2 = pk(kB, kB), G2 = pk(kG, kG), D2 = pk(kD, kD), IK2 = pk(kIK, kIK), K2 = pk(kK, kK);
  f2 x[N], e[P], o[P];
#pragma unroll
  for (int j = 0; j < N; ++j) x[j] = pk(lo[j], hi[j]);
#pragma unroll
  for (int k = 0; k < P; ++k) e[k] = mul2(x[2 * k], S);
  f2 n0 = shfl_down2(e[0]);
  if (right) n0 = e[P - 1];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const f2 nxt = k + 1 < P ? e[k + 1] : n0;
    o[k] = FUSED_SUM ? fma2(x[2 * k], S, nxt) : add2(e[k], nxt);
  }
#pragma unroll
  for (int k = 0; k < P; ++k) o[k] = fma2(x[2 * k + 1], S, mul2(o[k], A2));
  f2 p1 = shfl_up2(o[P - 1]);
  if (left) p1 = o[0];
#pragma unroll
  for (int k = P - 1; k >= 0; --k) e[k] = fma2(add2(k ? o[k - 1] : p1, o[k]), B2, e[k]);
  n0 = shfl_down2(e[0]);
  if (right) n0 = e[P - 1];
#pragma unroll
  for (int k = 0; k < P; ++k) o[k] = fma2(add2(e[k], k + 1 < P ? e[k + 1] : n0), G2, o[k]);
  p1 = shfl_up2(o[P - 1]);
  if (left) p1 = o[0];
#pragma unroll
  for (int k = P - 1; k >= 0; --k) e[k] = mul2(fma2(add2(k ? o[k - 1] : p1, o[k]), D2, e[k]), IK2);
#pragma unroll
  for (int k = 0; k < P; ++k) {
    upk(e[k], lo[2 * k], hi[2 * k]);
    upk(mul2(o[k], K2), lo[2 * k + 1], hi[2 * k + 1]);
  }
}

// ---- vertical analysis pipelines (state per column) ------------------------------------------------
// 5/3: step m consumes rows 2m+1, 2m+2 and produces row pair m.  d[n] = x[2n+1] - ((x[2n]+x[2n+2])>>1),
// s[n] = x[2n] + ((d[n-1]+d[n]+2)>>2)  (DWTGenerator.cu:72-76,137-157).
template <int N>
struct V53 {
  int xe[N], dp[N];
  static constexpr int kDelay = 0, kWarm = 1;
  __device__ __forceinline__ void init(const int (&r)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) { xe[j] = r[j]; dp[j] = 0; }
  }
  // top: first pair of the image (d[-1] := d[0]); the bottom rule (x[N] := x[N-2]) is applied by the caller to x2
  __device__ __forceinline__ void step(const int (&xo)[N], const int (&x2)[N], bool top, int (&lo)[N], int (&hi)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const int d = xo[j] - ((xe[j] + x2[j]) >> 1);
      const int p = top ? d : dp[j];
      lo[j] = xe[j] + ((p + d + 2) >> 2);
      hi[j] = d; dp[j] = d; xe[j] = x2[j];
    }
  }
};
// 9/7: step m consumes rows 2m+1, 2m+2 and produces row pair m-1 (scales 1/K, K still pending)
// (DWTGenerator.cu:91-104,184-227).
template <int N>
struct V97 {
  float xe[N], o1p[N], e1p[N], o2pp[N];
  static constexpr int kDelay = 1, kWarm = 2;
  __device__ __forceinline__ void init(const float (&r)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) { xe[j] = r[j]; o1p[j] = 0.f; e1p[j] = 0.f; o2pp[j] = 0.f; }
  }
  // m0: this is step 0 of the image (o1[-1] := o1[0]); m1: step 1 (o2[-1] := o2[0]);
  // flush: one step past the last pair (e1[n/2] := e1[n/2-1]); xo/x2 are ignored then
  __device__ __forceinline__ void step(const float (&xo)[N], const float (&x2)[N], bool m0, bool m1, bool flush,
                                       float (&lo)[N], float (&hi)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float o1 = __fmaf_rn(__fadd_rn(xe[j], x2[j]), kA, xo[j]);
      const float op = m0 ? o1 : o1p[j];
      float e1 = __fmaf_rn(__fadd_rn(op, o1), kB, xe[j]);
      if (flush) e1 = e1p[j];
      const float o2 = __fmaf_rn(__fadd_rn(e1p[j], e1), kG, op);       // o2[m-1]
      const float opp = m1 ? o2 : o2pp[j];
      lo[j] = __fmaf_rn(__fadd_rn(opp, o2), kD, e1p[j]);               // e2[m-1]
      hi[j] = o2;
      xe[j] = x2[j]; o1p[j] = o1; e1p[j] = e1; o2pp[j] = o2;
    }
  }
};

// V97::step without border rules, two columns per instruction
template <int N>
__device__ __forceinline__ void v97_step_packed(V97<N>& v, const float (&xo)[N], const float (&x2)[N], float (&lo)[N], float (&hi)[N]) {
  const f2 A2 = pk(kA, kA), B2 = pk(kB, kB), G2 = pk(kG, kG), D2 = pk(kD, kD);
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    const f2 xe = pk(v.xe[j], v.xe[j + 1]), xn = pk(x2[j], x2[j + 1]), o1p = pk(v.o1p[j], v.o1p[j + 1]);
    const f2 e1p = pk(v.e1p[j], v.e1p[j + 1]), o2pp = pk(v.o2pp[j], v.o2pp[j + 1]);
    const f2 o1 = fma2(add2(xe, xn), A2, pk(xo[j], xo[j + 1]));
    const f2 e1 = fma2(add2(o1p, o1), B2, xe);
    const f2 o2 = fma2(add2(e1p, e1), G2, o1p);   // o2[m-1]
    const f2 e2 = fma2(add2(o2pp, o2), D2, e1p);  // e2[m-1]
    upk(e2, lo[j], lo[j + 1]); upk(o2, hi[j], hi[j + 1]);
    v.xe[j] = x2[j]; v.xe[j + 1] = x2[j + 1];
    upk(o1, v.o1p[j], v.o1p[j + 1]); upk(e1, v.e1p[j], v.e1p[j + 1]); upk(o2, v.o2pp[j], v.o2pp[j + 1]);
  }
}

template <bool LOSSY, int N> struct VSel { using type = V53<N>; using T = int; };
template <int N> struct VSel<true, N> { using type = V97<N>; using T = float; };

// 8 pixels of a u8 row -> T (px - 128; for 5/3 the shift is applied to the LL output instead, which is
// exact because every high-pass step is shift-invariant and the low-pass steps carry the offset through)
__device__ __forceinline__ void unpack8(uint2 p, float (&v)[8]) {
  const f2 C2 = pk(-8388736.0f, -8388736.0f);  // 2^23 + b - (2^23 + 128), exact
#pragma unroll
  for (int j = 0; j < 4; j += 2) {
    upk(add2(pk(__uint_as_float(__byte_perm(p.x, 0x4B000000u, 0x7540u + j)), __uint_as_float(__byte_perm(p.x, 0x4B000000u, 0x7541u + j))), C2),
        v[j], v[j + 1]);
    upk(add2(pk(__uint_as_float(__byte_perm(p.y, 0x4B000000u, 0x7540u + j)), __uint_as_float(__byte_perm(p.y, 0x4B000000u, 0x7541u + j))), C2),
        v[4 + j], v[5 + j]);
  }
}
__device__ __forceinline__ void unpack8(uint2 p, int (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] = (int)__byte_perm(p.x, 0u, 0x4440u + j) - 128;
    v[4 + j] = (int)__byte_perm(p.y, 0u, 0x4440u + j) - 128;
  }
}

// ---- cp.async staging: every lane copies its own bytes of future rows into a per-warp shared-memory ring and
// reads them back itself, so no barrier is needed; depth D steps hides the DRAM latency at low occupancy ----
__device__ __forceinline__ void cp_async8(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T> struct Row8 { typename std::conditional<std::is_same<T, float>::value, float4, int4>::type a, b; };
__device__ __forceinline__ void unpack8(const Row8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void unpack8(const Row8<int>& r, int (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}

template <typename OT> struct V4;
template <> struct V4<int> { using type = int4; };
template <> struct V4<float> { using type = float4; };
template <typename OT> struct V2;
template <> struct V2<int> { using type = int2; };
template <> struct V2<float> { using type = float2; };

// Subband / pixel stores are written once and not read again by this transform: streaming (evict-first).
template <typename OT>
__device__ __forceinline__ void st4(OT* p, OT a, OT b, OT c, OT d) {
  typename V4<OT>::type t; t.x = a; t.y = b; t.z = c; t.w = d;
  __stcs(reinterpret_cast<typename V4<OT>::type*>(p), t);
}
// LL planes are read back by the next launch: ask L2 to keep them (evict-last) so that the read is an L2
// hit instead of a second trip to HBM (126 MB of L2; the host loop sizes sub-batches accordingly).
__device__ __forceinline__ unsigned long long l2_keep_policy() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned bits(float v) { return __float_as_uint(v); }
__device__ __forceinline__ unsigned bits(int v) { return (unsigned)v; }
template <typename T>
__device__ __forceinline__ void st4_keep(T* p, T a, T b, T c, T d) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(bits(a)), "r"(bits(b)), "r"(bits(c)), "r"(bits(d)),
               "l"(l2_keep_policy())
               : "memory");
}
template <typename T>
__device__ __forceinline__ void st2_keep(T* p, T a, T b) {
  asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(p), "r"(bits(a)), "r"(bits(b)), "l"(l2_keep_policy()) : "memory");
}
template <typename OT>
__device__ __forceinline__ void st2(OT* p, OT a, OT b) {
  typename V2<OT>::type t; t.x = a; t.y = b;
  __stcs(reinterpret_cast<typename V2<OT>::type*>(p), t);
}

// Stores the three (or four) subband rows of one row pair.  NS = samples per lane per row (8 or 4).
template <bool LOSSY, bool QS1, typename OT, typename T, int NS>
__device__ __forceinline__ void emit_pair(const T (&lo)[NS], const T (&hi)[NS], OT* row_lo, OT* row_hi, int hw, const float (&q)[4], float qs,
                                          bool ll_to_out, T* ll_dst) {
  constexpr int P = NS / 2;
  OT v[4][P];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    if constexpr (LOSSY) {
      // c * Q * qs, two rounded multiplies left to right (DWTGenerator.cu:403-419), low and high row packed
      f2 e = mul2(pk(lo[2 * k], hi[2 * k]), pk(q[0], q[2])), o = mul2(pk(lo[2 * k + 1], hi[2 * k + 1]), pk(q[1], q[3]));
      if (!QS1) { e = mul2(e, pk(qs, qs)); o = mul2(o, pk(qs, qs)); }
      float t0, t1, t2, t3;
      upk(e, t0, t2); upk(o, t1, t3);
      v[0][k] = cvt<OT>(t0); v[1][k] = cvt<OT>(t1); v[2][k] = cvt<OT>(t2); v[3][k] = cvt<OT>(t3);
    } else {
      v[0][k] = lo[2 * k]; v[1][k] = lo[2 * k + 1]; v[2][k] = hi[2 * k]; v[3][k] = hi[2 * k + 1];
    }
  }
  if constexpr (P == 4) {
    st4<OT>(row_lo + hw, v[1][0], v[1][1], v[1][2], v[1][3]);
    st4<OT>(row_hi, v[2][0], v[2][1], v[2][2], v[2][3]);
    st4<OT>(row_hi + hw, v[3][0], v[3][1], v[3][2], v[3][3]);
    if (ll_to_out) st4<OT>(row_lo, v[0][0], v[0][1], v[0][2], v[0][3]);
    else if (ll_dst) st4_keep<T>(ll_dst, lo[0], lo[2], lo[4], lo[6]);
  } else {
    st2<OT>(row_lo + hw, v[1][0], v[1][1]);
    st2<OT>(row_hi, v[2][0], v[2][1]);
    st2<OT>(row_hi + hw, v[3][0], v[3][1]);
    if (ll_to_out) st2<OT>(row_lo, v[0][0], v[0][1]);
    else if (ll_dst) st2_keep<T>(ll_dst, lo[0], lo[2]);
  }
}

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

// One work item = one 256-column strip x one chunk of row pairs of one frame, one level (LEVELS = 1; LL -> scratch or,
// if last, -> out) or two fused levels (LEVELS = 2, u8 input only).  SRC_U8: the input is the u8 frame (level 1), else
// the 32-bit LL plane of the level above.  COH: that plane was written by other SMs during this very launch (queue
// kernel below), so it is read through L2 (ld.global.cg) instead of the non-coherent path.
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
template <bool LOSSY, typename OT, bool SRC_U8, int LEVELS, bool QS1, bool COH, bool RGB = false>
__device__ __forceinline__ void fwd_item(const FastFwdArgs& a, const int strip, const int chunk, const int frame, FwdRing* const ringp) {
  static_assert(!RGB || (SRC_U8 && LEVELS == 1), "the fused colour transform reads u8 pixels and feeds one level");
  using T = typename VSel<LOSSY, 8>::T;
  using VP1 = typename VSel<LOSSY, 8>::type;
  using VP2 = typename VSel<LOSSY, 4>::type;
  using RowT = typename std::conditional<SRC_U8, uint2, Row8<T>>::type;
  constexpr bool kPf2 = kPf2On && LOSSY;
  const int lane = threadIdx.x & 31;
  constexpr int kHalo = LEVELS;                        // halo lanes per side
  constexpr int kEff = (32 - 2 * kHalo) * 8;           // effective columns per strip
  const int Wc = a.Wc, Hc = a.Hc, hw1 = Wc >> 1, hh1 = Hc >> 1, hw2 = Wc >> 2, hh2 = Hc >> 2;
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

  // first-level row pairs [n0, n1) and second-level row pairs [m2_0, m2_1) are written by this chunk
  const int n0 = chunk * a.rp, n1 = min(n0 + a.rp, hh1);
  const int m2_0 = n0 >> 1, m2_1 = n1 >> 1;
  const int M0 = max(m2_0 - VP2::kWarm, 0);  // first second-level step; it starts from LL row 2*M0
  // first / last first-level pair whose output is needed (by that level itself and by the second one)
  const int P0 = LEVELS == 2 ? min(n0, 2 * M0) : n0;
  const int Plast = LEVELS == 2 ? min(max(n1 - 1, 2 * (m2_1 - 1 + VP2::kDelay) + 2), hh1 - 1) : n1 - 1;

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
      unpack8(*reinterpret_cast<const uint2*>(&ring[slot][r][lane][0]), v);
    }
  };
  auto load_row = [&](int y) -> RowT {
    if constexpr (SRC_U8) {
      return __ldg(reinterpret_cast<const uint2*>(src8 + (size_t)row_index(y) * a.src_pitch));
    } else {
      RowT r;
      const auto* p4 = reinterpret_cast<const decltype(r.a)*>(srcT + (size_t)row_index(y) * a.src_pitch);
      if constexpr (COH) { r.a = __ldcg(p4); r.b = __ldcg(p4 + 1); }
      else { r.a = __ldg(p4); r.b = __ldg(p4 + 1); }
      return r;
    }
  };

  VP1 v1;
  VP2 v2;
  T ll_odd[4] = {};  // buffered odd LL row for the second level
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
      unpack8(load_row(2 * ms), r);
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
                                 LEVELS == 1 && a.last, (LEVELS == 1 && !a.last) ? llbase + (size_t)pr * a.ll_pitch + oc1 : nullptr);
    }
    if constexpr (LEVELS == 2) {
      // ---- second level in the same warp: LL row R arrives, 4 samples per lane (LL columns c0/2 .. c0/2+3) ----
      const int R = pr;
      if (R < 2 * M0) continue;
      T row[4] = {lo[0], lo[2], lo[4], lo[6]};
      if (R == 2 * M0) { v2.init(row); continue; }
      const bool lastrow = R == hh1 - 1;  // bottom of the image: LL row hh1 does not exist
      if (R & 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ll_odd[j] = row[j];
        if (!lastrow) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) row[j] = v2.xe[j];  // x[n] := x[n-2]
      }
      // normally one step (M = R/2 - 1); at the bottom row the last step and, for 9/7, the flush step
      const int Mfirst = lastrow ? (R - 1) >> 1 : (R >> 1) - 1;
      const int Mlast = lastrow ? Mfirst + VP2::kDelay : Mfirst;
      for (int M = Mfirst; M <= Mlast; ++M) {
        T lo2[4], hi2[4];
        if constexpr (LOSSY) v2.step(ll_odd, row, M == 0, M == 1, M == hh2, lo2, hi2);
        else v2.step(ll_odd, row, M == 0, lo2, hi2);
        const int p2 = M - VP2::kDelay;
        if (p2 < m2_0 || p2 >= m2_1) continue;
        if constexpr (LOSSY) {
          // the same lanes hold the second level's borders
          if (pair_variant(p2, a.nW2, a.ls2)) { hfwd97<4, true>(lo2, kIK, left, right); hfwd97<4, true>(hi2, kK, left, right); }
          else { hfwd97<4, false>(lo2, kIK, left, right); hfwd97<4, false>(hi2, kK, left, right); }
        } else {
          hfwd53<4>(lo2, left, right);
          hfwd53<4>(hi2, left, right);
        }
        if (!wlane) continue;
        const int oc2 = c0 >> 2;
        emit_pair<LOSSY, QS1, OT, T, 4>(lo2, hi2, out + (size_t)p2 * a.pitch + oc2, out + (size_t)(p2 + hh2) * a.pitch + oc2, hw2, a.q2, qs,
                                   a.last != 0, a.last ? nullptr : llbase + (size_t)p2 * a.ll_pitch + oc2);
      }
    }
  }
}

#ifndef PCS_FWD_MINB_U8
#define PCS_FWD_MINB_U8 4
#endif
template <bool LOSSY, typename OT, bool SRC_U8, int LEVELS, bool QS1, bool RGB = false>
__global__ void __launch_bounds__(128, LEVELS == 2 ? 3 : SRC_U8 ? PCS_FWD_MINB_U8 : PCS_FWD_MINB) dwt_forward_fused(const FastFwdArgs a) {
  __shared__ __align__(16) FwdRing s_ring[SRC_U8 ? 4 : 1][RGB ? 3 : 1];
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= a.nstrips * a.nchunks) return;
  fwd_item<LOSSY, OT, SRC_U8, LEVELS, QS1, false, RGB>(a, item % a.nstrips, item / a.nstrips, blockIdx.y,
                                                       &s_ring[SRC_U8 ? (threadIdx.x >> 5) : 0][0]);
}

// quantiser steps, [level-1][LL,HL,LH,HH] (DWT/DWTGenerator.cuh:168-179)
const float kQ[10][4] = {
    {1.965908f, 1.0112865f, 1.0112865f, 0.52021784f}, {4.1224113f, 1.9968134f, 1.9968134f, 0.96721643f},
    {8.416739f, 4.1833673f, 4.1833673f, 2.0792568f},  {16.935543f, 8.534108f, 8.534108f, 4.3004827f},
    {33.924816f, 17.166693f, 17.166693f, 8.686718f},  {67.87687f, 34.385098f, 34.385098f, 17.41882f},
    {135.76744f, 68.7964f, 68.7964f, 34.860676f},     {271.5416f, 137.60588f, 137.60588f, 69.73287f},
    {543.0866f, 275.21814f, 275.21814f, 139.47136f},  {1086.1624f, 550.43286f, 550.43286f, 278.94202f}};

void variant_params(int Hc, int* nW, int* ls) {
  *nW = (Hc - 8 + 9) / 10;
  *ls = (*nW - 1) * 10 - ((10 - ((Hc - 8) % 10)) % 10);
}

// Row pairs per chunk.  A chunk costs (rp + warm-up) serial steps and the grid runs in ceil(CTAs / resident CTAs)
// waves, so pick the rp that minimises waves * steps: long chunks amortise the warm-up rows, but a grid that
// spills a few CTAs into a second wave doubles the time of the small, latency-bound levels.
int choose_rp(int nstrips, int hh, int frames, int warm, int step, long long slots = 148LL * 4) {
  // slots: CTAs of 128 threads the GPU holds at once (148 SMs x 4 at <= 128 registers per thread)
  int best = step; long long best_cost = -1;
  for (int rp = step; rp <= 128; rp += step) {
    const long long ctas = (long long)((nstrips * ((hh + rp - 1) / rp) + 3) / 4) * frames;
    const long long cost = ((ctas + slots - 1) / slots) * (rp + warm);
    if (best_cost < 0 || cost <= best_cost) { best = rp; best_cost = cost; }
  }
  return best;
}

// CTAs of kernel K the whole GPU holds at once (its register count decides how many per SM)
template <typename K>
long long resident_ctas(K kernel) {
  int dev = 0, sms = 148, per_sm = 3;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0);
  return (long long)sms * std::max(per_sm, 1);
}

// The chunk length is chosen per kernel variant, for the number of CTAs that variant really keeps resident.
template <bool LOSSY, typename OT, bool SRC_U8, int LEVELS, bool QS1, bool RGB = false>
void launch_k(FastFwdArgs a, int frames, cudaStream_t st) {
  static const long long slots = resident_ctas(dwt_forward_fused<LOSSY, OT, SRC_U8, LEVELS, QS1, RGB>);
  const int hh1 = a.Hc >> 1;
  a.rp = choose_rp(a.nstrips, hh1, frames, LEVELS == 2 ? 12 : 3, 4, slots);
  a.nchunks = (hh1 + a.rp - 1) / a.rp;
  dim3 grid((a.nstrips * a.nchunks + 3) / 4, frames);
  dwt_forward_fused<LOSSY, OT, SRC_U8, LEVELS, QS1, RGB><<<grid, 128, 0, st>>>(a);
}
template <bool LOSSY, typename OT, bool SRC_U8, int LEVELS, bool RGB = false>
void launch(const FastFwdArgs& a, int frames, cudaStream_t st) {
  if (a.qs == 1.0f) launch_k<LOSSY, OT, SRC_U8, LEVELS, true, RGB>(a, frames, st);
  else launch_k<LOSSY, OT, SRC_U8, LEVELS, false, RGB>(a, frames, st);
}


// ---- all levels in ONE persistent launch: a dependency-ordered work queue ---------------------------------------
// The per-level launches leave three things on the table: the gaps and partially filled last waves of wl launches,
// the latency-bound deep levels running alone on an empty GPU, and a trip to HBM and back for every LL plane.  Here
// the items of ALL levels of ALL frames sit in one list; warps take tickets in list order.  An item of level l >= 2
// lists the chunks of level l-1 (same frame) whose LL rows it reads and waits until their per-chunk counters say
// all strips are finished.  The host orders the list so that every item comes after the items it depends on
// (ticket order == a topological order: a waiting warp only ever waits for warps that already run, so the queue
// cannot deadlock whatever the residency) and at least `lag` tickets after the last of them (about a wave: by the
// time a warp picks the item up, its producers have normally finished and their LL rows are still in L2).
constexpr int kMaxQLevels = 7;  // the codestream header cannot carry more (Annex C.1)
struct FwdQueueArgs {
  FastFwdArgs lv[kMaxQLevels];
  const int4* items;   // {level | frame << 8, strip, chunk, dep_lo | dep_hi << 16}
  unsigned* ctr;       // [0]: ticket; [1 + (frame * nlevels + level) * maxchunks + chunk]: finished strips
  int nitems, nlevels, maxchunks;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

#ifndef PCS_FWDQ_MINB
#define PCS_FWDQ_MINB 4
#endif
template <bool LOSSY, typename OT, bool SRC_U8, bool QS1>
__global__ void __launch_bounds__(128, PCS_FWDQ_MINB) dwt_forward_queue_kernel(const __grid_constant__ FwdQueueArgs q) {
  __shared__ __align__(16) FwdRing s_ring[SRC_U8 ? 4 : 1];
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(q.ctr, 1u);
    t = __shfl_sync(kFullMask, t, 0);
    if (t >= (unsigned)q.nitems) break;
    const int4 it = __ldg(q.items + t);
    const int level = it.x & 0xff, frame = it.x >> 8;
    unsigned* const cnt = q.ctr + 1 + ((size_t)frame * q.nlevels + level) * q.maxchunks;
    if (level == 0) {
      fwd_item<LOSSY, OT, SRC_U8, 1, QS1, false>(q.lv[0], it.y, it.z, frame, &s_ring[SRC_U8 ? (threadIdx.x >> 5) : 0]);
      if constexpr (SRC_U8) cp_async_wait<0>();  // the ring is reused by this warp's next item
    } else {
      if (lane == 0) {
        const unsigned* const dep = cnt - q.maxchunks;
        const unsigned need = (unsigned)q.lv[level - 1].nstrips;
        // (the spin bound only turns a host-side ordering bug into wrong output instead of a hung GPU)
        unsigned spins = 0;
        for (int c = it.w & 0xffff; c <= (it.w >> 16); ++c)
          while (ld_acquire(dep + c) < need && ++spins < (1u << 23)) __nanosleep(64);
      }
      __syncwarp();
      fwd_item<LOSSY, OT, false, 1, QS1, true>(q.lv[level], it.y, it.z, frame, nullptr);
    }
    if (level + 1 < q.nlevels) {  // publish: every lane's LL stores, then the count
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(cnt + it.z, 1u);
    }
  }
}

struct QueueGeo { int nstrips, nchunks, rp, hh; };

// List scheduling with a lag (see above).  Level-1 items in frame / chunk / strip order form the backbone; an item
// of a deeper level becomes ready when the last chunk it reads has been placed and is placed `lag` tickets later.
std::vector<int4> build_queue(const QueueGeo* g, int nlevels, int frames, bool lossy, int lag) {
  const int warm = lossy ? 2 : 1, delay = lossy ? 1 : 0;
  struct Pending { long long at; int level, frame, chunk; };
  auto later = [](const Pending& x, const Pending& y) {
    if (x.at != y.at) return x.at > y.at;
    if (x.frame != y.frame) return x.frame > y.frame;
    if (x.level != y.level) return x.level > y.level;
    return x.chunk > y.chunk;
  };
  std::priority_queue<Pending, std::vector<Pending>, decltype(later)> pend(later);
  // placed[(frame, level, chunk)] = ticket of the chunk's LAST strip, -1 while not placed
  int maxc = 0;
  for (int l = 0; l < nlevels; ++l) maxc = std::max(maxc, g[l].nchunks);
  std::vector<long long> placed((size_t)frames * nlevels * maxc, -1);
  auto deps = [&](int l, int c, int* lo, int* hi) {  // chunks of level l-1 read by chunk c of level l
    const int n0 = c * g[l].rp, n1 = std::min(n0 + g[l].rp, g[l].hh);
    const int Hc = 2 * g[l].hh;  // rows of the level's input == LL rows of the level above
    const int r0 = std::max(0, 2 * (n0 - warm)), r1 = std::min(Hc - 1, 2 * (n1 - 1 + delay) + 2);
    *lo = r0 / g[l - 1].rp; *hi = std::min(r1 / g[l - 1].rp, g[l - 1].nchunks - 1);
  };
  std::vector<int4> out;
  auto place = [&](int l, int f, int c) {
    int lo = 0, hi = 0;
    if (l) deps(l, c, &lo, &hi);
    for (int s = 0; s < g[l].nstrips; ++s) out.push_back(make_int4(l | (f << 8), s, c, lo | (hi << 16)));
    placed[((size_t)f * nlevels + l) * maxc + c] = (long long)out.size() - 1;
    if (l + 1 < nlevels) {  // which chunks of the next level did this complete?
      for (int c2 = 0; c2 < g[l + 1].nchunks; ++c2) {
        int lo2, hi2; deps(l + 1, c2, &lo2, &hi2);
        if (c < lo2 || c > hi2) continue;
        long long last = -1; bool all = true;
        for (int d = lo2; d <= hi2; ++d) {
          const long long t = placed[((size_t)f * nlevels + l) * maxc + d];
          if (t < 0) { all = false; break; }
          last = std::max(last, t);
        }
        if (all && last == (long long)out.size() - 1) pend.push({last + lag, l + 1, f, c2});
      }
    }
  };
  auto flush = [&](bool all) {
    while (!pend.empty() && (all || pend.top().at <= (long long)out.size())) {
      const Pending p = pend.top(); pend.pop();
      place(p.level, p.frame, p.chunk);
    }
  };
  for (int f = 0; f < frames; ++f)
    for (int c = 0; c < g[0].nchunks; ++c) { flush(false); place(0, f, c); }
  flush(true);
  return out;
}

struct QueueTable { int4* d_items = nullptr; int nitems = 0, maxchunks = 0; QueueGeo geo[kMaxQLevels]; };
struct QueueKey {
  int Wa, Ha, nlevels, lossy, frames, rp1, lag, dev;
  bool operator<(const QueueKey& o) const { return std::memcmp(this, &o, sizeof(*this)) < 0; }
};
std::map<QueueKey, QueueTable> g_queue_tables;
std::mutex g_queue_mutex;

int env_int(const char* name, int dflt) { const char* e = std::getenv(name); return e ? std::atoi(e) : dflt; }

int queue_resident_warps() {
  static int n = [] {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * PCS_FWDQ_MINB * 4;
  }();
  return n;
}

int queue_geo(int Wa, int Ha, int nlevels, int rp1, QueueGeo* geo) {
  int maxchunks = 0;
  for (int l = 0; l < nlevels; ++l) {
    const int Wc = Wa >> l, hh = Ha >> (l + 1);
    geo[l].nstrips = (Wc + 239) / 240;
    geo[l].hh = hh;
    geo[l].rp = std::min(std::max(8, (rp1 >> l) & ~3), (hh + 3) & ~3);
    geo[l].nchunks = (hh + geo[l].rp - 1) / geo[l].rp;
    maxchunks = std::max(maxchunks, geo[l].nchunks);
  }
  return maxchunks;
}
int queue_rp1() {
  static const int rp1 = std::max(8, env_int("PCS_DWTQ_RP", 64) & ~3);
  return rp1;
}

// Levels 1..nlevels of this geometry as one queue (items table cached per geometry and batch size).
const QueueTable* queue_table(int Wa, int Ha, int nlevels, bool lossy, int frames) {
  const int rp1 = queue_rp1();
  static const int lag_pct = env_int("PCS_DWTQ_LAG", 125);  // percent of the resident warps
  const int lag = (int)((long long)queue_resident_warps() * lag_pct / 100);
  int dev = 0;
  cudaGetDevice(&dev);
  QueueKey key;
  std::memset(&key, 0, sizeof key);
  key.Wa = Wa; key.Ha = Ha; key.nlevels = nlevels; key.lossy = lossy; key.frames = frames; key.rp1 = rp1; key.lag = lag; key.dev = dev;
  std::lock_guard<std::mutex> lk(g_queue_mutex);
  auto itf = g_queue_tables.find(key);
  if (itf != g_queue_tables.end()) return &itf->second;
  QueueTable t;
  t.maxchunks = queue_geo(Wa, Ha, nlevels, rp1, t.geo);
  const std::vector<int4> items = build_queue(t.geo, nlevels, frames, lossy, lag);
  t.nitems = (int)items.size();
  if (cudaMalloc(&t.d_items, items.size() * sizeof(int4)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(t.d_items, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  return &g_queue_tables.emplace(key, t).first->second;
}

template <bool LOSSY, typename OT, bool SRC_U8, bool QS1>
void launch_queue_k(const FwdQueueArgs& q, cudaStream_t st) {
  static const int resident = [] {  // CTAs the whole GPU holds at once: a persistent grid needs no more
    int dev = 0, sms = 148, per_sm = PCS_FWDQ_MINB;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dwt_forward_queue_kernel<LOSSY, OT, SRC_U8, QS1>, 128, 0);
    return sms * std::max(per_sm, 1);
  }();
  const int grid = std::min(resident, (q.nitems + 3) / 4);
  dwt_forward_queue_kernel<LOSSY, OT, SRC_U8, QS1><<<grid, 128, 0, st>>>(q);
}
template <bool LOSSY, typename OT, bool SRC_U8>
void launch_queue(const FwdQueueArgs& q, cudaStream_t st) {
  if (q.lv[0].qs == 1.0f) launch_queue_k<LOSSY, OT, SRC_U8, true>(q, st);
  else launch_queue_k<LOSSY, OT, SRC_U8, false>(q, st);
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
  int nstrips, nchunks, rp;
};

// 5/3 horizontal synthesis (DWTGenerator.cu:79-85,295-308): x = s0,d0,s1,d1,... -> e0,o0,e1,o1,...
template <int N>
__device__ __forceinline__ void hinv53(int (&x)[N], bool left, bool right) {
  constexpr int P = N / 2;
  int pd = __shfl_up_sync(kFullMask, x[N - 1], 1);
  if (left) pd = x[1];
  int e[P];
#pragma unroll
  for (int k = 0; k < P; ++k) e[k] = x[2 * k] - (((k ? x[2 * k - 1] : pd) + x[2 * k + 1] + 2) >> 2);
  int ne = __shfl_down_sync(kFullMask, e[0], 1);
  if (right) ne = e[P - 1];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    x[2 * k + 1] = x[2 * k + 1] + ((e[k] + (k + 1 < P ? e[k + 1] : ne)) >> 1);
    x[2 * k] = e[k];
  }
}

// 9/7 horizontal synthesis (DWTGenerator.cu:110-122,326-339); divisions by K and 1/K as multiplications by
// the rounded reciprocals, as in dwt.cu and the oracle.
template <int N>
__device__ __forceinline__ void hinv97(float (&x)[N], float rK, float rIK, bool left, bool right) {
  constexpr int P = N / 2;
  float e[P], o[P];
#pragma unroll
  for (int k = 0; k < P; ++k) o[k] = __fmul_rn(x[2 * k + 1], rK);
  float p1 = __shfl_up_sync(kFullMask, o[P - 1], 1);
  if (left) p1 = o[0];
#pragma unroll
  for (int k = 0; k < P; ++k) e[k] = __fmaf_rn(-__fadd_rn(k ? o[k - 1] : p1, o[k]), kD, __fmul_rn(x[2 * k], rIK));
  float n0 = __shfl_down_sync(kFullMask, e[0], 1);
  if (right) n0 = e[P - 1];
#pragma unroll
  for (int k = 0; k < P; ++k) o[k] = __fmaf_rn(-__fadd_rn(e[k], k + 1 < P ? e[k + 1] : n0), kG, o[k]);
  p1 = __shfl_up_sync(kFullMask, o[P - 1], 1);
  if (left) p1 = o[0];
#pragma unroll
  for (int k = P - 1; k >= 0; --k) e[k] = __fmaf_rn(-__fadd_rn(k ? o[k - 1] : p1, o[k]), kB, e[k]);
  n0 = __shfl_down_sync(kFullMask, e[0], 1);
  if (right) n0 = e[P - 1];
#pragma unroll
  for (int k = 0; k < P; ++k) o[k] = __fmaf_rn(-__fadd_rn(e[k], k + 1 < P ? e[k + 1] : n0), kA, o[k]);
#pragma unroll
  for (int k = 0; k < P; ++k) { x[2 * k] = e[k]; x[2 * k + 1] = o[k]; }
}

// The same on both rows of a row pair at once (packed pairs, see hfwd97_rows).
template <int N>
__device__ __forceinline__ void hinv97_rows(float (&lo)[N], float (&hi)[N], bool left, bool right) {
  constexpr int P = N / 2;
  const float rK = 1.0f / kK, rIK = 1.0f / kIK;
  // fma(-(a+b), k, x) == fma(a+b, -k, x) bit for bit, so the negation moves into the constants
  const f2 rK2 = pk(rK, rK), rIK2 = pk(rIK, rIK), nD = pk(-kD, -kD), nG = pk(-kG, -kG), nB = pk(-kB, -kB), nA = pk(-kA, -kA);
  f2 x[N], e[P], o[P];
#pragma unroll
  for (int j = 0; j < N; ++j) x[j] = pk(lo[j], hi[j]);
#pragma unroll
  for (int k = 0; k < P; ++k) o[k] = mul2(x[2 * k + 1], rK2);
  f2 p1 = shfl_up2(o[P - 1]);
  if (left) p1 = o[0];
#pragma unroll
  for (int k = 0; k < P; ++k) e[k] = fma2(add2(k ? o[k - 1] : p1, o[k]), nD, mul2(x[2 * k], rIK2));
  f2 n0 = shfl_down2(e[0]);
  if (right) n0 = e[P - 1];
#pragma unroll
  for (int k = 0; k < P; ++k) o[k] = fma2(add2(e[k], k + 1 < P ? e[k + 1] : n0), nG, o[k]);
  p1 = shfl_up2(o[P - 1]);
  if (left) p1 = o[0];
#pragma unroll
  for (int k = P - 1; k >= 0; --k) e[k] = fma2(add2(k ? o[k - 1] : p1, o[k]), nB, e[k]);
  n0 = shfl_down2(e[0]);
  if (right) n0 = e[P - 1];
#pragma unroll
  for (int k = 0; k < P; ++k) o[k] = fma2(add2(e[k], k + 1 < P ? e[k + 1] : n0), nA, o[k]);
#pragma unroll
  for (int k = 0; k < P; ++k) { upk(e[k], lo[2 * k], hi[2 * k]); upk(o[k], lo[2 * k + 1], hi[2 * k + 1]); }
}

// vertical synthesis pipelines.  step(m) takes row pair m (after the horizontal pass) and releases an even
// row kEven(m) and an odd row kOdd(m).
template <int N>
struct IV53 {  // x[2n] = s[n] - ((d[n-1]+d[n]+2)>>2);  x[2n-1] = d[n-1] + ((x[2n-2]+x[2n])>>1)  (DWTGenerator.cu:79-85,160-181)
  int dp[N], xp[N];
  static constexpr int kLead = 1, kTail = 0;  // steps before the first / after the last pair of the chunk
  __device__ __forceinline__ static int even_row(int m) { return 2 * m; }
  __device__ __forceinline__ static int odd_row(int m) { return 2 * m - 1; }
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < N; ++j) { dp[j] = 0; xp[j] = 0; }
  }
  __device__ __forceinline__ void step(const int (&s)[N], const int (&d)[N], int (&re)[N], int (&ro)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const int xe = s[j] - ((dp[j] + d[j] + 2) >> 2);
      ro[j] = dp[j] + ((xp[j] + xe) >> 1);
      re[j] = xe; xp[j] = xe; dp[j] = d[j];
    }
  }
};
template <int N>
struct IV97 {  // DWTGenerator.cu:110-122,230-272
  float opp[N], e1p[N], o1pp[N], e0pp[N];
  static constexpr int kLead = 2, kTail = 1;
  __device__ __forceinline__ static int even_row(int m) { return 2 * (m - 1); }
  __device__ __forceinline__ static int odd_row(int m) { return 2 * m - 3; }
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < N; ++j) { opp[j] = 0.f; e1p[j] = 0.f; o1pp[j] = 0.f; e0pp[j] = 0.f; }
  }
  __device__ __forceinline__ void step(const float (&lo)[N], const float (&hi)[N], float (&re)[N], float (&ro)[N]) {
    const float rK = 1.0f / kK, rIK = 1.0f / kIK;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float op = __fmul_rn(hi[j], rK);
      const float e1 = __fmaf_rn(-__fadd_rn(opp[j], op), kD, __fmul_rn(lo[j], rIK));
      const float o1p = __fmaf_rn(-__fadd_rn(e1p[j], e1), kG, opp[j]);    // o1[m-1]
      const float e0p = __fmaf_rn(-__fadd_rn(o1pp[j], o1p), kB, e1p[j]);  // x[2(m-1)]
      ro[j] = __fmaf_rn(-__fadd_rn(e0pp[j], e0p), kA, o1pp[j]);           // x[2(m-2)+1]
      re[j] = e0p;
      opp[j] = op; e1p[j] = e1; o1pp[j] = o1p; e0pp[j] = e0p;
    }
  }
};
// IV97::step, two columns per instruction
template <int N>
__device__ __forceinline__ void iv97_step_packed(IV97<N>& v, const float (&lo)[N], const float (&hi)[N], float (&re)[N], float (&ro)[N]) {
  const float rK = 1.0f / kK, rIK = 1.0f / kIK;
  const f2 rK2 = pk(rK, rK), rIK2 = pk(rIK, rIK), nD = pk(-kD, -kD), nG = pk(-kG, -kG), nB = pk(-kB, -kB), nA = pk(-kA, -kA);
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    const f2 opp = pk(v.opp[j], v.opp[j + 1]), e1p = pk(v.e1p[j], v.e1p[j + 1]), o1pp = pk(v.o1pp[j], v.o1pp[j + 1]), e0pp = pk(v.e0pp[j], v.e0pp[j + 1]);
    const f2 op = mul2(pk(hi[j], hi[j + 1]), rK2);
    const f2 e1 = fma2(add2(opp, op), nD, mul2(pk(lo[j], lo[j + 1]), rIK2));
    const f2 o1p = fma2(add2(e1p, e1), nG, opp);     // o1[m-1]
    const f2 e0p = fma2(add2(o1pp, o1p), nB, e1p);   // x[2(m-1)]
    const f2 xo = fma2(add2(e0pp, e0p), nA, o1pp);   // x[2(m-2)+1]
    upk(xo, ro[j], ro[j + 1]); upk(e0p, re[j], re[j + 1]);
    upk(op, v.opp[j], v.opp[j + 1]); upk(e1, v.e1p[j], v.e1p[j + 1]); upk(o1p, v.o1pp[j], v.o1pp[j + 1]); upk(e0p, v.e0pp[j], v.e0pp[j + 1]);
  }
}

template <bool LOSSY, int N> struct IVSel { using type = IV53<N>; };
template <int N> struct IVSel<true, N> { using type = IV97<N>; };

// De-quantiser on subband read (DWTGenerator.cu:513-553): 0 -> 0, else (|q|+0.5)*sgn / Q / qs, the two
// divisions as multiplications by rounded reciprocals.  |q| -> float through the 2^23 mantissa trick
// (exact for |q| < 2^23, which the decoder guarantees) instead of an I2F on the quarter-rate pipe.
// Two quantisation indices -> reconstruction values (|q| + 0.5) * sgn(q) / Q [/ qs], 0 for q == 0
// (DWTGenerator.cu:513-553 with the divisions restated as the oracle's reciprocal multiplies).  s = 2q + sgn(q) is
// twice the reconstruction point as an integer (0 stays 0, the sign comes for free); the 1.5 * 2^23 magic number turns
// it into a float with one integer add and one exact fp32 add (|s| < 2^22: indices of 8-bit frames stay below 2^18
// at any depth), and s * (0.5 / Q) rounds exactly like (|q| + 0.5) * (1 / Q) because the factors 2 and 0.5 are exact.
// hrq = 0.5 / Q for the two samples.  5 instructions per sample instead of 8, two of them packed.
template <bool QS1>
__device__ __forceinline__ f2 dequant2(int q0, int q1, f2 hrq, f2 rqs2) {
  const int s0 = 2 * q0 + max(min(q0, 1), -1), s1 = 2 * q1 + max(min(q1, 1), -1);
  f2 v = add2(pk(__int_as_float(0x4B400000 + s0), __int_as_float(0x4B400000 + s1)), pk(-12582912.0f, -12582912.0f));
  v = mul2(v, hrq);
  if (!QS1) v = mul2(v, rqs2);
  return v;
}

__device__ __forceinline__ unsigned sat_u8(float v) {  // clamp(rn(v), 0, 255)
  unsigned r;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

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
      if constexpr (LOSSY) {  // clamp(rn((v + 128) + 0.01)), DecodingEngine.cu:714
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float t0, t1;
          upk(add2(add2(pk(v[j], v[j + 1]), pk(128.0f, 128.0f)), pk(0.01f, 0.01f)), t0, t1);
          b[j] = sat_u8(t0); b[j + 1] = sat_u8(t1);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = (unsigned)min(max(v[j] + 128, 0), 255);  // :728
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
                         int Wc, int Hc, int levels) {
  if (Wc % 8 || Hc % 2 || Wc < 64 || Hc < 16) return false;
  if (levels == 2 && (Wc % 32 || Hc % 16 || Hc < 64)) return false;
  if (reinterpret_cast<uintptr_t>(d_coeff) & 15) return false;
  if (ll && (reinterpret_cast<uintptr_t>(ll) & 15)) return false;
  if (src_u8) {
    if ((reinterpret_cast<uintptr_t>(src) & 7) || src_pitch % 8 || src_fstride % 8 || src_pitch != Wc || H > Hc) return false;
  } else {
    if ((reinterpret_cast<uintptr_t>(src) & 15) || src_pitch % 4 || src_fstride % 4) return false;
  }
  return true;
}

// `levels` (1 or 2) levels of the forward transform, starting at 1-based level `level`, in one launch.
int dwt_forward_fast(const void* src, bool src_u8, long long src_fstride, int src_pitch, int H, int Wc, int Hc, void* d_coeff,
                     long long out_fstride, int pitch, void* d_ll, long long ll_fstride, int ll_pitch, int level, int levels, int wl,
                     bool lossy, float qs, bool out_int, int frames, cudaStream_t st, bool rgb, long long comp_stride) {
  // rgb: src = planar u8 RGB frames (src_fstride per FRAME, comp_stride between colour planes), `frames` counts
  // component planes (3 per frame) and the colour transform is fused into the load (one level only)
  FastFwdArgs a{};
  a.comp_stride = comp_stride;
  a.src = src; a.src_fstride = src_fstride; a.src_pitch = src_pitch; a.H = src_u8 ? H : Hc; a.Wc = Wc; a.Hc = Hc;
  a.out = d_coeff; a.out_fstride = out_fstride; a.pitch = pitch;
  a.ll = d_ll; a.ll_fstride = ll_fstride; a.ll_pitch = ll_pitch;
  a.last = (level + levels - 1 == wl);
  for (int i = 0; i < 4; ++i) { a.q1[i] = kQ[level - 1][i]; a.q2[i] = kQ[level < 10 ? level : 9][i]; }
  a.qs = qs;
  variant_params(Hc, &a.nW1, &a.ls1);
  variant_params(Hc >> 1, &a.nW2, &a.ls2);
  const int eff = (32 - 2 * levels) * 8;
  a.nstrips = (Wc + eff - 1) / eff;
  if (src_u8) {
    if (levels == 2) {
      if (!lossy) launch<false, int, true, 2>(a, frames, st);
      else if (out_int) launch<true, int, true, 2>(a, frames, st);
      else launch<true, float, true, 2>(a, frames, st);
    } else if (rgb) {
      if (!lossy) launch<false, int, true, 1, true>(a, frames, st);
      else if (out_int) launch<true, int, true, 1, true>(a, frames, st);
      else launch<true, float, true, 1, true>(a, frames, st);
    } else {
      if (!lossy) launch<false, int, true, 1>(a, frames, st);
      else if (out_int) launch<true, int, true, 1>(a, frames, st);
      else launch<true, float, true, 1>(a, frames, st);
    }
  } else {
    if (!lossy) launch<false, int, false, 1>(a, frames, st);
    else if (out_int) launch<true, int, false, 1>(a, frames, st);
    else launch<true, float, false, 1>(a, frames, st);
  }
  ++g_launches;
  return (int)cudaGetLastError();
}

// ---- queue path: host side --------------------------------------------------------------------------------------
static size_t queue_scratch_elems(int Wa, int Ha, int wl) {
  size_t n = 0;
  for (int l = 1; l < wl; ++l) n += (size_t)(Wa >> l) * (Ha >> l);
  return n;
}
static size_t queue_ll_offset(int Wa, int Ha, int l) {
  size_t n = 0;
  for (int k = 1; k < l; ++k) n += (size_t)(Wa >> k) * (Ha >> k);
  return n;
}

// Bytes of counter workspace a queue launch over `frames` frames needs (one area per concurrent call).
size_t dwt_queue_ws_bytes(int Wa, int Ha, int wl, int frames) {
  QueueGeo geo[kMaxQLevels];
  const int nl = std::min(wl, kMaxQLevels);
  const int maxchunks = queue_geo(Wa, Ha, nl, queue_rp1(), geo);
  return (1 + (size_t)frames * nl * maxchunks) * sizeof(unsigned);
}

// Levels 1..*done of the forward transform in one persistent launch; *done = 0 (and nothing launched) if the queue
// does not apply to this geometry / batch, in which case the per-level path takes all levels.  src: the u8 frames
// (src_u8) or 32-bit colour-transformed planes [frames][Ha][Wa].
int dwt_forward_queue(const void* src, bool src_u8, int W, int H, int Wa, int Ha, void* d_coeff, void* d_scratch, int wl, bool lossy,
                      float qs, bool out_int, int frames, void* ws, cudaStream_t st, int* done) {
  *done = 0;
  static const int mode = env_int("PCS_DWT_QUEUE", 0);
  static const int min_mpix = env_int("PCS_DWTQ_MIN_MPIX", 8);
  const size_t A = (size_t)Wa * Ha, S = queue_scratch_elems(Wa, Ha, wl);
  if (!mode || !ws || S % 4 || (long long)frames * (long long)A < (long long)min_mpix * 1000000LL || frames >= (1 << 22)) return 0;
  FwdQueueArgs q{};
  int nl = 0;
  for (int l = 1; l <= wl && nl < kMaxQLevels; ++l) {
    const int Wc = Wa >> (l - 1), Hc = Ha >> (l - 1);
    const bool u8 = l == 1 && src_u8;
    const void* lsrc = l == 1 ? src : static_cast<const char*>(d_scratch) + queue_ll_offset(Wa, Ha, l - 1) * 4;
    const long long sfs = u8 ? (long long)W * H : l == 1 ? (long long)A : (long long)S;
    const int sp = u8 ? W : Wc;
    void* ll = l == wl ? nullptr : static_cast<char*>(d_scratch) + queue_ll_offset(Wa, Ha, l) * 4;
    if (!dwt_forward_fast_ok(lsrc, d_coeff, ll, u8, sp, sfs, H, Wc, Hc, 1)) break;
    FastFwdArgs& a = q.lv[nl++];
    a.src = lsrc; a.src_fstride = sfs; a.src_pitch = sp; a.H = u8 ? H : Hc; a.Wc = Wc; a.Hc = Hc;
    a.out = d_coeff; a.out_fstride = (long long)A; a.pitch = Wa;
    a.ll = ll; a.ll_fstride = (long long)S; a.ll_pitch = Wa >> l;
    a.last = l == wl;
    for (int i = 0; i < 4; ++i) { a.q1[i] = kQ[l - 1][i]; a.q2[i] = kQ[l < 10 ? l : 9][i]; }
    a.qs = qs;
    variant_params(Hc, &a.nW1, &a.ls1);
    variant_params(Hc >> 1, &a.nW2, &a.ls2);
  }
  if (nl < 2) return 0;
  const QueueTable* t = queue_table(Wa, Ha, nl, lossy, frames);
  if (!t) return (int)cudaErrorMemoryAllocation;
  for (int i = 0; i < nl; ++i) { q.lv[i].nstrips = t->geo[i].nstrips; q.lv[i].rp = t->geo[i].rp; q.lv[i].nchunks = t->geo[i].nchunks; }
  q.items = t->d_items; q.nitems = t->nitems; q.nlevels = nl; q.maxchunks = t->maxchunks;
  q.ctr = static_cast<unsigned*>(ws);
  cudaError_t e = cudaMemsetAsync(ws, 0, (1 + (size_t)frames * nl * t->maxchunks) * sizeof(unsigned), st);
  if (e != cudaSuccess) return (int)e;
  if (src_u8) {
    if (!lossy) launch_queue<false, int, true>(q, st);
    else if (out_int) launch_queue<true, int, true>(q, st);
    else launch_queue<true, float, true>(q, st);
  } else {
    if (!lossy) launch_queue<false, int, false>(q, st);
    else if (out_int) launch_queue<true, int, false>(q, st);
    else launch_queue<true, float, false>(q, st);
  }
  ++g_launches;
  *done = nl;
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
                     int frames, cudaStream_t st, bool rgb, long long comp_stride) {
  // rgb (dst_u8 only): dst = planar u8 RGB frames (dst_fstride per FRAME, comp_stride between colour planes), `frames`
  // counts frames, d_q / ll hold three component planes per frame
  FastInvArgs a{};
  a.comp_stride = comp_stride;
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
