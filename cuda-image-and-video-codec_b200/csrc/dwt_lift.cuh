// dwt_lift.cuh -- device building blocks shared by the streaming DWT kernels (dwt_fast.cu: one level per launch;
// dwt_fused.cu: levels 1+2 in one CTA): packed fp32 pairs, the horizontal / vertical lifting steps of the 5/3 and 9/7
// transforms in the operation shapes of the reference binary (DWT/DWTGenerator.cu:72-122,137-339 of the reference),
// quantiser / de-quantiser, pixel pack / unpack, subband stores, launch-shape helpers.  Bit-exactness notes are with each
// function; oracle/picsong_oracle.c is the restatement they are tested against.
#pragma once
#include <algorithm>
#include <type_traits>

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

// hfwd97_rows with the row variant chosen at run time (warp-uniform): the two variants differ only in the first lifting
// sum of each pair (FMA-contracted or not), so one copy of the code with a uniform branch around those N/2
// instructions serves both -- half the instruction-cache footprint of calling the two template variants.
template <int N>
__device__ __forceinline__ void hfwd97_rows_rt(float (&lo)[N], float (&hi)[N], bool left, bool right, bool fused) {
  constexpr int P = N / 2;
  const f2 S = pk(kIK, kK), A2 = pk(kA, kA), B2 = pk(kB, kB), G2 = pk(kG, kG), D2 = pk(kD, kD), IK2 = pk(kIK, kIK), K2 = pk(kK, kK);
  f2 x[N], e[P], o[P];
#pragma unroll
  for (int j = 0; j < N; ++j) x[j] = pk(lo[j], hi[j]);
#pragma unroll
  for (int k = 0; k < P; ++k) e[k] = mul2(x[2 * k], S);
  f2 n0 = shfl_down2(e[0]);
  if (right) n0 = e[P - 1];
  if (fused) {
#pragma unroll
    for (int k = 0; k < P; ++k) o[k] = fma2(x[2 * k], S, k + 1 < P ? e[k + 1] : n0);
  } else {
#pragma unroll
    for (int k = 0; k < P; ++k) o[k] = add2(e[k], k + 1 < P ? e[k + 1] : n0);
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

// Two independent row pairs A and B through hfwd97_rows_rt in ONE instruction stream, stage by stage: the shuffle latency of
// one pair (four dependent ~25-cycle exchanges) is covered by the arithmetic of the other.  At 3-4 eligible warps per
// scheduler the lifting chains are latency-bound otherwise.
template <int N>
__device__ __forceinline__ void hfwd97_rows2(float (&loA)[N], float (&hiA)[N], float (&loB)[N], float (&hiB)[N], bool left, bool right,
                                             bool fusedA, bool fusedB) {
  constexpr int P = N / 2;
  const f2 S = pk(kIK, kK), A2 = pk(kA, kA), B2 = pk(kB, kB), G2 = pk(kG, kG), D2 = pk(kD, kD), IK2 = pk(kIK, kIK), K2 = pk(kK, kK);
  f2 xa[N], ea[P], oa[P], xb[N], eb[P], ob[P];
#pragma unroll
  for (int j = 0; j < N; ++j) { xa[j] = pk(loA[j], hiA[j]); xb[j] = pk(loB[j], hiB[j]); }
#pragma unroll
  for (int k = 0; k < P; ++k) { ea[k] = mul2(xa[2 * k], S); eb[k] = mul2(xb[2 * k], S); }
  f2 na = shfl_down2(ea[0]), nb = shfl_down2(eb[0]);
  if (right) { na = ea[P - 1]; nb = eb[P - 1]; }
  if (fusedA) {
#pragma unroll
    for (int k = 0; k < P; ++k) oa[k] = fma2(xa[2 * k], S, k + 1 < P ? ea[k + 1] : na);
  } else {
#pragma unroll
    for (int k = 0; k < P; ++k) oa[k] = add2(ea[k], k + 1 < P ? ea[k + 1] : na);
  }
  if (fusedB) {
#pragma unroll
    for (int k = 0; k < P; ++k) ob[k] = fma2(xb[2 * k], S, k + 1 < P ? eb[k + 1] : nb);
  } else {
#pragma unroll
    for (int k = 0; k < P; ++k) ob[k] = add2(eb[k], k + 1 < P ? eb[k + 1] : nb);
  }
#pragma unroll
  for (int k = 0; k < P; ++k) { oa[k] = fma2(xa[2 * k + 1], S, mul2(oa[k], A2)); ob[k] = fma2(xb[2 * k + 1], S, mul2(ob[k], A2)); }
  f2 pa = shfl_up2(oa[P - 1]), pb = shfl_up2(ob[P - 1]);
  if (left) { pa = oa[0]; pb = ob[0]; }
#pragma unroll
  for (int k = P - 1; k >= 0; --k) { ea[k] = fma2(add2(k ? oa[k - 1] : pa, oa[k]), B2, ea[k]); eb[k] = fma2(add2(k ? ob[k - 1] : pb, ob[k]), B2, eb[k]); }
  na = shfl_down2(ea[0]); nb = shfl_down2(eb[0]);
  if (right) { na = ea[P - 1]; nb = eb[P - 1]; }
#pragma unroll
  for (int k = 0; k < P; ++k) { oa[k] = fma2(add2(ea[k], k + 1 < P ? ea[k + 1] : na), G2, oa[k]); ob[k] = fma2(add2(eb[k], k + 1 < P ? eb[k + 1] : nb), G2, ob[k]); }
  pa = shfl_up2(oa[P - 1]); pb = shfl_up2(ob[P - 1]);
  if (left) { pa = oa[0]; pb = ob[0]; }
#pragma unroll
  for (int k = P - 1; k >= 0; --k) {
    ea[k] = mul2(fma2(add2(k ? oa[k - 1] : pa, oa[k]), D2, ea[k]), IK2);
    eb[k] = mul2(fma2(add2(k ? ob[k - 1] : pb, ob[k]), D2, eb[k]), IK2);
  }
#pragma unroll
  for (int k = 0; k < P; ++k) {
    upk(ea[k], loA[2 * k], hiA[2 * k]); upk(mul2(oa[k], K2), loA[2 * k + 1], hiA[2 * k + 1]);
    upk(eb[k], loB[2 * k], hiB[2 * k]); upk(mul2(ob[k], K2), loB[2 * k + 1], hiB[2 * k + 1]);
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

// The same step with the first two lifting stages packed (two columns per instruction) and the last two scalar.  The
// horizontal pass that follows wants (low row, high row) of a COLUMN side by side in a register pair; a packed last
// stage leaves (column j, column j+1) of a ROW side by side instead and costs a 2x2 register transpose per column pair
// (~5 MOVs).  Scalar results can be allocated straight into the pairs the horizontal pass reads: 16 more arithmetic
// instructions per step, ~38 fewer moves.  Same operations, same rounding (all .rn.ftz), bit-identical.
template <int N>
__device__ __forceinline__ void v97_step_mixed(V97<N>& v, const float (&xo)[N], const float (&x2)[N], float (&lo)[N], float (&hi)[N]) {
  const f2 A2 = pk(kA, kA), B2 = pk(kB, kB);
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    const f2 xe = pk(v.xe[j], v.xe[j + 1]), xn = pk(x2[j], x2[j + 1]), o1p = pk(v.o1p[j], v.o1p[j + 1]);
    const f2 o1 = fma2(add2(xe, xn), A2, pk(xo[j], xo[j + 1]));
    const f2 e1 = fma2(add2(o1p, o1), B2, xe);
    float e1c[2];
    upk(e1, e1c[0], e1c[1]);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float o2 = __fmaf_rn(__fadd_rn(v.e1p[j + c], e1c[c]), kG, v.o1p[j + c]);   // o2[m-1]
      lo[j + c] = __fmaf_rn(__fadd_rn(v.o2pp[j + c], o2), kD, v.e1p[j + c]);          // e2[m-1]
      hi[j + c] = o2;
      v.o2pp[j + c] = o2; v.e1p[j + c] = e1c[c];
    }
    v.xe[j] = x2[j]; v.xe[j + 1] = x2[j + 1];
    upk(o1, v.o1p[j], v.o1p[j + 1]);
  }
}

template <bool LOSSY, int N> struct VSel { using type = V53<N>; using T = int; };
template <int N> struct VSel<true, N> { using type = V97<N>; using T = float; };

// 8 pixels of a u8 row -> T (px - 128; for 5/3 the shift is applied to the LL output instead, which is
// exact because every high-pass step is shift-invariant and the low-pass steps carry the offset through)
// `offset` = the level shift 2^(bps-1) (128 for 8-bit samples, 0 for signed input; Engines/CodingEngine.cu:581-588)
__device__ __forceinline__ void unpack8(uint2 p, float (&v)[8], int offset = 128) {
  const float c = -(8388608.0f + (float)offset);  // (2^23 + b) - (2^23 + offset), exact for offset <= 2^15
  const f2 C2 = pk(c, c);
#pragma unroll
  for (int j = 0; j < 4; j += 2) {
    upk(add2(pk(__uint_as_float(__byte_perm(p.x, 0x4B000000u, 0x7540u + j)), __uint_as_float(__byte_perm(p.x, 0x4B000000u, 0x7541u + j))), C2),
        v[j], v[j + 1]);
    upk(add2(pk(__uint_as_float(__byte_perm(p.y, 0x4B000000u, 0x7540u + j)), __uint_as_float(__byte_perm(p.y, 0x4B000000u, 0x7541u + j))), C2),
        v[4 + j], v[5 + j]);
  }
}
__device__ __forceinline__ void unpack8(uint2 p, int (&v)[8], int offset = 128) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] = (int)__byte_perm(p.x, 0u, 0x4440u + j) - offset;
    v[4 + j] = (int)__byte_perm(p.y, 0u, 0x4440u + j) - offset;
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

}  // namespace
}  // namespace pcs
