/*
 * picsong_oracle.c -- CPU restatement of PICSONG's encode/decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (cuda-image-and-video-codec_b200/) links,
 * imports or executes this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may, and there only as the checker / reported baseline.
 *
 * Parity pinning: the reference (13Karl/CUDA-Image-and-Video-codec) has no tests, golden vectors or
 * CPU path.  This restatement is pinned against outputs of the UNMODIFIED reference binary
 * (oracle/_ref/PICSONG_ref, built by oracle/build_ref.sh from /root/reference) run on a B200 by
 * oracle/run_ref_golden.py; those outputs are committed under tests/golden/ (see DESIGN.md).
 *
 * All `file:line` citations are relative to /root/reference/CUDA_ImCod/.  This is a restatement of
 * behaviour in plain scalar C (whole-image lifting, 32 virtual lanes for the bit-plane coder), not a
 * copy of the reference's CUDA code.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CB 64            /* BPC/BPCEngine.cuh:29-36  CBLOCK_WIDTH = 64 (two columns x 32 lanes: fixed) */
/* CBLOCK_LENGTH: 64 rows, or 32 when the reference is rebuilt with -DCBLOCK_LENGTH=32 (BPCEngine.cuh:33-36; BASELINE
 * configs[2]'s sweep point).  With 32 rows the reference still clamps its codeword counter at 4094 / 4095 while a block
 * owns only 2048 words, so it is only defined while no block needs more than 2046 codewords; tests stay inside that. */
static int g_cb_rows = 64;
void pcso_set_cb_rows(int rows) { g_cb_rows = rows == 32 ? 32 : 64; }
#define CBR g_cb_rows
#define CBSZ (CB * CBR)  /* words reserved per codeblock in the sparse codestream */

/* ------------------------------------------------------------------------------------------
 * Geometry and padding
 * ---------------------------------------------------------------------------------------- */

/* SupportFunctions/AuxiliarFunctions.cpp:22-26 fixImageProportions: pad both dims to x64. */
void pcso_adapted_dims(int W, int H, int* Wa, int* Ha) {
  *Wa = (int)ceilf((float)W / CB) * CB;
  *Ha = (int)ceilf((float)H / CBR) * CBR;
}

/* IO/IOManager.ipp:72-112 loadFrameCAdaptedSizes: edge-inclusive mirror, columns then rows:
 * column W+j <- column W-1-j; then row H+r <- row H-1-r (of the already widened rows). */
void pcso_pad_mirror(const uint8_t* in, int W, int H, uint8_t* out, int Wa, int Ha) {
  for (int y = 0; y < H; y++) {
    memcpy(out + (size_t)y * Wa, in + (size_t)y * W, (size_t)W);
    for (int j = 0; j < Wa - W; j++) out[(size_t)y * Wa + W + j] = out[(size_t)y * Wa + W - 1 - j];
  }
  for (int r = 0; r < Ha - H; r++)
    memcpy(out + (size_t)(H + r) * Wa, out + (size_t)(H - 1 - r) * Wa, (size_t)Wa);
}

/* ------------------------------------------------------------------------------------------
 * DWT.  DWT/DWTGenerator.cu computes each level per 64x18 warp tile with overlap; because the
 * lifting dependency radius (2 for 5/3, 4 for 9/7) is covered by the discarded overlap, the
 * result equals whole-image lifting with whole-sample symmetric extension at the ADAPTED image
 * borders (SURVEY.md Annex A.2/A.3, verified there by emulating the tile maths).  That identity
 * is what is restated here.  Mallat placement: DWTGenerator.cu:699-725 (coordinates) and
 * :403-433 (even row/even col -> LL, even row/odd col -> HL, odd row/even col -> LH, odd/odd -> HH).
 * Forward: columns (vertical) then rows (:805-806, :928-929); inverse: rows then columns
 * (:1057-1058, :1118-1119).
 * ---------------------------------------------------------------------------------------- */

/* DWT/DWTGenerator.cu:72-76, :137-157: d[n] = x[2n+1] - ((x[2n]+x[2n+2])>>1), x[N] := x[N-2];
 * s[n] = x[2n] + ((d[n-1]+d[n]+2)>>2), d[-1] := d[0].  Arithmetic shifts on int32. */
static void fwd53_1d(int32_t* x, int n, int st) {
  for (int k = 1; k < n; k += 2) {
    int32_t a = x[(k - 1) * st], c = (k + 1 < n) ? x[(k + 1) * st] : x[(k - 1) * st];
    x[k * st] -= (a + c) >> 1;
  }
  for (int k = 0; k < n; k += 2) {
    int32_t a = (k > 0) ? x[(k - 1) * st] : x[st], c = x[(k + 1) * st];
    x[k * st] += (a + c + 2) >> 2;
  }
}

/* DWT/DWTGenerator.cu:79-85, :160-181 inverse of the above (even samples first, then odd). */
static void inv53_1d(int32_t* x, int n, int st) {
  for (int k = 0; k < n; k += 2) {
    int32_t a = (k > 0) ? x[(k - 1) * st] : x[st], c = x[(k + 1) * st];
    x[k * st] -= (a + c + 2) >> 2;
  }
  for (int k = 1; k < n; k += 2) {
    int32_t a = x[(k - 1) * st], c = (k + 1 < n) ? x[(k + 1) * st] : x[(k - 1) * st];
    x[k * st] += (a + c) >> 1;
  }
}

/* DWT/DWTGenerator.cuh:16-22 */
static const float A97 = -1.586134342059924f, B97 = -0.052980118572961f, G97 = 0.882911075530934f,
                   D97 = 0.443506852043971f, K97 = 1.230174104914001f, IK97 = 0.812893066f;

/* The reference is built with -use_fast_math => flush-to-zero on every fp32 result. */
static inline float ftz(float v) { return (fabsf(v) < 1.17549435e-38f) ? copysignf(0.0f, v) : v; }

/* Forward 9/7 lifting, DWT/DWTGenerator.cu:91-104, :184-227, :311-323.  `*b += (a+c)*k` is one add
 * then one FMA under nvcc's default -fmad=true; step four is FMA then multiply by 1/K; odd samples
 * are then scaled by K.
 *
 * Bit-exactness note (pinned against the reference binary's SASS and its B200 output): the
 * compiler contracts ACROSS the vertical->horizontal boundary.  The last operation of the vertical
 * pass on every sample is a bare multiply by S (S = 1/K on even rows, K on odd rows).  For ODD
 * columns that product has a single use -- the `*b + ...` of horizontal step one -- so it is fused
 * there: o = fma(v, S, (eL+eR)*alpha) with the product (eL+eR)*alpha rounded separately
 * (SASS: FADD, FMUL 1.586..., FFMA v,S,-t).  Even columns are materialised as v*S because
 * they have several uses.  fwd97_vert therefore leaves samples UNSCALED and fwd97_horz applies
 * the row scale S. */
#define XL(k) x[((k) > 0 ? (k) - 1 : 1) * st]
#define XR(k) x[((k) + 1 < n ? (k) + 1 : (k) - 1) * st]
static void fwd97_vert(float* x, int n, int st) { /* columns: steps 1-4, final scaling deferred */
  for (int k = 1; k < n; k += 2) x[k * st] = ftz(fmaf(ftz(XL(k) + XR(k)), A97, x[k * st]));
  for (int k = 0; k < n; k += 2) x[k * st] = ftz(fmaf(ftz(XL(k) + XR(k)), B97, x[k * st]));
  for (int k = 1; k < n; k += 2) x[k * st] = ftz(fmaf(ftz(XL(k) + XR(k)), G97, x[k * st]));
  for (int k = 0; k < n; k += 2) x[k * st] = ftz(fmaf(ftz(XL(k) + XR(k)), D97, x[k * st]));
}
static void fwd97_horz(float* x, int n, float S, int fused_sum) { /* one row; S = deferred vertical scale */
  const int st = 1;
  /* step one needs the UNSCALED even samples when the compiler also fused the own-sample scaling into
   * the neighbour sum (fused_sum, see fwd97_row_variant): sum = fma(v, S, right) instead of v*S + right */
  for (int k = 1; k < n; k += 2) {
    float right = ftz(XR(k) * S), sum;
    if (fused_sum) sum = ftz(fmaf(XL(k), S, right));
    else sum = ftz(ftz(XL(k) * S) + right);
    x[k] = ftz(fmaf(x[k], S, ftz(sum * A97)));
  }
  for (int k = 0; k < n; k += 2) x[k] = ftz(x[k] * S);
  for (int k = 0; k < n; k += 2) x[k] = ftz(fmaf(ftz(XL(k) + XR(k)), B97, x[k]));
  for (int k = 1; k < n; k += 2) x[k] = ftz(fmaf(ftz(XL(k) + XR(k)), G97, x[k]));
  for (int k = 0; k < n; k += 2) x[k] = ftz(ftz(fmaf(ftz(XL(k) + XR(k)), D97, x[k])) * IK97);
  for (int k = 1; k < n; k += 2) x[k] = ftz(x[k] * K97);
}

/* Which rows use the fused neighbour sum.  The reference kernel is fully unrolled over its 64x18
 * register tile and the compiler (nvcc 12.9, sm_100) made that extra contraction only for tile rows
 * 12..17 (recovered by symbolic execution of the binary's SASS).  A row's tile-local
 * index follows the reference's vertical tiling (DWTGenerator.cu:699-744,1281-1285): tiles every 10
 * rows, tile t writes local rows 4..13 (the top tile 0..13, the bottom tile 4..17, aligned to the
 * image bottom).  Rows that both a middle tile and the shifted bottom tile write (when (Hc-8) % 10 != 0)
 * are a WRITE RACE in the reference between two warps whose values can differ in the last bit; the
 * bottom tile lives in the last blocks of the grid and wins whenever the grid spans more than one wave
 * (observed: 512x512 level 1), while on tiny grids the middle tile has been seen to win (256x256 level
 * 1).  The bottom tile is modelled.  This only ever affects 2 image rows per level by <= 1 ulp. */
static int fwd97_row_variant(int y, int Hc) {
  int nWy = (Hc - 8 + 9) / 10;
  if (nWy <= 1) return y >= 12;
  int last_start = (nWy - 1) * 10 - ((10 - ((Hc - 8) % 10)) % 10);
  if (y >= last_start + 4) return (y - last_start) >= 12; /* bottom tile (last blocks of the grid) */
  if (y < 14) return y >= 12;                              /* top tile */
  return (y - 4) % 10 >= 8;                                /* middle tile t = (y-4)/10, local row y-10t */
}

/* DWT/DWTGenerator.cu:110-122, :230-272, :326-339.  The reference's two divisions per sample are
 * -use_fast_math approximate divisions (x * rcp(c)); they are modelled as multiplication by the
 * fp32-rounded reciprocal.  The inverse 9/7 is therefore tolerance-checked, never bit-checked,
 * against the reference (north_star: decoded PSNR within 0.01 dB). */
static void inv97_1d(float* x, int n, int st) {
  const float rK = 1.0f / K97, rIK = 1.0f / IK97;
  for (int k = 1; k < n; k += 2) x[k * st] = ftz(x[k * st] * rK);
  for (int k = 0; k < n; k += 2)
    x[k * st] = ftz(fmaf(-ftz(XL(k) + XR(k)), D97, ftz(x[k * st] * rIK)));
  for (int k = 1; k < n; k += 2) x[k * st] = ftz(fmaf(-ftz(XL(k) + XR(k)), G97, x[k * st]));
  for (int k = 0; k < n; k += 2) x[k * st] = ftz(fmaf(-ftz(XL(k) + XR(k)), B97, x[k * st]));
  for (int k = 1; k < n; k += 2) x[k * st] = ftz(fmaf(-ftz(XL(k) + XR(k)), A97, x[k * st]));
#undef XL
#undef XR
}

/* DWT/DWTGenerator.cuh:168-179 (same table as L2Norm, BPC/BPCEngine.cuh:158-169) */
static const float QSTEP[10][4] = {
    {1.965908f, 1.0112865f, 1.0112865f, 0.52021784f}, {4.1224113f, 1.9968134f, 1.9968134f, 0.96721643f},
    {8.416739f, 4.1833673f, 4.1833673f, 2.0792568f},  {16.935543f, 8.534108f, 8.534108f, 4.3004827f},
    {33.924816f, 17.166693f, 17.166693f, 8.686718f},  {67.87687f, 34.385098f, 34.385098f, 17.41882f},
    {135.76744f, 68.7964f, 68.7964f, 34.860676f},     {271.5416f, 137.60588f, 137.60588f, 69.73287f},
    {543.0866f, 275.21814f, 275.21814f, 139.47136f},  {1086.1624f, 550.43286f, 550.43286f, 278.94202f}};

/* Forward 5/3: u8 pixels (W_a x H_a, already padded) -> int32 Mallat plane out[0..A).
 * Level shift Engines/CodingEngine.cu:581-588 (px - 2^(bps-1)); host loop DWTGenerator.cu:1268-1342. */
/* Level shift of the grey path: 2^(bps-1) for unsigned samples (-bps, default 8), 0 for -signedOrUnsigned 1
 * (Engines/CodingEngine.cu:581-588,644-656; Engines/DecodingEngine.cu:706-729).  Samples are bytes whatever -bps says
 * (IO/IOManager.ipp:72-112 reads unsigned char), and the decoder clamps to [0, 255] in either case. */
static int g_sample_offset = 128;
void pcso_set_sample_offset(int offset) { g_sample_offset = offset; }
static int dwt53_forward_core(int Wa, int Ha, int wl, int32_t* out);
int pcso_dwt53_forward(const uint8_t* px, int Wa, int Ha, int wl, int32_t* out) {
  size_t A = (size_t)Wa * Ha;
  for (size_t i = 0; i < A; i++) out[i] = (int32_t)px[i] - g_sample_offset;
  return dwt53_forward_core(Wa, Ha, wl, out);
}
/* RGB path: the plane is already level-shifted and colour-transformed (Engines/CodingEngine.cu:613) */
int pcso_dwt53_forward_plane(const int32_t* in, int Wa, int Ha, int wl, int32_t* out) {
  memcpy(out, in, (size_t)Wa * Ha * sizeof(int32_t));
  return dwt53_forward_core(Wa, Ha, wl, out);
}
static int dwt53_forward_core(int Wa, int Ha, int wl, int32_t* out) {
  size_t A = (size_t)Wa * Ha;
  int32_t* tmp = (int32_t*)malloc(A * sizeof(int32_t));
  if (!tmp) return -1;
  int Wc = Wa, Hc = Ha;
  for (int l = 1; l <= wl; l++) {
    for (int x = 0; x < Wc; x++) fwd53_1d(out + x, Hc, Wa);           /* vertical */
    for (int y = 0; y < Hc; y++) fwd53_1d(out + (size_t)y * Wa, Wc, 1); /* horizontal */
    for (int y = 0; y < Hc; y++)
      for (int x = 0; x < Wc; x++)
        tmp[(size_t)((y >> 1) + (y & 1) * (Hc >> 1)) * Wc + (x >> 1) + (x & 1) * (Wc >> 1)] =
            out[(size_t)y * Wa + x];
    for (int y = 0; y < Hc; y++) memcpy(out + (size_t)y * Wa, tmp + (size_t)y * Wc, Wc * sizeof(int32_t));
    Wc >>= 1; Hc >>= 1;
  }
  free(tmp);
  return 0;
}

/* Inverse 5/3: int32 Mallat plane -> u8 pixels.  Host loop DWTGenerator.cu:1349-1424; pixel post
 * Engines/DecodingEngine.cu:719-729 clamp(v+128,0,255); narrowing IO/IOManager.ipp:332-335. */
static int dwt53_inverse_core(const int32_t* q, int Wa, int Ha, int wl, uint8_t* px, int32_t* iout);
int pcso_dwt53_inverse(const int32_t* q, int Wa, int Ha, int wl, uint8_t* px) {
  return dwt53_inverse_core(q, Wa, Ha, wl, px, NULL);
}
/* RGB path: the reconstructed plane before the inverse colour transform (no +128, no clamp) */
int pcso_dwt53_inverse_plane(const int32_t* q, int Wa, int Ha, int wl, int32_t* out) {
  return dwt53_inverse_core(q, Wa, Ha, wl, NULL, out);
}
static int dwt53_inverse_core(const int32_t* q, int Wa, int Ha, int wl, uint8_t* px, int32_t* iout) {
  size_t A = (size_t)Wa * Ha;
  int32_t* w = (int32_t*)malloc(A * sizeof(int32_t));
  int32_t* tmp = (int32_t*)malloc(A * sizeof(int32_t));
  if (!w || !tmp) return -1;
  memcpy(w, q, A * sizeof(int32_t));
  for (int l = wl; l >= 1; l--) {
    int Wc = Wa >> (l - 1), Hc = Ha >> (l - 1);
    for (int y = 0; y < Hc; y++)
      for (int x = 0; x < Wc; x++)
        tmp[(size_t)y * Wc + x] =
            w[(size_t)((y >> 1) + (y & 1) * (Hc >> 1)) * Wa + (x >> 1) + (x & 1) * (Wc >> 1)];
    for (int y = 0; y < Hc; y++) inv53_1d(tmp + (size_t)y * Wc, Wc, 1);  /* horizontal */
    for (int x = 0; x < Wc; x++) inv53_1d(tmp + x, Hc, Wc);              /* vertical */
    for (int y = 0; y < Hc; y++) memcpy(w + (size_t)y * Wa, tmp + (size_t)y * Wc, Wc * sizeof(int32_t));
  }
  for (size_t i = 0; i < A; i++) {
    if (iout) iout[i] = w[i];
    if (px) {
      int32_t v = w[i] + g_sample_offset;
      px[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
  free(w); free(tmp);
  return 0;
}

/* Forward 9/7 + quantiser multiply: u8 -> f32 Mallat plane.  Quantiser on subband write,
 * DWTGenerator.cu:403-419: c * Q[l-1][sb] * qs (two fp32 multiplies, left to right); LL only at
 * the last level.  Level index: currentLevel = wl - currentWLevel (:1292), i.e. 0 = finest. */
static int dwt97_forward_core(int Wa, int Ha, int wl, float qs, float* out);
int pcso_dwt97_forward(const uint8_t* px, int Wa, int Ha, int wl, float qs, float* out) {
  size_t A = (size_t)Wa * Ha;
  for (size_t i = 0; i < A; i++) out[i] = (float)((int)px[i] - g_sample_offset);
  return dwt97_forward_core(Wa, Ha, wl, qs, out);
}
/* RGB path: colour-transformed float plane in (Engines/CodingEngine.cu:693) */
int pcso_dwt97_forward_plane(const float* in, int Wa, int Ha, int wl, float qs, float* out) {
  memcpy(out, in, (size_t)Wa * Ha * sizeof(float));
  return dwt97_forward_core(Wa, Ha, wl, qs, out);
}
static int dwt97_forward_core(int Wa, int Ha, int wl, float qs, float* out) {
  size_t A = (size_t)Wa * Ha;
  float* tmp = (float*)malloc(A * sizeof(float));
  if (!tmp) return -1;
  int Wc = Wa, Hc = Ha;
  for (int l = 1; l <= wl; l++) {
    for (int x = 0; x < Wc; x++) fwd97_vert(out + x, Hc, Wa);
    for (int y = 0; y < Hc; y++) fwd97_horz(out + (size_t)y * Wa, Wc, (y & 1) ? K97 : IK97, fwd97_row_variant(y, Hc));
    for (int y = 0; y < Hc; y++)
      for (int x = 0; x < Wc; x++) {
        int sb = (x & 1) + 2 * (y & 1); /* 0 LL, 1 HL, 2 LH, 3 HH */
        float v = out[(size_t)y * Wa + x];
        if (sb != 0 || l == wl) v = ftz(ftz(v * QSTEP[l - 1][sb]) * qs);
        tmp[(size_t)((y >> 1) + (y & 1) * (Hc >> 1)) * Wc + (x >> 1) + (x & 1) * (Wc >> 1)] = v;
      }
    for (int y = 0; y < Hc; y++) memcpy(out + (size_t)y * Wa, tmp + (size_t)y * Wc, Wc * sizeof(float));
    Wc >>= 1; Hc >>= 1;
  }
  free(tmp);
  return 0;
}

/* Inverse 9/7 with de-quantisation on subband read, DWTGenerator.cu:513-553:
 * q==0 -> 0 else (|q| + 0.5f) * sgn(q) / Q[l-1][sb] / qs (fast-math divisions, modelled as
 * reciprocal multiplies); LL de-quantised only at the coarsest level (:1052-1053).
 * Pixel post Engines/DecodingEngine.cu:706-715: clamp(rn(v + 128 + 0.01f), 0, 255).
 * If fout != NULL the un-rounded reconstruction (v, before +128) is also returned. */
int pcso_dwt97_inverse(const int32_t* q, int Wa, int Ha, int wl, float qs, uint8_t* px, float* fout) {
  size_t A = (size_t)Wa * Ha;
  float* w = (float*)malloc(A * sizeof(float));
  float* tmp = (float*)malloc(A * sizeof(float));
  if (!w || !tmp) return -1;
  const float rqs = 1.0f / qs;
  for (int l = wl; l >= 1; l--) {
    int Wc = Wa >> (l - 1), Hc = Ha >> (l - 1);
    for (int y = 0; y < Hc; y++)
      for (int x = 0; x < Wc; x++) {
        int sb = (x & 1) + 2 * (y & 1);
        size_t src = (size_t)((y >> 1) + (y & 1) * (Hc >> 1)) * Wa + (x >> 1) + (x & 1) * (Wc >> 1);
        float v;
        if (sb == 0 && l != wl) {
          v = w[src];
        } else {
          int32_t qi = q[src];
          if (qi == 0) v = 0.0f;
          else {
            float mag = (fabsf((float)qi) + 0.5f) * (qi < 0 ? -1.0f : 1.0f);
            v = ftz(ftz(mag * (1.0f / QSTEP[l - 1][sb])) * rqs);
          }
        }
        tmp[(size_t)y * Wc + x] = v;
      }
    for (int y = 0; y < Hc; y++) inv97_1d(tmp + (size_t)y * Wc, Wc, 1);
    for (int x = 0; x < Wc; x++) inv97_1d(tmp + x, Hc, Wc);
    for (int y = 0; y < Hc; y++) memcpy(w + (size_t)y * Wa, tmp + (size_t)y * Wc, Wc * sizeof(float));
  }
  for (size_t i = 0; i < A; i++) {
    if (fout) fout[i] = w[i];
    if (!px) continue;
    float r = nearbyintf((w[i] + (float)g_sample_offset) + 0.01f); /* __float2int_rn: round half to even */
    r = r > 255.0f ? 255.0f : r;
    r = r < 0.0f ? 0.0f : r;
    px[i] = (uint8_t)r;
  }
  free(w); free(tmp);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Colour transforms of the RGB path.  Planes are u8, already mirror-padded.
 * Plane order out: 0 = Y, 1 = B-G (Cb), 2 = R-G (Cr)  (Engines/CodingEngine.cu:357-405).
 * ---------------------------------------------------------------------------------------- */
static int32_t floordiv4(int32_t v) { return (v >= 0) ? v / 4 : -((-v + 3) / 4); }
/* Engines/CodingEngine.cu:357-379 RGBTransformLossless: floor((R+2G+B)/4), B-G, R-G on level-shifted samples */
void pcso_rct_forward(const uint8_t* r, const uint8_t* g, const uint8_t* b, size_t n, int32_t* y, int32_t* cb, int32_t* cr) {
  for (size_t i = 0; i < n; i++) {
    int32_t R = (int32_t)r[i] - 128, G = (int32_t)g[i] - 128, B = (int32_t)b[i] - 128;
    y[i] = floordiv4(R + 2 * G + B); cb[i] = B - G; cr[i] = R - G;
  }
}
/* Engines/DecodingEngine.cu:599-624: G = Y - floor((Cb+Cr)/4), R = Cr + G, B = Cb + G, +128, clamp */
void pcso_rct_inverse(const int32_t* y, const int32_t* cb, const int32_t* cr, size_t n, uint8_t* r, uint8_t* g, uint8_t* b) {
  for (size_t i = 0; i < n; i++) {
    int32_t G = y[i] - floordiv4(cb[i] + cr[i]), R = cr[i] + G, B = cb[i] + G;
    R += 128; G += 128; B += 128;
    r[i] = (uint8_t)(R < 0 ? 0 : (R > 255 ? 255 : R));
    g[i] = (uint8_t)(G < 0 ? 0 : (G > 255 ? 255 : G));
    b[i] = (uint8_t)(B < 0 ? 0 : (B > 255 ? 255 : B));
  }
}
/* Engines/CodingEngine.cu:384-405 RGBTransformLossy with the matrix of CodingEngine.cuh:25.  Operation shapes as
 * in the reference binary (sm_100 SASS): t = G*m1; t = fma(R, m0, t); t = fma(B, m2, t). */
void pcso_ict_forward(const uint8_t* r, const uint8_t* g, const uint8_t* b, size_t n, float* y, float* cb, float* cr) {
  static const float M[3][3] = {{0.299f, 0.587f, 0.114f}, {-0.168736f, -0.331264f, 0.5f}, {0.5f, -0.418688f, -0.081312f}};
  float* out[3] = {y, cb, cr};
  for (size_t i = 0; i < n; i++) {
    float R = (float)r[i] - 128.0f, G = (float)g[i] - 128.0f, B = (float)b[i] - 128.0f;
    for (int k = 0; k < 3; k++) out[k][i] = ftz(fmaf(B, M[k][2], ftz(fmaf(R, M[k][0], ftz(G * M[k][1])))));
  }
}
/* Engines/DecodingEngine.cu:628-650 with DecodingEngine.cuh:41: rn(M*v + 0.01f) + 128, clamp.  SASS shapes:
 * R = fma(Cr, 1.402, Y); G = fma(Cr, -0.714136, fma(Cb, -0.344136, Y)); B = fma(Cb, 1.772, Y). */
void pcso_ict_inverse(const float* y, const float* cb, const float* cr, size_t n, uint8_t* r, uint8_t* g, uint8_t* b) {
  for (size_t i = 0; i < n; i++) {
    float v[3];
    v[0] = ftz(fmaf(cr[i], 1.402f, y[i]));
    v[1] = ftz(fmaf(cr[i], -0.714136f, ftz(fmaf(cb[i], -0.344136f, y[i]))));
    v[2] = ftz(fmaf(cb[i], 1.772f, y[i]));
    uint8_t* o[3] = {r, g, b};
    for (int k = 0; k < 3; k++) {
      int t = (int)nearbyintf(v[k] + 0.01f) + 128;
      o[k][i] = (uint8_t)(t < 0 ? 0 : (t > 255 ? 255 : t));
    }
  }
}

/* BPC/BPCEngine.cu:49,55 readCoefficients: (int)c -- truncation toward zero (dead-zone). */
void pcso_truncate(const float* c, size_t n, int32_t* q) {
  for (size_t i = 0; i < n; i++) q[i] = (int32_t)c[i];
}

/* ------------------------------------------------------------------------------------------
 * LUT loading (Annex D).  Engines/Engine.cu:101-141,190-210; IO/IOManager.ipp:363-386,404-612.
 * ---------------------------------------------------------------------------------------- */

/* header.txt: eight `NAME;int` lines: bit-planes, subbands, refinement ctx, sign ctx,
 * significance ctx, precision, files, bit-plane files (Engine.cu:190-210). */
int pcso_lut_header(const char* folder, int hdr[8]) {
  char path[1024];
  snprintf(path, sizeof path, "%s/header.txt", folder); /* Engine.cu:105 appends "/header.txt" */
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  char line[256];
  int i = 0;
  while (i < 8 && fgets(line, sizeof line, f)) {
    char* s = strchr(line, ';');
    if (!s) continue;
    hdr[i++] = atoi(s + 1);
  }
  fclose(f);
  return i == 8 ? 0 : -2;
}

int pcso_lut_size(const int hdr[8], int wl) {
  int nBp = hdr[0], nSub = hdr[1], cR = hdr[2], cG = hdr[3], cS = hdr[4];
  return (nSub * nBp * cR * wl + nBp * cR) + (nSub * nBp * cS * wl + nBp * cS) + (nSub * nBp * cG * wl + nBp * cG);
}

/* One section (ref / sig / sign) of IOManager.ipp:446-536.  Rows are `lvl sb bp : v...`.
 * A row whose bp <= previous bp starts a new group: first pad the previous group's planes
 * prev+1..nBp-1 with 64, then advance the base by nBp*ctx; *then* the stop test
 * (lvl+1 > wl && sb > 0); then store at base + bp*ctx.  Everything else stays 0 (the reference
 * relies on fresh-page zeroing of its malloc; SURVEY note 5). */
static int lut_section(const char* path, int32_t* table, int base, int nBp, int ctx, int wl, int cap) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  int i = base, prev = -1, lvl, sb, bp, v[9];
  for (;;) {
    if (fscanf(f, "%d %d %d :", &lvl, &sb, &bp) != 3) break;
    int ok = 1;
    for (int c = 0; c < ctx; c++) if (fscanf(f, "%d", &v[c]) != 1) ok = 0;
    if (!ok) break;
    if (bp <= prev) {
      for (int z = 0; z < (nBp - prev - 1) * ctx; z++) {
        int idx = i + prev * ctx + z + ctx;
        if (idx >= 0 && idx < cap) table[idx] = 64;
      }
      i += nBp * ctx;
    }
    if (lvl + 1 > wl && sb > 0) break;
    prev = bp;
    for (int c = 0; c < ctx; c++) {
      int idx = i + bp * ctx + c;
      if (idx >= 0 && idx < cap) table[idx] = v[c];
    }
  }
  fclose(f);
  return 0;
}

/* Greyscale, k==0, cp==2: with LUT_N_FILES==3 the component suffixes are R,G,B and grey uses
 * "R" (Engine.cu:124-139 loads i+1 -> "R","G","B"; CodingEngine.cu:662 passes table 0);
 * with LUT_N_FILES==1 the suffix is empty (IOManager.ipp:433-444).  File suffix `_0` = k==0. */
int pcso_lut_load_component(const char* folder, int wl, int component, int hdr[8], int32_t* table, int cap);
int pcso_lut_load(const char* folder, int wl, int hdr[8], int32_t* table, int cap) {
  return pcso_lut_load_component(folder, wl, 0, hdr, table, cap);
}
/* component 0,1,2 -> files "R","G","B" (Engine.cu:124-139: the RGB path codes component i with table i) */
int pcso_lut_load_component_file(const char* folder, int wl, int component, int file, int hdr[8], int32_t* table, int cap);
int pcso_lut_load_component(const char* folder, int wl, int component, int hdr[8], int32_t* table, int cap) {
  return pcso_lut_load_component_file(folder, wl, component, 0, hdr, table, cap);
}
/* k > 0 (Engine.cu:12-100): the tables of the bit-plane files `_0 .. _(AMOUNT_OF_BITPLANE_FILES-1)` back to back,
 * each pcso_lut_size() ints long (the file index only enters the file name, IOManager.ipp:445-447);
 * initializeLUTPointers then adds s * size for LUT set s (BPCEngine.cu:333). */
int pcso_lut_load_component_k(const char* folder, int wl, int component, int hdr[8], int32_t* table, int cap) {
  if (pcso_lut_header(folder, hdr)) return -1;
  int n = pcso_lut_size(hdr, wl), files = hdr[7];
  if (files < 1 || (long long)n * files > cap) return -2;
  for (int j = 0; j < files; j++) {
    int r = pcso_lut_load_component_file(folder, wl, component, j, hdr, table + (size_t)j * n, n);
    if (r < 0) return r;
  }
  return n * files;
}
int pcso_lut_load_component_file(const char* folder, int wl, int component, int file, int hdr[8], int32_t* table, int cap) {
  if (pcso_lut_header(folder, hdr)) return -1;
  int nBp = hdr[0], nSub = hdr[1], cR = hdr[2], cG = hdr[3], cS = hdr[4];
  int n = pcso_lut_size(hdr, wl);
  if (n > cap) return -2;
  memset(table, 0, (size_t)n * sizeof(int32_t));
  static const char* kComp[3] = {"R", "G", "B"};
  if (component < 0 || component > 2) return -4;
  const char* comp = (hdr[6] == 1) ? "" : kComp[component];
  int R = nSub * nBp * cR * wl + nBp * cR, S = nSub * nBp * cS * wl + nBp * cS;
  char path[1024];
  /* -LUTFolder is used as a plain string prefix for the table files (must end in '/') */
  snprintf(path, sizeof path, "%sref%s.txt_%d", folder, comp, file);
  if (lut_section(path, table, 0, nBp, cR, wl, n)) return -3;
  snprintf(path, sizeof path, "%ssig%s.txt_%d", folder, comp, file);
  if (lut_section(path, table, R, nBp, cS, wl, n)) return -3;
  snprintf(path, sizeof path, "%ssign%s.txt_%d", folder, comp, file);
  if (lut_section(path, table, R + S, nBp, cG, wl, n)) return -3;
  return n;
}

/* ------------------------------------------------------------------------------------------
 * BPC-PaCo bit-plane coder, 2 coding passes, k = 0 (Annex B).  32 virtual lanes per codeblock.
 * ---------------------------------------------------------------------------------------- */

typedef struct {
  int nBp, nSub, cR, cG, cS, prec; /* header.txt values */
  int wl;
  const int32_t* lut;
  int lutN;
  float k;     /* complexity-scalability factor (-k); 0 = every plane by the two coding passes */
} bpc_cfg;

/* BPCEngine.cu:143-170 findSubband.  NOTE: the kernel calls it with the THREAD's coordinates
 * (:1975-1976, :1990): X = block X + 2*lane, Y = block Y.  The class -- and with it the three LUT
 * pointers -- is therefore per LANE: in a block wider than its subband (e.g. block 0 when the LL
 * band is narrower than 64) lanes right of the subband edge use the HL tables of a finer level.
 * (Pinned by the reference binary: blocks 0 of the golden streams only match with this rule.) */
static void find_subband(int X, int Y, int Wa, int Ha, int wl, int* lvl, int* sb) {
  for (int a = 1; a <= wl; a++) {
    if (X >= (Wa >> a) || Y >= (Ha >> a)) {
      *lvl = a - 1;
      *sb = (X >= (Wa >> a)) ? ((Y >= (Ha >> a)) ? 2 : 0) : 1;
      return;
    }
  }
  *lvl = wl; *sb = 0;
}

/* BPCEngine.cu:329-350 initializeLUTPointers for each of the 32 lanes of a block at (X, Y). */
static void lane_pointers(const bpc_cfg* cfg, int X, int Y, int Wa, int Ha, int MSB, int* refP, int* sigP, int* sgnP) {
  int R = cfg->nSub * cfg->nBp * cfg->cR * cfg->wl + cfg->nBp * cfg->cR;
  int Sg = cfg->nSub * cfg->nBp * cfg->cS * cfg->wl + cfg->nBp * cfg->cS;
  for (int l = 0; l < 32; l++) {
    int lvl, sb;
    find_subband(X + 2 * l, Y, Wa, Ha, cfg->wl, &lvl, &sb);
    int g = (lvl * cfg->nSub + sb) * cfg->nBp + MSB;
    refP[l] = g * cfg->cR;
    sigP[l] = R + g * cfg->cS;
    sgnP[l] = R + Sg + g * cfg->cG;
  }
}

/* BPCEngine.cu:252-293 computeSignContext(h, v) */
static int sign_ctx_hv(int h, int v) {
  if (h == 0) return v == 0 ? 0 : (v > 0 ? 2 : 3);
  if (h > 0) return v == 0 ? 4 : (v > 0 ? 6 : 0);
  return v == 0 ? 5 : (v > 0 ? 1 : 7);
}
/* BPCEngine.cu:296-308: +-1 by bit 0 for SIGNIFICANT (bit 31) neighbours */
static int sgn_of(uint32_t s) { return (s >> 31) ? ((s & 1) ? -1 : 1) : 0; }
static int sign_ctx(uint32_t up, uint32_t left, uint32_t right, uint32_t down) {
  return sign_ctx_hv(sgn_of(left) + sgn_of(right), sgn_of(up) + sgn_of(down));
}

/* LUT read with the reference's unguarded flat indexing (MSB >= 15 walks into the next group);
 * reads beyond the table are undefined in the reference -- here they return 0. */
static inline int lut_at(const bpc_cfg* c, int idx) { return (idx >= 0 && idx < c->lutN) ? c->lut[idx] : 0; }

typedef struct {
  uint32_t S[32][2 * CB + 4];   /* state words, [lane][2*row + col] (Annex B.2) */
  uint32_t lower[32], size[32]; /* per-lane interval */
  uint32_t slot[32];            /* reserved codeword (absolute index into the block's words) */
  uint32_t code[32];            /* decoder: current codeword */
  int count;                    /* per-warp codeword counter (codeStreamShared, BPCEngine.cu:1970) */
  int32_t* words;               /* block's words: words[0] = MSB, words[1+j] = codeword j */
  int prec;
} cbstate;

/* BPCEngine.cu:371-399 / :405-442, first half: lanes in `need` (those executing THIS call whose
 * interval is exhausted) restart and reserve slots in ascending lane order:
 * slot = 1 + min(count + rank, 4094); then count = min(count + howMany, 4095). */
static void ac_reserve(cbstate* s, uint32_t need, int decode) {
  if (!need) return;
  int rank = 0;
  for (int l = 0; l < 32; l++)
    if ((need >> l) & 1) {
      s->lower[l] = 0;
      s->size[l] = 0xFFFF;
      int j = s->count + rank;
      if (j > 4094) j = 4094;
      s->slot[l] = 1 + (uint32_t)j;
      if (decode) s->code[l] = (uint32_t)s->words[s->slot[l]];
      rank++;
    }
  s->count += rank;
  if (s->count > 4095) s->count = 4095;
}

/* BPCEngine.cu:386-398 */
static inline void ac_encode(cbstate* s, int l, uint32_t b, int p) {
  uint32_t t = ((s->size[l] * (uint32_t)p) >> s->prec) + b;
  if (b == 0) s->size[l] = t;
  else { s->size[l] -= t; s->lower[l] += t; }
  if (s->size[l] == 0) s->words[s->slot[l]] = (int32_t)s->lower[l];
}
/* BPCEngine.cu:424-440 */
static inline uint32_t ac_decode(cbstate* s, int l, int p) {
  uint32_t t = ((s->size[l] * (uint32_t)p) >> s->prec) + 1;
  uint32_t lim = s->lower[l] + t;
  if (s->code[l] >= lim) { s->size[l] -= t; s->lower[l] = lim; return 1; }
  s->size[l] = t - 1;
  return 0;
}

/* Neighbour snapshot for one (row, col) step, taken before any lane updates in this step
 * (shuffles executed by all lanes before the divergent part, BPCEngine.cu:446-461, 770-843).
 * n[0..7] = up-left, up, up-right, left, right, bottom-left, bottom, bottom-right. */
static void neighbours(const cbstate* s, int l, int i, int c, uint32_t n[8]) {
  const uint32_t(*S)[2 * CB + 4] = s->S;
#define AT(lane, row, col) (((row) < 0 || (row) >= CBR || (lane) < 0 || (lane) > 31) ? 0u : S[lane][2 * (row) + (col)])
  if (c == 0) {
    n[0] = AT(l - 1, i - 1, 1); n[1] = AT(l, i - 1, 0); n[2] = AT(l, i - 1, 1);
    n[3] = AT(l - 1, i, 1);                              n[4] = AT(l, i, 1);
    n[5] = AT(l - 1, i + 1, 1); n[6] = AT(l, i + 1, 0); n[7] = AT(l, i + 1, 1);
  } else {
    n[0] = AT(l, i - 1, 0); n[1] = AT(l, i - 1, 1); n[2] = AT(l + 1, i - 1, 0);
    n[3] = AT(l, i, 0);                              n[4] = AT(l + 1, i, 0);
    n[5] = AT(l, i + 1, 0); n[6] = AT(l, i + 1, 1); n[7] = AT(l + 1, i + 1, 0);
  }
#undef AT
}
static inline int sig_ctx(const uint32_t n[8]) { /* BPCEngine.cu:222-230 */
  int c = 0;
  for (int k = 0; k < 8; k++) c += (int)(n[k] >> 31);
  return c;
}

/* ---- k > 0: the lowest planes of a block are coded coefficient by coefficient ("bulk" mode, BPCEngine.cu:1285-1662) ---- */
/* Planes [0, c) go to the bulk pass, c = max(floor(MSB * (k / L2Norm)), 0) (:1683-1692); L2Norm is the quantiser table
 * (BPCEngine.cuh:158-169 == QSTEP), row = the lane's level (LL: level - 1), column 3 - subband as the reference indexes it. */
static int bulk_planes(const bpc_cfg* cfg, int lvl, int sb, int MSB) {
  float norm = (cfg->wl == lvl) ? QSTEP[lvl - 1 > 0 ? lvl - 1 : 0][0] : QSTEP[lvl][3 - sb];
  int c = (int)floorf((float)MSB * (cfg->k / norm));
  return c > 0 ? c : 0;
}
/* computeContextBulk :236-246: neighbours that became significant at plane >= B (bits 28:24 of the state word) */
static inline int sig_ctx_bulk(const uint32_t n[8], int B) {
  int c = 0;
  for (int k = 0; k < 8; k++) c += (int)(((n[k] >> 24) & 31) >= (uint32_t)B);
  return c;
}
/* computeSignContextBulk :310-322: only neighbours significant since plane >= cur contribute their sign */
static int sgn_of_bulk(uint32_t s, int cur) { return ((s >> 31) && ((s >> 24) & 31) >= (uint32_t)cur) ? ((s & 1) ? -1 : 1) : 0; }
static int sign_ctx_bulk(uint32_t up, uint32_t left, uint32_t right, uint32_t down, int cur) {
  return sign_ctx_hv(sgn_of_bulk(left, cur) + sgn_of_bulk(right, cur), sgn_of_bulk(up, cur) + sgn_of_bulk(down, cur));
}

/* One codeblock, encode.  Kernel body BPCEngine.cu:1973-2025; Encode :1668-1721; SPP :770-843 +
 * :490-516; MRP :1249-1261 + :726-736.  q = coefficient plane (already (int)-truncated for the
 * lossy path), words = this block's 4096-word slot (pre-filled with -1 by the caller, :2431-2440). */
static int encode_block(const bpc_cfg* cfg, const int32_t* q, int Wa, int Ha, int cb, int32_t* words) {
  int nbx = Wa / CB, X = (cb % nbx) * CB, Y = (cb / nbx) * CBR;
  cbstate* s = (cbstate*)calloc(1, sizeof(cbstate));
  s->words = words; s->prec = cfg->prec;
  uint32_t ormag = 0;
  for (int l = 0; l < 32; l++)
    for (int i = 0; i < CBR; i++)
      for (int c = 0; c < 2; c++) {
        int32_t v = q[(size_t)(Y + i) * Wa + X + 2 * l + c];
        uint32_t mag = (uint32_t)(v < 0 ? -(int64_t)v : v);
        s->S[l][2 * i + c] = (mag << 1) + (v < 0 ? 1u : 0u); /* :41-63 */
        ormag |= (s->S[l][2 * i + c] >> 1);
      }
  int MSB = 32; /* :176-192: 32 - ffs(brev(max)); all-zero block => 32 */
  for (int b = 31; b >= 0; b--) if ((ormag >> b) & 1) { MSB = b; break; }
  words[0] = MSB;
  if (MSB != 32) {
    int refP[32], sigP[32], sgnP[32]; /* per lane, :329-350 */
    lane_pointers(cfg, X, Y, Wa, Ha, MSB, refP, sigP, sgnP);
    uint32_t n[32][8];
    /* k > 0 (:1683-1697): planes below `consec` are left to the bulk pass and the LUT set min(consec, MSB) is used.
     * The reference computes both per LANE; a block that straddles subbands makes its lanes leave the plane loop at
     * different planes, which the reference does not survive (its decoder aborts on such streams) -- lane 0 decides
     * here, exact for every block that lies inside one subband. */
    int consec = 0;
    if (cfg->k > 0.0f) {
      int lvl, sb;
      find_subband(X, Y, Wa, Ha, cfg->wl, &lvl, &sb);
      consec = bulk_planes(cfg, lvl, sb, MSB);
      int set = consec < MSB ? consec : MSB;
      int per = pcso_lut_size((const int[8]){cfg->nBp, cfg->nSub, cfg->cR, cfg->cG, cfg->cS, 0, 0, 0}, cfg->wl);
      for (int l = 0; l < 32; l++) { refP[l] += set * per; sigP[l] += set * per; sgnP[l] += set * per; }
    }
    int bp;
    for (bp = MSB; bp >= consec; bp--) {
      /* ---- SPP ---- */
      for (int i = 0; i < CBR; i++)
        for (int c = 0; c < 2; c++) {
          uint32_t act = 0, one = 0, need;
          for (int l = 0; l < 32; l++) {
            neighbours(s, l, i, c, n[l]);
            if (!(s->S[l][2 * i + c] >> 31)) act |= 1u << l;
          }
          /* significance call */
          need = 0;
          for (int l = 0; l < 32; l++) if (((act >> l) & 1) && s->size[l] == 0) need |= 1u << l;
          ac_reserve(s, need, 0);
          for (int l = 0; l < 32; l++)
            if ((act >> l) & 1) {
              uint32_t b = (s->S[l][2 * i + c] >> (bp + 1)) & 1;
              ac_encode(s, l, b, lut_at(cfg, sigP[l] + sig_ctx(n[l])));
              if (b) one |= 1u << l;
            }
          /* sign call for lanes whose bit was 1 */
          need = 0;
          for (int l = 0; l < 32; l++) if (((one >> l) & 1) && s->size[l] == 0) need |= 1u << l;
          ac_reserve(s, need, 0);
          for (int l = 0; l < 32; l++)
            if ((one >> l) & 1) {
              uint32_t* w = &s->S[l][2 * i + c];
              *w |= 1u << 31;
              *w |= (uint32_t)bp << 24;
              int sc = sign_ctx(n[l][1], n[l][3], n[l][4], n[l][6]);
              ac_encode(s, l, ((*w & 1) == (uint32_t)(sc & 1)) ? 0 : 1, lut_at(cfg, sgnP[l] + (sc >> 1)));
            }
        }
      /* ---- MRP ---- */
      for (int i = 0; i < CBR; i++)
        for (int c = 0; c < 2; c++) {
          uint32_t act = 0, need = 0;
          for (int l = 0; l < 32; l++) {
            uint32_t w = s->S[l][2 * i + c];
            if ((w >> 29) & 1) act |= 1u << l;
            else if (w >> 31) s->S[l][2 * i + c] |= 1u << 29;
          }
          for (int l = 0; l < 32; l++) if (((act >> l) & 1) && s->size[l] == 0) need |= 1u << l;
          ac_reserve(s, need, 0);
          for (int l = 0; l < 32; l++)
            if ((act >> l) & 1) ac_encode(s, l, (s->S[l][2 * i + c] >> (bp + 1)) & 1, lut_at(cfg, refP[l]));
        }
      for (int l = 0; l < 32; l++) { refP[l] -= cfg->cR; sigP[l] -= cfg->cS; sgnP[l] -= cfg->cG; } /* :353-358 */
    }
    if (bp >= 0) {
      /* ---- bulk pass from plane B = bp down to 0, one coefficient at a time (:1285-1420, :1640-1650): rows in order,
       * column 0 then column 1; context and neighbour snapshot are taken ONCE per coefficient, before its planes ---- */
      const int B = bp;
      for (int i = 0; i < CBR; i++)
        for (int c = 0; c < 2; c++) {
          int ctx[32];
          uint32_t any = 0;
          for (int l = 0; l < 32; l++) {
            neighbours(s, l, i, c, n[l]);
            if (!(s->S[l][2 * i + c] >> 31)) any = 1;
          }
          for (int l = 0; l < 32; l++) ctx[l] = any ? (B != 0 ? sig_ctx_bulk(n[l], B) : sig_ctx(n[l])) : 0;
          for (int cur = B; cur >= 0; cur--) {
            const int d = B - cur;
            uint32_t sigm = 0, insig = 0, one = 0, need = 0;
            for (int l = 0; l < 32; l++) { if (s->S[l][2 * i + c] >> 31) sigm |= 1u << l; else insig |= 1u << l; }
            /* refinement of the already significant (:1292-1295) */
            for (int l = 0; l < 32; l++) if (((sigm >> l) & 1) && s->size[l] == 0) need |= 1u << l;
            ac_reserve(s, need, 0);
            for (int l = 0; l < 32; l++)
              if ((sigm >> l) & 1) ac_encode(s, l, (s->S[l][2 * i + c] >> (cur + 1)) & 1, lut_at(cfg, refP[l] - cfg->cR * d));
            /* significance of the others (:1297-1299) */
            need = 0;
            for (int l = 0; l < 32; l++) if (((insig >> l) & 1) && s->size[l] == 0) need |= 1u << l;
            ac_reserve(s, need, 0);
            for (int l = 0; l < 32; l++)
              if ((insig >> l) & 1) {
                uint32_t b = (s->S[l][2 * i + c] >> (cur + 1)) & 1;
                ac_encode(s, l, b, lut_at(cfg, sigP[l] + ctx[l] - cfg->cS * d));
                if (b) one |= 1u << l;
              }
            /* sign of those that just became significant (:1301-1310) */
            need = 0;
            for (int l = 0; l < 32; l++) if (((one >> l) & 1) && s->size[l] == 0) need |= 1u << l;
            ac_reserve(s, need, 0);
            for (int l = 0; l < 32; l++)
              if ((one >> l) & 1) {
                uint32_t* w = &s->S[l][2 * i + c];
                *w |= 1u << 31;
                *w |= (uint32_t)cur << 24;
                int sc = sign_ctx_bulk(n[l][1], n[l][3], n[l][4], n[l][6], cur);
                ac_encode(s, l, ((*w & 1) == (uint32_t)(sc & 1)) ? 0 : 1, lut_at(cfg, sgnP[l] + (sc >> 1) - cfg->cG * d));
              }
          }
        }
    }
    for (int l = 0; l < 32; l++) words[s->slot[l]] = (int32_t)s->lower[l]; /* flush :1719 */
  }
  int size = s->count + 1; /* :2010 */
  if (size == CBSZ) {      /* expansion fallback :2015-2024, :1904-1912 */
    for (int l = 0; l < 32; l++)
      for (int k = 0; k < 2 * CBR; k++) words[l * 2 * CBR + k] = (int32_t)(s->S[l][k] & 0xFFFF);
  }
  free(s);
  return size;
}

/* One codeblock, decode.  Kernel BPCEngine.cu:2173-2213; Decode :1776-1837; SPP :930-1003 +
 * :558-594; MRP :1267-1279 + :742-762; output :94-111. */
static void decode_block(const bpc_cfg* cfg, const int32_t* words, int size, int Wa, int Ha, int cb, int32_t* q) {
  int nbx = Wa / CB, X = (cb % nbx) * CB, Y = (cb / nbx) * CBR;
  cbstate* s = (cbstate*)calloc(1, sizeof(cbstate));
  s->words = (int32_t*)words; s->prec = cfg->prec;
  int MSB = words[0];
  if (size == CBSZ) { /* :2197-2200, :1914-1922 copyEntireCodeblock */
    for (int l = 0; l < 32; l++)
      for (int k = 0; k < 2 * CBR; k++) s->S[l][k] = (uint32_t)words[l * 2 * CBR + k];
  } else if (MSB != 32) {
    int refP[32], sigP[32], sgnP[32];
    lane_pointers(cfg, X, Y, Wa, Ha, MSB, refP, sigP, sgnP);
    uint32_t mask = 0x3u << MSB; /* :1791 */
    if (MSB == 0) mask &= 0x2;   /* :1804 */
    uint32_t n[32][8];
    int consec = 0;              /* k > 0: see encode_block (:1793-1810) */
    if (cfg->k > 0.0f) {
      int lvl, sb;
      find_subband(X, Y, Wa, Ha, cfg->wl, &lvl, &sb);
      consec = bulk_planes(cfg, lvl, sb, MSB);
      int set = consec < MSB ? consec : MSB;
      int per = pcso_lut_size((const int[8]){cfg->nBp, cfg->nSub, cfg->cR, cfg->cG, cfg->cS, 0, 0, 0}, cfg->wl);
      for (int l = 0; l < 32; l++) { refP[l] += set * per; sigP[l] += set * per; sgnP[l] += set * per; }
    }
    int bp;
    for (bp = MSB; bp >= consec; bp--) {
      for (int i = 0; i < CBR; i++)
        for (int c = 0; c < 2; c++) {
          uint32_t act = 0, one = 0, need = 0;
          for (int l = 0; l < 32; l++) {
            neighbours(s, l, i, c, n[l]);
            if (!(s->S[l][2 * i + c] >> 31)) act |= 1u << l;
          }
          for (int l = 0; l < 32; l++) if (((act >> l) & 1) && s->size[l] == 0) need |= 1u << l;
          ac_reserve(s, need, 1);
          for (int l = 0; l < 32; l++)
            if ((act >> l) & 1)
              if (ac_decode(s, l, lut_at(cfg, sigP[l] + sig_ctx(n[l])))) one |= 1u << l;
          need = 0;
          for (int l = 0; l < 32; l++) if (((one >> l) & 1) && s->size[l] == 0) need |= 1u << l;
          ac_reserve(s, need, 1);
          for (int l = 0; l < 32; l++)
            if ((one >> l) & 1) {
              uint32_t* w = &s->S[l][2 * i + c];
              *w |= mask;
              *w |= 1u << 31;
              *w |= (uint32_t)bp << 24;
              int sc = sign_ctx(n[l][1], n[l][3], n[l][4], n[l][6]);
              uint32_t sym = ac_decode(s, l, lut_at(cfg, sgnP[l] + (sc >> 1)));
              sym = ((sym & 1) == (uint32_t)(sc & 1)) ? 0 : 1;
              *w |= sym;
            }
        }
      for (int i = 0; i < CBR; i++)
        for (int c = 0; c < 2; c++) {
          uint32_t act = 0, need = 0;
          for (int l = 0; l < 32; l++) {
            uint32_t w = s->S[l][2 * i + c];
            if ((w >> 29) & 1) act |= 1u << l;
            else if (w >> 31) s->S[l][2 * i + c] |= 1u << 29;
          }
          for (int l = 0; l < 32; l++) if (((act >> l) & 1) && s->size[l] == 0) need |= 1u << l;
          ac_reserve(s, need, 1);
          for (int l = 0; l < 32; l++)
            if ((act >> l) & 1) {
              uint32_t sym = ac_decode(s, l, lut_at(cfg, refP[l]));
              uint32_t* w = &s->S[l][2 * i + c];
              *w &= ~mask;
              *w |= mask & (((sym << 1) + 1) << bp);
            }
        }
      mask >>= 1;               /* :1822 */
      if (bp == 1) mask = 0x2;  /* :1824 */
      for (int l = 0; l < 32; l++) { refP[l] -= cfg->cR; sigP[l] -= cfg->cS; sgnP[l] -= cfg->cG; }
    }
    if (bp >= 0) { /* bulk pass, inverse of the one in encode_block (:1487-1530, :1535-1650) */
      const int B = bp;
      const uint32_t mask0 = mask;
      for (int i = 0; i < CBR; i++)
        for (int c = 0; c < 2; c++) {
          int ctx[32];
          uint32_t any = 0;
          for (int l = 0; l < 32; l++) {
            neighbours(s, l, i, c, n[l]);
            if (!(s->S[l][2 * i + c] >> 31)) any = 1;
          }
          for (int l = 0; l < 32; l++) ctx[l] = any ? (B != 0 ? sig_ctx_bulk(n[l], B) : sig_ctx(n[l])) : 0;
          uint32_t m = mask0;  /* the mask is passed by value: every coefficient restarts from the plane-B mask */
          for (int cur = B; cur >= 0; cur--) {
            const int dd = B - cur;
            uint32_t sigm = 0, insig = 0, one = 0, need = 0;
            for (int l = 0; l < 32; l++) { if (s->S[l][2 * i + c] >> 31) sigm |= 1u << l; else insig |= 1u << l; }
            for (int l = 0; l < 32; l++) if (((sigm >> l) & 1) && s->size[l] == 0) need |= 1u << l;
            ac_reserve(s, need, 1);
            for (int l = 0; l < 32; l++)
              if ((sigm >> l) & 1) {
                uint32_t sym = ac_decode(s, l, lut_at(cfg, refP[l] - cfg->cR * dd));
                uint32_t* w = &s->S[l][2 * i + c];
                *w &= ~m;
                *w |= m & (((sym << 1) + 1) << cur);
              }
            need = 0;
            for (int l = 0; l < 32; l++) if (((insig >> l) & 1) && s->size[l] == 0) need |= 1u << l;
            ac_reserve(s, need, 1);
            for (int l = 0; l < 32; l++)
              if ((insig >> l) & 1)
                if (ac_decode(s, l, lut_at(cfg, sigP[l] + ctx[l] - cfg->cS * dd))) one |= 1u << l;
            need = 0;
            for (int l = 0; l < 32; l++) if (((one >> l) & 1) && s->size[l] == 0) need |= 1u << l;
            ac_reserve(s, need, 1);
            for (int l = 0; l < 32; l++)
              if ((one >> l) & 1) {
                uint32_t* w = &s->S[l][2 * i + c];
                *w |= m;
                *w |= 1u << 31;
                *w |= (uint32_t)cur << 24;
                int sc = sign_ctx_bulk(n[l][1], n[l][3], n[l][4], n[l][6], cur);
                uint32_t sym = ac_decode(s, l, lut_at(cfg, sgnP[l] + (sc >> 1) - cfg->cG * dd));
                sym = ((sym & 1) == (uint32_t)(sc & 1)) ? 0 : 1;
                *w |= sym;
              }
            m >>= 1;
            if (cur == 1) m = 0x2;
          }
        }
    }
  }
  for (int l = 0; l < 32; l++)
    for (int i = 0; i < CBR; i++)
      for (int c = 0; c < 2; c++) {
        uint32_t w = s->S[l][2 * i + c];
        int32_t v = (int32_t)((w & 0xFFFFFF) >> 1);
        q[(size_t)(Y + i) * Wa + X + 2 * l + c] = (w & 1) ? -v : v;
      }
  free(s);
}

static void mkcfg(bpc_cfg* c, const int hdr[8], int wl, const int32_t* lut, int lutN) {
  c->nBp = hdr[0]; c->nSub = hdr[1]; c->cR = hdr[2]; c->cG = hdr[3]; c->cS = hdr[4]; c->prec = hdr[5];
  c->wl = wl; c->lut = lut; c->lutN = lutN; c->k = 0.0f;
}

/* Whole frame.  cs = sparse int32 codestream [A] (filled with -1 here, BPCEngine.cu:2431-2440),
 * sizes = int32[nCB].  Codeblocks cb0 <= cb < cb1 only (cb1 <= 0 means all) so that callers can
 * time / check a bounded sample. */
int pcso_bpc_encode(const int32_t* q, int Wa, int Ha, int wl, const int hdr[8], const int32_t* lut,
                    int lutN, int32_t* cs, int32_t* sizes, int cb0, int cb1) {
  bpc_cfg cfg; mkcfg(&cfg, hdr, wl, lut, lutN);
  int nCB = (Wa / CB) * (Ha / CBR);
  if (cb1 <= 0 || cb1 > nCB) cb1 = nCB;
  for (int cb = cb0; cb < cb1; cb++) {
    int32_t* w = cs + (size_t)cb * CBSZ;
    for (int k = 0; k < CBSZ; k++) w[k] = -1;
    sizes[cb] = encode_block(&cfg, q, Wa, Ha, cb, w);
  }
  return 0;
}

/* The same with the complexity-scalability factor k (ENCODER only; lut = pcso_lut_load_component_k tables). */
int pcso_bpc_encode_k(const int32_t* q, int Wa, int Ha, int wl, const int hdr[8], const int32_t* lut,
                      int lutN, float k, int32_t* cs, int32_t* sizes, int cb0, int cb1) {
  bpc_cfg cfg; mkcfg(&cfg, hdr, wl, lut, lutN);
  cfg.k = k;
  int nCB = (Wa / CB) * (Ha / CBR);
  if (cb1 <= 0 || cb1 > nCB) cb1 = nCB;
  for (int cb = cb0; cb < cb1; cb++) {
    int32_t* w = cs + (size_t)cb * CBSZ;
    for (int j = 0; j < CBSZ; j++) w[j] = -1;
    sizes[cb] = encode_block(&cfg, q, Wa, Ha, cb, w);
  }
  return 0;
}

int pcso_bpc_decode_k(const int32_t* cs, const int32_t* sizes, int Wa, int Ha, int wl, const int hdr[8],
                      const int32_t* lut, int lutN, float k, int32_t* q, int cb0, int cb1) {
  bpc_cfg cfg; mkcfg(&cfg, hdr, wl, lut, lutN);
  cfg.k = k;
  int nCB = (Wa / CB) * (Ha / CBR);
  if (cb1 <= 0 || cb1 > nCB) cb1 = nCB;
  for (int cb = cb0; cb < cb1; cb++) decode_block(&cfg, cs + (size_t)cb * CBSZ, sizes[cb], Wa, Ha, cb, q);
  return 0;
}

int pcso_bpc_decode(const int32_t* cs, const int32_t* sizes, int Wa, int Ha, int wl, const int hdr[8],
                    const int32_t* lut, int lutN, int32_t* q, int cb0, int cb1) {
  bpc_cfg cfg; mkcfg(&cfg, hdr, wl, lut, lutN);
  int nCB = (Wa / CB) * (Ha / CBR);
  if (cb1 <= 0 || cb1 > nCB) cb1 = nCB;
  for (int cb = cb0; cb < cb1; cb++) decode_block(&cfg, cs + (size_t)cb * CBSZ, sizes[cb], Wa, Ha, cb, q);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Header (Annex C.1) and bit-stream builder (Annex C.2)
 * ---------------------------------------------------------------------------------------- */

/* BitStreamBuilder/BitStreamBuilder.cpp:35-94 setExtraInformation */
void pcso_header_pack(int W, int H, int components, int cp, int cbW, int cbH, int wl, int bitDepth,
                      int lossy, float qs, int isRGB, int endian, int bps, int isSigned, int frames,
                      float k, uint16_t out[9]) {
  int sz = W * H * components;
  int qi = (int)(qs * 10000);
  out[0] = (uint16_t)(sz & 0xFFFF);
  out[1] = (uint16_t)(((uint32_t)sz & 0xFFFF0000u) >> 16);
  out[2] = (uint16_t)((cp == 2 ? 0 : 1) | (cbH << 1) | (cbW << 8) | ((wl & 1) << 15));
  out[3] = (uint16_t)(((wl & 7) >> 1) | (bitDepth << 3) | ((lossy ? 1 : 0) << 10) | ((qi & 31) << 11));
  out[4] = (uint16_t)((qi >> 5) | ((components & 127) << 9));
  out[5] = (uint16_t)((components >> 7) | ((isRGB ? 1 : 0) << 7) | (H << 8));
  out[6] = (uint16_t)((H >> 8) | (endian << 8) | (bps << 9) | ((isSigned ? 1 : 0) << 14) | ((frames & 1) << 15));
  out[7] = (uint16_t)((frames >> 1) & 0xFFFF);
  out[8] = (uint16_t)((int)(k * 1000));
}

/* Engines/DecodingEngine.cu:567-585 getExtraInformation -> 16 ints:
 * [0] size [1] cp [2] cbH [3] cbW [4] wl [5] bitDepth [6] lossy [7] qs*10000 [8] components
 * [9] isRGB [10] H [11] endian [12] bps [13] signed [14] frames [15] k*1000 */
void pcso_header_unpack(const uint16_t e[9], int b[16]) {
  b[0] = e[0] | (e[1] << 16);
  b[1] = (e[2] & 1) == 1 ? 3 : 2;
  b[2] = (e[2] >> 1) & 127;
  b[3] = (e[2] >> 8) & 127;
  b[4] = ((e[2] >> 15) & 1) | ((e[3] & 7) << 1);
  b[5] = (e[3] >> 3) & 127;
  b[6] = (e[3] >> 10) & 1;
  b[7] = ((e[3] >> 11) & 31) | ((e[4] & 511) << 5);
  b[8] = ((e[4] >> 9) & 127) | ((e[5] & 127) << 9);
  b[9] = (e[5] >> 7) & 1;
  b[10] = ((e[5] >> 8) & 255) | ((e[6] & 255) << 8);
  b[11] = (e[6] >> 8) & 1;
  b[12] = (e[6] >> 9) & 31;
  b[13] = (e[6] >> 14) & 1;
  b[14] = ((e[6] >> 15) & 1) | (e[7] << 1);
  b[15] = e[8];
}

/* BitStreamBuilder.cu:198-229 binarySearchLUTBSB + :234-265: 256-step coarse index. */
static void bsb_index(const int32_t* P, int n, int V, int* range, int32_t table[257]) {
  const int steps = 256;
  *range = (int)ceilf((float)V / (float)steps);
  table[0] = 0;
  for (int t = 1; t <= steps; t++) {
    int value = *range * t, lo = 0, hi = n, mid = (lo + hi) / 2;
    while (mid < hi) {
      if (value - P[mid] >= 0) lo = mid + 1; else hi = mid;
      mid = (lo + hi) / 2;
    }
    table[t] = mid;
  }
}

/* BitStreamBuilder.cu:33-101 binarySearchLUTBS, for codeword index t. P has n+1 readable ints. */
static void bsb_locate(const int32_t* P, int n, int t, int range, const int32_t table[257], int* cb,
                       int* bsStart, int* csStart, int* realStart, int* size, int* headerThread) {
  int pos = (int)((float)t / (float)range), lo, hi;
  if (pos == 0) { lo = table[0]; hi = table[1]; }
  else if (pos > 0 && pos < 256) { lo = table[pos - 1]; hi = table[pos + 1]; }
  else { lo = table[pos - 1]; hi = n - 1; }
  int mid = (lo + hi) / 2;
  while (mid < hi) {
    if (t - P[mid] >= 0) lo = mid + 1; else hi = mid;
    mid = (lo + hi) / 2;
  }
  if (P[hi] <= t && hi + 1 < n) hi++;
  if (hi == 0) {
    *bsStart = t; *cb = 0; *csStart = 0; *size = P[0]; *headerThread = 0; *realStart = t;
  } else {
    *cb = hi; *bsStart = t - hi; *csStart = CBSZ * hi; *realStart = CBSZ * hi + t - P[hi - 1];
    *size = P[hi] - P[hi - 1]; *headerThread = P[hi - 1];
  }
}

/* createBitStream, BitStreamBuilder.cpp:100-114 + BitStreamBuilder.cu:106-137,270-323: literal
 * thread-per-codeword restatement.  hdr9 == NULL => frame index > 0 (0xFFFF x 9).
 * Returns the frame length in uint16 units = P[n-1] + n + 10 (one trailing 0xFFFF). */
int pcso_bsb_pack(const int32_t* cs, const int32_t* sizes, int nCB, const uint16_t* hdr9, uint16_t* out) {
  int32_t* P = (int32_t*)malloc((size_t)(nCB + 1) * sizeof(int32_t));
  int32_t table[257];
  int acc = 0, range;
  for (int i = 0; i < nCB; i++) { acc += sizes[i]; P[i] = acc; }
  P[nCB] = 0; /* the reference allocates nCB+1 ints; the extra one is never initialised */
  int V = acc, total = V + nCB + 10;
  memset(out, 0xFF, (size_t)total * sizeof(uint16_t));
  if (hdr9) memcpy(out, hdr9, 9 * sizeof(uint16_t));
  bsb_index(P, nCB, V, &range, table);
  int headerOffset = -1 + 9 + nCB * 2;
  for (int t = 0; t < V; t++) {
    int cb, bs, csS, real, size, ht;
    bsb_locate(P, nCB, t, range, table, &cb, &bs, &csS, &real, &size, &ht);
    if (t == ht) { out[9 + 2 * cb] = (uint16_t)cs[csS]; out[9 + 2 * cb + 1] = (uint16_t)size; }
    else out[headerOffset + bs] = (uint16_t)cs[real];
  }
  free(P);
  return total;
}

/* Closed form of the same layout (Annex C.2): out[9+2cb] = word 0, out[10+2cb] = size, the
 * size-1 codewords contiguous at 9 + 2n + (P[cb-1] - cb).  Used by tests to show the literal
 * search above and the closed form agree. */
int pcso_bsb_pack_closed(const int32_t* cs, const int32_t* sizes, int nCB, const uint16_t* hdr9, uint16_t* out) {
  int V = 0;
  for (int i = 0; i < nCB; i++) V += sizes[i];
  int total = V + nCB + 10;
  memset(out, 0xFF, (size_t)total * sizeof(uint16_t));
  if (hdr9) memcpy(out, hdr9, 9 * sizeof(uint16_t));
  int Pprev = 0;
  for (int cb = 0; cb < nCB; cb++) {
    out[9 + 2 * cb] = (uint16_t)cs[(size_t)cb * CBSZ];
    out[10 + 2 * cb] = (uint16_t)sizes[cb];
    for (int j = 1; j < sizes[cb]; j++) out[9 + 2 * nCB + (Pprev - cb) + (j - 1)] = (uint16_t)cs[(size_t)cb * CBSZ + j];
    Pprev += sizes[cb];
  }
  return total;
}

/* createCodeStream, BitStreamBuilder.cpp:119-153 + BitStreamBuilder.cu:142-171: sizes parsed from
 * the per-block header (in[10+2cb]); cs memset to -1; literal per-codeword scatter, u16 -> int32
 * zero-extended. */
int pcso_bsb_unpack(const uint16_t* in, int nCB, int32_t* cs, int32_t* sizes) {
  int32_t* P = (int32_t*)malloc((size_t)(nCB + 1) * sizeof(int32_t));
  int32_t table[257];
  int acc = 0, range;
  for (int i = 0; i < nCB; i++) { sizes[i] = in[10 + 2 * i]; acc += sizes[i]; P[i] = acc; }
  P[nCB] = 0;
  memset(cs, 0xFF, (size_t)nCB * CBSZ * sizeof(int32_t));
  int V = acc;
  bsb_index(P, nCB, V, &range, table);
  int headerOffset = -1 + 9 + nCB * 2;
  for (int t = 0; t < V; t++) {
    int cb, bs, csS, real, size, ht;
    bsb_locate(P, nCB, t, range, table, &cb, &bs, &csS, &real, &size, &ht);
    if (t == ht) cs[real] = in[9 + 2 * cb];
    else cs[real] = in[headerOffset + bs];
  }
  free(P);
  return V + nCB + 10;
}
