// dwt_fused.cu -- levels 1 AND 2 of the forward DWT in one CTA for sm_100a: the LL plane of level 1 never reaches HBM.
//
// Replaces the first two iterations of the reference's host loop DWTEngine::DWTForward (DWT/DWTGenerator.cu:1268-1342:
// one kernel + one cudaStreamSynchronize per level, LL ping-pong through global memory) together with offsetImage
// (Engines/CodingEngine.cu:581-588) and the mirror padding (IO/IOManager.ipp:72-112).  Arithmetic, border rules, row
// variants and quantiser are those of dwt_lift.cuh / oracle/picsong_oracle.c, bit for bit: the per-level kernels of
// dwt_fast.cu produce the same bytes (tests/test_gpu_parity.py::test_fused_levels_match_per_level_launches).
//
// Why: per-level launches move 7.4 B/px (the 4 B/px LL1 plane is written by level 1 and read back by level 2) where the
// transform needs 5 B/px; both launches already run at ~86 % of the HBM peak, so only removing that round trip helps.
//
// (Measured steps of this design are in DESIGN.md 4.1.)
// Design (one CTA = 6 warps, a column strip of 912 pixels x a chunk of rows, streaming downwards):
//   * warps 0-3, "producers": level 1.  A warp owns 256 pixel columns (8 per lane, lanes 0 / 31 are halo), keeps the
//     vertical lifting state in registers and does the horizontal lifting with four shuffles per row pair -- the
//     pipeline of dwt_fast.cu.  Its pixel rows arrive through a private ring of TMA tile loads
//     (cp.async.bulk.tensor.3d -> mbarrier complete_tx; lane 0 issues two 256-byte row boxes per step, kDepth steps
//     ahead): no per-lane address arithmetic, out-of-image columns are zero-filled by the TMA unit, the row index
//     folding of the mirror padding is done once per row by the issuing lane.  HL1 / LH1 / HH1 go to HBM with 16-byte
//     streaming stores; the LL1 row goes into a shared-memory ring of kLLSlots rows.
//   * warps 4-5, "consumers": level 2 at FULL lane density.  A warp owns 256 LL1 columns (8 per lane) of the ring, so
//     level 2 costs a quarter of level 1's instructions instead of the half an in-warp fusion pays (4 samples per
//     lane).  Hand-over per ring row: "row stored" is a named barrier (producers bar.arrive, consumers bar.sync: the
//     consumers wait about half of the time by design and a parked warp costs no issue slots); "row read" is an
//     mbarrier (elected consumer lane arrives, producers try_wait, which normally succeeds at once).  A named barrier
//     is a rendezvous of EVERY warp that bar.syncs on it: used for "row read" it locked the four producers into step
//     at every row (12 % of the kernel's stall samples sat there; the mbarrier is worth 1.1-1.4 %, profiles/r2n_*).
//     There is no CTA-wide barrier in the loop.  HL2 / LH2 / HH2 go to the Mallat plane, LL2 to the scratch plane the
//     per-level kernels continue from (or, quantised, to the Mallat plane when wl == 2).
//   * CTA pitch 912 = 2 x 456 LL1 columns: producers deliver 4 x 120 = 480 valid LL1 columns, the consumers' halo lanes
//     take 8 on either side, and 912 keeps every TMA box start a multiple of 16 bytes (the unit rejects anything else).
//     At the right image edge only the warps that have columns run.
#include <cuda.h>

#include <cstdlib>

#include "dwt_lift.cuh"

namespace pcs {

namespace {

constexpr int kProd = 4, kCons = 2, kThreads = 32 * (kProd + kCons);
constexpr int kPitchLL = 456, kPitch = 2 * kPitchLL;  // LL1 / pixel columns a CTA owns (912 = 16 * 57: TMA boxes start 16-byte aligned)
constexpr int kStages = 4;                            // two-step stages (4 rows x 256 B) a producer's TMA ring runs ahead
constexpr int kLLSlots = 6;                           // LL1 rows in the hand-over ring (two named barriers each)
constexpr int kLLRow = 512;                           // 32-bit words per ring row (480 written, 500 read)

struct FusedArgs {
  const unsigned char* src; long long src_fstride; int src_pitch;  // u8 frames (for the one row each chunk starts from)
  int H, Wc, Hc;                                      // original height; level-1 input size = adapted frame size
  void* out; long long out_fstride; int pitch;        // Mallat plane
  void* ll; long long ll_fstride; int ll_pitch;       // LL2 destination (scratch) unless `last`
  int last;                                           // wl == 2: LL2 goes to `out`, quantised
  float q1[4], q2[4], qs;                             // quantiser steps of levels 1 / 2: LL, HL, LH, HH
  int nW1, ls1, nW2, ls2;                             // row-variant tiling of the reference per level (9/7 only)
  int nctax, nchunks, rp;                             // rp: level-1 row pairs per chunk (multiple of 4)
  int offset;                                         // level shift 2^(bps-1), 0 for signed samples
};

// ---- mbarrier / TMA (PTX; SASS: SYNCS.*, UTMALDG) ----------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LAB_DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "LAB_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// one box of the u8 frames at (x, y, frame) -> shared memory (256 columns x 4 rows or x 1 row, says the tensor map); `dep` ties the issue to data the warp has read
// out of that very slot (the slot is private to the warp, so program order is the only hand-back it needs)
__device__ __forceinline__ void tma_load_rows(void* dst, const CUtensorMap* tmap, int x, int y, int z, unsigned long long* bar, unsigned dep) {
  asm volatile(
      "{\n\t"
      ".reg .b32 t;\n\t"
      "mov.b32 t, %6;\n\t"
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n\t"
      "}" ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "r"(dep)
      : "memory");
}

// Named barriers for the "row stored" side of the LL1 hand-over (ids 1 .. kLLSlots): a waiting warp is parked by the
// hardware and spends no issue slots, unlike a polled mbarrier -- the consumers wait about half of the time by design.
__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// pair_variant() for 32 consecutive row pairs at once (lane i evaluates pair base + i), refreshed when n leaves the window:
// two or three instructions per step instead of the dozen of the closed form
struct VariantMask {
  unsigned mask = 0; int base = -0x40000000;
  __device__ __forceinline__ bool get(int n, int nW, int ls) {
    if ((n >> 5) != base) {
      base = n >> 5;
      mask = __ballot_sync(kFullMask, pair_variant((base << 5) + (int)(threadIdx.x & 31), nW, ls));
    }
    return (mask >> (n & 31)) & 1u;
  }
};

template <typename T> __device__ __forceinline__ void sts4(void* p, T a, T b, T c, T d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(bits(a)), "r"(bits(b)), "r"(bits(c)), "r"(bits(d)) : "memory");
}
__device__ __forceinline__ void lds8(const void* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void lds8(const void* p, int (&v)[8]) {
  const int4 a = *reinterpret_cast<const int4*>(p), b = *(reinterpret_cast<const int4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

#ifndef PCS_FUSED_MINB
#define PCS_FUSED_MINB 4
#endif
// tmap4: u8 frames with a 256 x 4 box (the four rows of a two-step stage in one copy); tmap1: 256 x 1 box, used row by
// row where the rows of a stage are not four consecutive rows of the frame (mirror folds at the top / bottom / padding).
template <bool LOSSY, typename OT, bool QS1>
__global__ void __launch_bounds__(kThreads, PCS_FUSED_MINB) dwt_forward_l12_kernel(const __grid_constant__ CUtensorMap tmap4,
                                                                                  const __grid_constant__ CUtensorMap tmap1, const FusedArgs a) {
  using T = typename VSel<LOSSY, 8>::T;
  using VP = typename VSel<LOSSY, 8>::type;
  __shared__ __align__(128) unsigned char s_in[kProd][kStages][4][256];
  __shared__ __align__(16) unsigned s_ll[kLLSlots][kLLRow];
  __shared__ __align__(8) unsigned long long s_full[kProd][kStages];
  __shared__ __align__(8) unsigned long long s_empty[kLLSlots];  // "the consumers have read this ring row" (count: consumer warps)

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(kFullMask, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler: TMA operands stay in uniform registers
  const int ctax = blockIdx.x % a.nctax, chunk = blockIdx.x / a.nctax, frame = blockIdx.y;
  const int Wc = a.Wc, Hc = a.Hc, hw1 = Wc >> 1, hh1 = Hc >> 1, hw2 = Wc >> 2, hh2 = Hc >> 2;
  const int X0 = ctax * kPitch;                       // first pixel column this CTA owns
  const int remLL = min(kPitchLL, hw1 - (X0 >> 1));   // LL1 columns this CTA owns
  const int ncons = remLL > 240 ? 2 : 1;
  const int nprod = min(kProd, (remLL + 20 + 119) / 120);  // valid LL1 columns needed: 12 before the first owned one, 8 after the last

  if (threadIdx.x == 0) {
    for (int p = 0; p < kProd; ++p)
      for (int d = 0; d < kStages; ++d) mbar_init(&s_full[p][d], 1);
    for (int d = 0; d < kLLSlots; ++d) mbar_init(&s_empty[d], (unsigned)ncons);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // chunk geometry: level-1 row pairs [n0, n1) and level-2 row pairs [m2_0, m2_1) are written by this CTA
  const int n0 = chunk * a.rp, n1 = min(n0 + a.rp, hh1);
  const int m2_0 = n0 >> 1, m2_1 = n1 >> 1;
  const int M0 = max(m2_0 - VP::kWarm, 0);            // first level-2 step; it starts from LL1 row 2 * M0
  const int P0 = 2 * M0;                              // first level-1 pair whose output is needed (<= n0)
  const int Plast = min(2 * (m2_1 - 1 + VP::kDelay) + 2, hh1 - 1);  // last one (level 2 reads one or three rows past n1)
  const int hand = 32 * (nprod + ncons);  // threads that meet at a hand-over barrier
  const float qs = a.qs;
  OT* out = static_cast<OT*>(a.out) + (size_t)frame * a.out_fstride;
  asm volatile("" : "+l"(out));  // keep the frame's base pointer in registers: the compiler otherwise rebuilds it from the parameters at every store

  if (warp < kProd) {
    // =============================== producer: level 1 on 256 pixel columns ===============================
    const int p = warp;
    if (p >= nprod) return;
    const int xw = X0 - 32 + 240 * p;                 // first column of the warp's 256 (a multiple of 16; may be < 0 or reach past Wc)
    const int c0 = xw + 8 * lane;
    const bool inimg = c0 >= 0 && c0 < Wc;
    const bool wlane = lane >= 1 && lane <= 30 && inimg && c0 >= X0 && c0 < X0 + kPitch;  // stores: owned columns only
    const bool left = c0 == 0, right = c0 + 8 == Wc;
    auto row_index = [&](int y) -> int {  // whole-sample fold at the borders of the level input, then the pad mapping
      int ys = y < 0 ? -y : (y >= Hc ? 2 * (Hc - 1) - y : y);
      return ys < a.H ? ys : 2 * a.H - 1 - ys;
    };
    const int ms = P0 - VP::kWarm, mend = Plast + VP::kDelay;
    const int nstages = (mend - ms + 2) >> 1;         // a stage = two steps = rows 2m+1 .. 2m+4 of its first step m
    auto issue = [&](int j, int slot, unsigned dep) {
      if (lane == 0 && j < nstages) {
        const int y0 = 2 * (ms + 2 * j) + 1;
        mbar_expect_tx(&s_full[p][slot], 1024);
        if (y0 >= 0 && y0 + 3 < a.H) {
          tma_load_rows(&s_in[p][slot][0][0], &tmap4, xw, y0, frame, &s_full[p][slot], dep);
        } else {
#pragma unroll
          for (int r = 0; r < 4; ++r) tma_load_rows(&s_in[p][slot][r][0], &tmap1, xw, row_index(y0 + r), frame, &s_full[p][slot], dep);
        }
      }
    };
#pragma unroll
    for (int d = 0; d < kStages; ++d) issue(d, d, 0u);
    VP v1;
    {
      T r[8];
      const int cl = min(max(c0, 0), Wc - 8);  // lanes outside the image read inside it; their values are never used
      unpack8(__ldg(reinterpret_cast<const uint2*>(a.src + (size_t)frame * a.src_fstride + (size_t)row_index(2 * ms) * a.src_pitch + cl)), r, a.offset);
      v1.init(r);
    }
    int ls = 0; bool lap0 = true;  // ring slot of the next LL1 row; first lap: the slots are free
    unsigned epar = 0;             // parity of the s_empty phase the next lap waits for
    const int oc1 = c0 >> 1;
    VariantMask vm;
    // stores and hand-over of row pair pr (after the horizontal pass)
    auto finish = [&](int pr, T (&lo)[8], T (&hi)[8]) {
      if (pr < P0 || pr > Plast) return;  // warm-up / the unused second half of the last stage (warp-uniform)
      if (wlane && pr >= n0 && pr < n1)
        emit_pair<LOSSY, QS1, OT, T, 8>(lo, hi, out + (size_t)pr * a.pitch + oc1, out + (size_t)(pr + hh1) * a.pitch + oc1, hw1, a.q1, qs,
                                        false, nullptr);
      // ---- hand the LL1 row to the consumers ----
      if (!lap0) mbar_wait(&s_empty[ls], epar);  // the consumers have read what the slot held a lap ago
      if (lane >= 1 && lane <= 30) sts4<T>(&s_ll[ls][120 * p + 4 * (lane - 1)], lo[0], lo[2], lo[4], lo[6]);
      bar_arrive(1 + ls, hand);
      if (++ls == kLLSlots) { ls = 0; if (!lap0) epar ^= 1u; lap0 = false; }
    };
    int slot = 0; unsigned par = 0;
#pragma unroll 1
    for (int j = 0; j < nstages; ++j) {
      mbar_wait(&s_full[p][slot], par);
      __syncwarp();  // converged: each load below is one warp instruction, and the re-arming lane cannot run ahead of a waiting one
      const unsigned char* const st = &s_in[p][slot][0][8 * lane];
      const uint2 ra = *reinterpret_cast<const uint2*>(st), rb = *reinterpret_cast<const uint2*>(st + 256);
      const uint2 rc = *reinterpret_cast<const uint2*>(st + 512), rd = *reinterpret_cast<const uint2*>(st + 768);
      issue(j + kStages, slot, ra.x ^ rb.y ^ rc.x ^ rd.y);
      if (++slot == kStages) { slot = 0; par ^= 1u; }
      // the two steps of the stage: rows 2m+1, 2m+2 (A) and 2m+3, 2m+4 (B) enter, row pairs prA and prA + 1 leave.  Both
      // vertical steps first, then the two horizontal passes as one interleaved stream (borders of this level come from
      // the folded loads, so no explicit rule in the vertical steps).
      const int prA = ms + 2 * j - VP::kDelay;
      T xo[8], x2[8], loA[8], hiA[8], loB[8], hiB[8];
      unpack8(ra, xo, a.offset); unpack8(rb, x2, a.offset);
      if constexpr (LOSSY) v97_step_mixed<8>(v1, xo, x2, loA, hiA);
      else v1.step(xo, x2, false, loA, hiA);
      unpack8(rc, xo, a.offset); unpack8(rd, x2, a.offset);
      if constexpr (LOSSY) v97_step_mixed<8>(v1, xo, x2, loB, hiB);
      else v1.step(xo, x2, false, loB, hiB);
      if constexpr (LOSSY) {
        hfwd97_rows2<8>(loA, hiA, loB, hiB, left, right, vm.get(prA, a.nW1, a.ls1), vm.get(prA + 1, a.nW1, a.ls1));
      } else {
        hfwd53<8>(loA, left, right); hfwd53<8>(hiA, left, right);
        hfwd53<8>(loB, left, right); hfwd53<8>(hiB, left, right);
      }
      finish(prA, loA, hiA);
      finish(prA + 1, loB, hiB);
    }
  } else {
    // =============================== consumer: level 2 on 256 LL1 columns ===============================
    const int c = warp - kProd;
    if (c >= ncons) return;
    const int cl0 = (X0 >> 1) - 8 + 240 * c + 8 * lane;  // first of the lane's 8 LL1 columns
    const int ridx = 4 + 240 * c + 8 * lane;             // ... and its position in a ring row (which starts at LL1 column X0/2 - 12)
    const bool inimg = cl0 >= 0 && cl0 < hw1;
    const bool wlane = lane >= 1 && lane <= 30 && inimg && cl0 >= (X0 >> 1) && cl0 < (X0 >> 1) + kPitchLL;
    const bool left = cl0 == 0, right = cl0 + 8 == hw1;
    T* const llbase = a.last ? nullptr : static_cast<T*>(a.ll) + (size_t)frame * a.ll_fstride;
    const int oc2 = cl0 >> 1;
    VariantMask vm;
    int ls = 0;  // ring slot of the next LL1 row
    auto take = [&](T (&row)[8]) {
      bar_sync(1 + ls, hand);                  // every producer has stored its part of the row
      lds8(&s_ll[ls][ridx], row);
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[ls]);  // release: the warp's loads above are ordered before it
      if (++ls == kLLSlots) ls = 0;
    };
    // horizontal pass, quantiser and stores of level-2 row pair p2
    auto finish = [&](int p2, T (&lo2)[8], T (&hi2)[8]) {
      if (p2 < m2_0 || p2 >= m2_1) return;
      if constexpr (LOSSY) {
        hfwd97_rows_rt<8>(lo2, hi2, left, right, vm.get(p2, a.nW2, a.ls2));
      } else {
        hfwd53<8>(lo2, left, right);
        hfwd53<8>(hi2, left, right);
      }
      if (wlane)
        emit_pair<LOSSY, QS1, OT, T, 8>(lo2, hi2, out + (size_t)p2 * a.pitch + oc2, out + (size_t)(p2 + hh2) * a.pitch + oc2, hw2, a.q2, qs,
                                        a.last != 0, a.last ? nullptr : llbase + (size_t)p2 * a.ll_pitch + oc2);
    };
    VP v2;
    {
      T r[8];
      take(r);
      v2.init(r);
    }
    // steps whose two rows 2M+1, 2M+2 both exist; M < 2 only at the top of the image (explicit border rules)
    const int Mreg = (Plast - 2) >> 1;
#pragma unroll 2
    for (int M = M0; M <= Mreg; ++M) {
      T xo[8], x2[8], lo2[8], hi2[8];
      take(xo); take(x2);
      if constexpr (LOSSY) {
        if (M >= 2) v97_step_mixed<8>(v2, xo, x2, lo2, hi2);
        else v2.step(xo, x2, M == 0, M == 1, false, lo2, hi2);
      } else {
        v2.step(xo, x2, M == 0, lo2, hi2);
      }
      finish(M - VP::kDelay, lo2, hi2);
    }
    if (Plast & 1) {
      // bottom of the image (Plast == hh1 - 1): the last odd row has no even partner, x[n] := x[n-2]; 9/7: one flush step more
      T xo[8], x2[8], lo2[8], hi2[8];
      take(xo);
#pragma unroll
      for (int j = 0; j < 8; ++j) x2[j] = v2.xe[j];
      for (int M = hh2 - 1; M <= hh2 - 1 + VP::kDelay; ++M) {
        if constexpr (LOSSY) v2.step(xo, x2, M == 0, M == 1, M == hh2, lo2, hi2);
        else v2.step(xo, x2, M == 0, lo2, hi2);
        finish(M - VP::kDelay, lo2, hi2);
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

template <bool LOSSY, typename OT, bool QS1>
void launch_fused_k(const CUtensorMap& tmap4, const CUtensorMap& tmap1, FusedArgs a, int frames, cudaStream_t st) {
  static const long long slots = [] {  // CTAs the whole GPU holds at once
    int dev = 0, sms = 148, per_sm = PCS_FUSED_MINB;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dwt_forward_l12_kernel<LOSSY, OT, QS1>, kThreads, 0);
    return (long long)sms * std::max(per_sm, 1);
  }();
  // Row pairs per chunk: a chunk costs (rp + warm-up) serial steps and the grid runs in ceil(CTAs / resident CTAs) waves
  const int hh1 = a.Hc >> 1, warm = LOSSY ? 10 : 4;
  static const int rp_env = [] { const char* e = std::getenv("PCS_FUSED_RP"); return e ? std::atoi(e) : 0; }();
  int best = 4; long long best_cost = -1;
  for (int rp = 4; rp <= 512; rp += 4) {
    const long long ctas = (long long)a.nctax * ((hh1 + rp - 1) / rp) * frames;
    const long long cost = ((ctas + slots - 1) / slots) * (rp + warm);
    if (best_cost < 0 || cost <= best_cost) { best = rp; best_cost = cost; }
  }
  a.rp = rp_env > 0 ? std::max(4, rp_env & ~3) : best;
  a.nchunks = (hh1 + a.rp - 1) / a.rp;
  dim3 grid(a.nctax * a.nchunks, frames);
  dwt_forward_l12_kernel<LOSSY, OT, QS1><<<grid, kThreads, 0, st>>>(tmap4, tmap1, a);
}

}  // namespace

// Can levels 1+2 of this geometry go through the fused kernel?  (u8 frames without horizontal padding, wl >= 2)
bool dwt_forward_l12_ok(const void* d_px, const void* d_coeff, const void* ll2, int W, int H, int Wa, int Ha, int wl) {
  if (wl < 2 || W != Wa || H > Ha || 2 * H < Ha || Wa % 64 || Ha % 64 || Wa < 64 || Ha < 64) return false;
  if ((reinterpret_cast<uintptr_t>(d_px) & 15) || (reinterpret_cast<uintptr_t>(d_coeff) & 15)) return false;
  if (ll2 && (reinterpret_cast<uintptr_t>(ll2) & 15)) return false;
  if (((long long)W * H) % 16) return false;
  return encode_tiled() != nullptr;
}

// Levels 1 and 2 of `frames` u8 frames in one launch.  ll2: scratch plane that receives LL2 (ignored when wl == 2).
int dwt_forward_l12(const uint8_t* d_px, int W, int H, int Wa, int Ha, void* d_coeff, long long out_fstride, void* ll2, long long ll_fstride,
                    int wl, bool lossy, float qs, bool out_int, int frames, cudaStream_t st, int offset) {
  alignas(64) CUtensorMap tmap4, tmap1;
  const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)frames};
  const cuuint64_t gstride[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
  const cuuint32_t box4[3] = {256, 4, 1}, box1[3] = {256, 1, 1}, estr[3] = {1, 1, 1};
  for (int k = 0; k < 2; ++k) {
    const CUresult r = encode_tiled()(k ? &tmap1 : &tmap4, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(d_px), gdim, gstride,
                                      k ? box1 : box4, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
  }
  FusedArgs a{};
  a.src = d_px; a.src_fstride = (long long)W * H; a.src_pitch = W;
  a.H = H; a.Wc = Wa; a.Hc = Ha; a.offset = offset;
  a.out = d_coeff; a.out_fstride = out_fstride; a.pitch = Wa;
  a.ll = ll2; a.ll_fstride = ll_fstride; a.ll_pitch = Wa >> 2;
  a.last = wl == 2;
  for (int i = 0; i < 4; ++i) { a.q1[i] = kQ[0][i]; a.q2[i] = kQ[1][i]; }
  a.qs = qs;
  variant_params(Ha, &a.nW1, &a.ls1);
  variant_params(Ha >> 1, &a.nW2, &a.ls2);
  a.nctax = (Wa + kPitch - 1) / kPitch;
  const bool qs1 = qs == 1.0f;
  if (!lossy) launch_fused_k<false, int, true>(tmap4, tmap1, a, frames, st);
  else if (out_int) { if (qs1) launch_fused_k<true, int, true>(tmap4, tmap1, a, frames, st); else launch_fused_k<true, int, false>(tmap4, tmap1, a, frames, st); }
  else { if (qs1) launch_fused_k<true, float, true>(tmap4, tmap1, a, frames, st); else launch_fused_k<true, float, false>(tmap4, tmap1, a, frames, st); }
  ++g_launches;
  return (int)cudaGetLastError();
}

}  // namespace pcs
