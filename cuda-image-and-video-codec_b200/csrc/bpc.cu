// bpc.cu -- BPC-PaCo bit-plane coder / decoder for sm_100a (2 coding passes, k = 0).
//
// Replaces kernelBPCCoder<T> / kernelBPCDecoder<T> and everything under them
// (BPC/BPCEngine.cu:41-63,94-111,143-192,222-308,329-358,371-516,558-594,726-843,930-1003,
// 1249-1279,1668-1721,1776-1837,1904-2026,2125-2215 of the reference).  The codestream is
// bit-identical: same visit order, same contexts, same per-lane fixed-length arithmetic coders,
// same codeword slot reservation order (SURVEY.md Annex B, pinned in oracle/picsong_oracle.c).
//
// What is different from the reference kernel (one warp per 64x64 codeblock, lane = 2 columns,
// 128 state words per thread in LOCAL memory, three shuffles + a global LUT load per visited sample):
//   * state is BIT-SLICED: per lane and column a 64-bit row mask for "significant", "sign",
//     "bit of the current plane"; no per-coefficient state words, no local memory;
//   * the ENCODER knows all magnitudes, so the significance state a neighbour has *at the moment it
//     is looked at* is a pure function of the data (visited earlier in the pass -> |q| >= 2^bp, else
//     |q| >= 2^(bp+1)).  All 4096 significance contexts and sign contexts of a plane are therefore
//     computed at once with carry-save adders on the row masks (a few dozen LOP3 per plane); the
//     serial loop only runs the interval coder.  No shuffles in the encoder's inner loop;
//   * the DECODER splits each context into a static part (state before the pass, bit-sliced) and the
//     few neighbours decoded earlier in the same pass, which travel as 1-2 bit flags (two shuffles
//     per row instead of six);
//   * probabilities of a (subband, plane) row are packed 9 x 7 bit into one 64-bit register per lane
//     (4 x 7 bit for the sign) -- no table loads in the inner loop; the class is per LANE, as in the
//     reference (findSubband is called with the thread's x, BPCEngine.cu:1975,1990);
//   * codeword slot reservation is an explicit __ballot_sync over "lanes of this call whose interval
//     is exhausted" instead of __activemask() inside divergent code;
//   * codewords are written as uint16 (the reference writes int32 and narrows later).
#include <algorithm>

#include "pcs_common.cuh"

namespace pcs {

namespace {

typedef unsigned long long u64;
constexpr unsigned kFull = 0xffffffffu;

template <typename W> __device__ __forceinline__ void add3(W a, W b, W c, W& s, W& cy) {
  s = a ^ b ^ c;
  cy = (a & b) | (c & (a ^ b));
}
// bit-sliced population count of 8 one-bit row masks -> 4 bit planes (value 0..8 per row); W = u64 (64 rows) or unsigned (32)
template <typename W> __device__ __forceinline__ void count8(W m0, W m1, W m2, W m3, W m4, W m5, W m6, W m7, W (&k)[4]) {
  W s1, c1, s2, c2, b0, c4, s5, c5;
  add3(m0, m1, m2, s1, c1);
  add3(m3, m4, m5, s2, c2);
  const W s3 = m6 ^ m7, c3 = m6 & m7;
  add3(s1, s2, s3, b0, c4);
  add3(c1, c2, c3, s5, c5);
  const W c6 = s5 & c4;
  k[0] = b0; k[1] = s5 ^ c4; k[2] = c5 ^ c6; k[3] = c5 & c6;
}
// rows [sh, sh + 32) of a 64-row mask; of the mask moved one row down (row i holds x[i-1]) / one row up (x[i+1])
__device__ __forceinline__ unsigned rows32(u64 x, int sh) { return (unsigned)(x >> sh); }
__device__ __forceinline__ unsigned rows32_prev(u64 x, int sh) { return (unsigned)((x << 1) >> sh); }
__device__ __forceinline__ unsigned rows32_next(u64 x, int sh) { return (unsigned)((x >> 1) >> sh); }

__device__ __forceinline__ u64 shfl_up64(u64 v, int lane) {
  unsigned lo = __shfl_up_sync(kFull, (unsigned)v, 1), hi = __shfl_up_sync(kFull, (unsigned)(v >> 32), 1);
  return lane == 0 ? 0ull : ((u64)hi << 32 | lo);  // outside the codeblock counts as 0 (BPCEngine.cu:465-484)
}
__device__ __forceinline__ u64 shfl_down64(u64 v, int lane) {
  unsigned lo = __shfl_down_sync(kFull, (unsigned)v, 1), hi = __shfl_down_sync(kFull, (unsigned)(v >> 32), 1);
  return lane == 31 ? 0ull : ((u64)hi << 32 | lo);
}

// BPCEngine.cu:143-170 findSubband with the THREAD's x: returns the LUT row group lvl*nSub + sb.
__device__ __forceinline__ int lane_group(int X, int Y, const BpcArgs& p) {
  for (int a = 1; a <= p.wl; ++a) {
    const bool bx = X >= (p.Wa >> a), by = Y >= (p.Ha >> a);
    if (bx || by) return (a - 1) * p.nSub + (bx ? (by ? 2 : 0) : 1);
  }
  return p.wl * p.nSub;
}

template <typename T> __device__ __forceinline__ void load2(const T* p, int& a, int& b);
template <> __device__ __forceinline__ void load2<int>(const int* p, int& a, int& b) {
  const int2 v = __ldg(reinterpret_cast<const int2*>(p)); a = v.x; b = v.y;
}
template <> __device__ __forceinline__ void load2<float>(const float* p, int& a, int& b) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  a = __float2int_rz(v.x); b = __float2int_rz(v.y);  // (int)c, BPCEngine.cu:49,55
}
__device__ __forceinline__ unsigned umag(int v) { return v < 0 ? 0u - (unsigned)v : (unsigned)v; }

// ---------------------------------------------------------------------------------------------
// encoder
// ---------------------------------------------------------------------------------------------

// In-register 32x32 bit-matrix transpose: on return w[k] bit i == (old w[i]) bit k.  5 stages of masked
// swaps, fully unrolled (all indices compile-time).
__device__ __forceinline__ void transpose32(unsigned (&w)[32]) {
  unsigned m = 0x0000FFFFu;
#pragma unroll
  for (int j = 16; j != 0; j >>= 1) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k & j) continue;
      const unsigned t = ((w[k] >> j) ^ w[k + j]) & m;
      w[k] ^= t << j;
      w[k + j] ^= t;
    }
    m ^= m << (j >> 1);
  }
}

// 4 bit planes (32 rows each) -> 4 words of 8 nibbles: afterwards nibble j of n[q] holds, for row 4j+q,
// the value plane0 | plane1<<1 | plane2<<2 | plane3<<3.
__device__ __forceinline__ void nibble_net(unsigned (&n)[4]) {
  unsigned t;
  t = ((n[0] >> 1) ^ n[1]) & 0x55555555u; n[1] ^= t; n[0] ^= t << 1;
  t = ((n[2] >> 1) ^ n[3]) & 0x55555555u; n[3] ^= t; n[2] ^= t << 1;
  t = ((n[0] >> 2) ^ n[2]) & 0x33333333u; n[2] ^= t; n[0] ^= t << 2;
  t = ((n[1] >> 2) ^ n[3]) & 0x33333333u; n[3] ^= t; n[1] ^= t << 2;
}
// same for 2 planes (value 0..3 per nibble)
__device__ __forceinline__ void nibble_net2(unsigned p0, unsigned p1, unsigned (&n)[4]) {
  const unsigned t = ((p0 >> 1) ^ p1) & 0x55555555u;
  const unsigned e = p0 ^ (t << 1), o = p1 ^ t;  // bit pairs of even / odd rows
  n[0] = e & 0x33333333u; n[2] = (e >> 2) & 0x33333333u;
  n[1] = o & 0x33333333u; n[3] = (o >> 2) & 0x33333333u;
}

// Four table look-ups at once: byte i of the result = table entry selected by nibble i of `sel`
// (entries 0..7 = bytes of lo|hi; bit 3 of a nibble is PRMT's sign-replication flag and is masked off, so
// context 8 reads entry 0 and is patched by fix8).
__device__ __forceinline__ unsigned lut4(unsigned lo, unsigned hi, unsigned sel) { return __byte_perm(lo, hi, sel & 0x7777u); }
// rows whose context is 8 looked up entry 0 above; replace it: delta = entry 8 - entry 0 (no inter-byte carry)
__device__ __forceinline__ unsigned fix8(unsigned probs, unsigned sel, unsigned delta) {
  const unsigned m = (sel >> 3) & 0x1111u;
  const unsigned spread = (m & 1u) | ((m & 0x10u) << 4) | ((m & 0x100u) << 8) | ((m & 0x1000u) << 12);
  return probs + spread * delta;
}

__device__ __forceinline__ unsigned byte_of(unsigned x, int k) { return __byte_perm(x, 0u, 0x4440u + (unsigned)k); }

// Per-lane interval coder, branch-free step (BPCEngine.cu:371-399).  `slot` is relative to out+1.
struct EncState { unsigned lower, size, slot, count, lt; };
// One call site of the interval coder for the whole warp.  `abw` / `anbw`: words whose bit `bitm` says
// "this lane codes a 1 here" / "codes a 0 here" (neither: the lane is idle at this call site).
// A completed codeword (size == 0) is not stored when it completes but when its lane starts the next one (or by the flush
// at the end of the block, which stores every lane's current word): the store then sits with the slot hand-out inside the
// part of the step that is skipped, with a warp-uniform branch, at the call sites where no lane starts a codeword.  The
// lane's first "completed word" is the initial state (lower = 0): its slot is WORDS - 2, which no codeword of a block that
// is not expanded can own (slots < count <= WORDS - 2; an expanded block is rewritten raw).
// Hand-scheduled as predicated PTX: nvcc's lowering of the equivalent C++ (kept below) costs ~50 % more
// instructions on the ALU pipe, which is what bounds this kernel.
#ifndef PCS_ENC_PTX
#define PCS_ENC_PTX 1
#endif
template <int WORDS>  // words a codeblock owns (64 x rows): slots are clamped to WORDS - 2 (BPCEngine.cu:382 for 64 rows)
__device__ __forceinline__ void enc_step(unsigned abw, unsigned anbw, unsigned bitm, unsigned p, EncState& c,
                                         unsigned short* out1, unsigned prec) {
#if PCS_ENC_PTX
  asm volatile(
      "{\n\t"
      ".reg .pred pab, panb, pact, pneed, pz;\n\t"
      ".reg .b32 t, m, x, pre, pm, r, sl;\n\t"
      ".reg .b64 ad;\n\t"
      "and.b32 t, %4, %6;\n\t"
      "setp.ne.u32 pab, t, 0;\n\t"
      "and.b32 t, %5, %6;\n\t"
      "setp.ne.u32 panb, t, 0;\n\t"
      "or.pred pact, pab, panb;\n\t"
      "setp.eq.and.u32 pneed, %1, 0, pact;\n\t"       // lanes that start a codeword at this call site,
      "vote.sync.any.pred pz, pneed, 0xffffffff;\n\t"
      "@!pz bra.uni LAB_ENC_NONEED;\n\t"               // warp-uniform: nobody starts a codeword here (about every other site)
      "vote.sync.ballot.b32 m, pneed, 0xffffffff;\n\t"  // served in lane order (BPCEngine.cu:376-384)
      "min.u32 sl, %2, %11;\n\t"
      "mad.wide.u32 ad, sl, 2, %10;\n\t"
      "@pneed st.global.u16 [ad], %0;\n\t"             // the codeword the lane has just completed (deferred store)
      "and.b32 x, m, %7;\n\t"
      "popc.b32 pre, x;\n\t"
      "popc.b32 pm, m;\n\t"
      "@pneed add.u32 %2, %3, pre;\n\t"
      "add.u32 %3, %3, pm;\n\t"
      "@pneed mov.u32 %1, 0xFFFF;\n\t"
      "@pneed mov.u32 %0, 0;\n\t"
      "LAB_ENC_NONEED:\n\t"
      "mad.lo.u32 r, %1, %8, %12;\n\t"                 // size * p + 2^prec ...
      "shr.u32 r, r, %9;\n\t"                          // ... >> prec = q + 1
      "@pab sub.u32 %1, %1, r;\n\t"
      "@pab add.u32 %0, %0, r;\n\t"
      "@panb add.u32 %1, r, -1;\n\t"
      "}"
      : "+r"(c.lower), "+r"(c.size), "+r"(c.slot), "+r"(c.count)
      : "r"(abw), "r"(anbw), "r"(bitm), "r"(c.lt), "r"(p), "r"(prec), "l"(out1), "n"(WORDS - 2), "r"(1u << prec)
      : "memory");
#else
  const bool ab = (abw & bitm) != 0, anb = (anbw & bitm) != 0;
  const bool act = ab || anb;
  const bool need = act && (c.size == 0);
  const unsigned m = __ballot_sync(kFull, need);
  if (m != 0) {
    if (need) out1[min(c.slot, (unsigned)(WORDS - 2))] = (unsigned short)c.lower;
    const unsigned pre = __popc(m & c.lt);
    if (need) { c.slot = c.count + pre; c.size = 0xFFFFu; c.lower = 0u; }
    c.count += __popc(m);
  }
  const unsigned q = (c.size * p) >> prec;
  if (ab) { c.size = c.size - q - 1u; c.lower = c.lower + q + 1u; }
  if (anb) c.size = q;
#endif
}

#ifndef PCS_SPP_UNROLL
#define PCS_SPP_UNROLL 1
#endif
constexpr int kSppUnroll = PCS_SPP_UNROLL;

// CTAs per SM the encoder is register-bounded for (32 KB of shared memory per CTA allow 6).  The coder step is short and
// latency-bound, so resident warps pay: 3.48 ms per 32 4K frames at 4 CTAs (128 registers), 3.15 at 5 (96), 2.98 at 6 (80
// registers, 56 bytes of spill; profiles/r2w_bpc_occupancy.json, spill traffic in profiles/r2x_bpc_metrics.md).
#ifndef PCS_ENC_MINB
#define PCS_ENC_MINB 6
#endif
template <typename T, int ROWS>  // ROWS: codeblock length, 64 or 32 (the reference's CBLOCK_LENGTH, BPC/BPCEngine.cuh:33-36)
__global__ void __launch_bounds__(128, PCS_ENC_MINB) bpc_encode_kernel(const T* __restrict__ coeff, unsigned short* __restrict__ cw,
                                                        int* __restrict__ sizes, const BpcArgs p) {
  // transposed magnitude bits of the warp's codeblock: word [32*half + k][lane], k < 16: plane k of the
  // lane's column 0, k >= 16: plane k-16 of column 1; bit i = row 32*half + i
  __shared__ unsigned s_pl[4][64][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int w = blockIdx.x * (blockDim.x >> 5) + wib;
  if (w >= p.total_cb) return;
  const int frame = (w / p.ncb) * p.pstep + p.pfirst, cb = w % p.ncb;
  constexpr int kHalves = ROWS / 32, kWords = kCB * ROWS;  // 32-row halves of a block; words a block owns
  const int bx = cb % p.nbx, by = cb / p.nbx;
  const T* src = coeff + (size_t)frame * p.A + (size_t)(by * ROWS) * p.Wa + bx * kCB + 2 * lane;
  const size_t blk = (size_t)frame * p.ncb + cb;  // index of this block in the codeword / size arrays
  unsigned short* out = cw + blk * kWords;
  unsigned (*pl)[32] = s_pl[wib];

  // pass 0: MSB of the block, sign masks (BPCEngine.cu:41-63,176-192), bit-plane transposition
  u64 g0 = 0, g1 = 0;
  unsigned ormag = 0;
#pragma unroll 1
  for (int h = 0; h < kHalves; ++h) {
    unsigned wv[32], sg0 = 0, sg1 = 0;
#pragma unroll
    for (int i = 31; i >= 0; --i) {
      int a, b; load2<T>(src + (size_t)(32 * h + i) * p.Wa, a, b);
      const unsigned ma = umag(a), mb = umag(b);
      ormag |= ma | mb;
      sg0 = __funnelshift_l((unsigned)a, sg0, 1);  // (sg0 << 1) | sign(a): row i ends up at bit i
      sg1 = __funnelshift_l((unsigned)b, sg1, 1);
      wv[i] = (ma & 0xFFFFu) | (mb << 16);
    }
    transpose32(wv);
#pragma unroll
    for (int k = 0; k < 32; ++k) pl[32 * h + k][lane] = wv[k];
    g0 |= (u64)sg0 << (32 * h); g1 |= (u64)sg1 << (32 * h);
  }
  ormag = __reduce_or_sync(kFull, ormag);
  const int MSB = ormag ? 31 - __clz(ormag) : 32;
  if (MSB == 32) {  // nothing is coded; sizeArray = 1 (BPCEngine.cu:1998-2010)
    if (lane == 0) { out[0] = 32; sizes[blk] = 1; }
    return;
  }
  __syncwarp();
  int row = lane_group(bx * kCB + 2 * lane, by * ROWS, p) * p.nBp + MSB;  // initializeLUTPointers :329-350
  const u64 Lg = shfl_up64(g1, lane), Rg = shfl_down64(g0, lane);       // neighbour lanes' signs
  unsigned short* const out1 = out + 1;
  const unsigned prec = (unsigned)p.prec;

  EncState ac; ac.lower = 0; ac.size = 0; ac.slot = kWords - 2;  /* dump slot for the deferred store of the initial state, see enc_step */ ac.count = 0; ac.lt = (1u << lane) - 1u;
  u64 s0 = 0, s1 = 0;  // significant BEFORE the current plane
  for (int bp = MSB; bp >= 0; --bp, --row) {
    u64 B0, B1;
    if (bp < 16) {
      B0 = (u64)pl[bp][lane]; B1 = (u64)pl[16 + bp][lane];
      if constexpr (kHalves == 2) { B0 |= (u64)pl[32 + bp][lane] << 32; B1 |= (u64)pl[48 + bp][lane] << 32; }
    } else {  // |q| >= 2^16: not reachable from 8-bit images with the shipped quantiser; generic slow path
      B0 = 0; B1 = 0;
#pragma unroll 4
      for (int i = 0; i < ROWS; ++i) {
        int a, b; load2<T>(src + (size_t)i * p.Wa, a, b);
        B0 |= (u64)((umag(a) >> bp) & 1u) << i; B1 |= (u64)((umag(b) >> bp) & 1u) << i;
      }
    }
    const uint4 tab = (row >= 0 && row < p.nrows) ? __ldg(p.tab4 + row) : make_uint4(0, 0, 0, 0);
    const unsigned d8 = (tab.z & 0xFFu) - (tab.x & 0xFFu), pref = (tab.z >> 8) & 0xFFu;
    const u64 a0 = s0 | B0, a1 = s1 | B1;  // significant AFTER this plane's SPP visit
    const u64 Lb = shfl_up64(s1, lane), La = shfl_up64(a1, lane);
    const u64 Rb = shfl_down64(s0, lane), Ra = shfl_down64(a0, lane);
    // ---- significance propagation pass: rows in order, column 0 then column 1 (BPCEngine.cu:770-843) ----
    // The context words of a 32-row half are built inside the loop, on 32-bit slices of the 64-row state words, so that only
    // those state words stay live across the coding loop (the 64-row versions cost 28 more registers for the same work).
#pragma unroll 1
    for (int h = 0; h < kHalves; ++h) {
      const int sh = 32 * h;
      const unsigned S0 = rows32(s0, sh), S1 = rows32(s1, sh);
      const unsigned A0 = ~S0, A1 = ~S1;
      const unsigned anyact = __reduce_or_sync(kFull, A0 | A1);  // rows of this half in which some lane codes a symbol
      if (anyact == 0) continue;
      const unsigned b0 = rows32(B0, sh), b1 = rows32(B1, sh), G0 = rows32(g0, sh), G1 = rows32(g1, sh);
      const unsigned a0s = S0 | b0, Lbs = rows32(Lb, sh), Ras = rows32(Ra, sh);
      // significance contexts (BPCEngine.cu:222-230 with the visit-order rule of :770-843):
      // rows above and same-row samples visited earlier use the "after" state, everything else "before".
      unsigned n0[4], n1[4];
      count8(rows32_prev(La, sh), rows32_prev(a0, sh), rows32_prev(a1, sh), Lbs, S1, rows32_next(Lb, sh), rows32_next(s0, sh),
             rows32_next(s1, sh), n0);
      count8(rows32_prev(a0, sh), rows32_prev(a1, sh), rows32_prev(Ra, sh), a0s, Ras, rows32_next(s0, sh), rows32_next(s1, sh),
             rows32_next(Rb, sh), n1);
      // sign contexts (BPCEngine.cu:252-308) bit-sliced: pos/neg contribution of L,R (h) and U,D (v)
      unsigned y0, I00, I01, y1, I10, I11;
      {
        const unsigned Lgs = rows32(Lg, sh);
        const unsigned Lp = Lbs & ~Lgs, Ln = Lbs & Lgs, Rp = S1 & ~G1, Rn = S1 & G1;
        const unsigned Us = rows32_prev(a0, sh), Ug = rows32_prev(g0, sh), Ds = rows32_next(s0, sh), Dg = rows32_next(g0, sh);
        const unsigned Up = Us & ~Ug, Un = Us & Ug, Dp = Ds & ~Dg, Dn = Ds & Dg;
        const unsigned hp = (Lp & ~Rn) | (Rp & ~Ln), hn = (Ln & ~Rp) | (Rn & ~Lp);
        const unsigned vp = (Up & ~Dn) | (Dp & ~Un), vn = (Un & ~Dp) | (Dn & ~Up);
        const unsigned both = (hp & vp) | (hn & vn);
        y0 = G0 ^ (hn | (~hp & vn));
        I00 = (~(hp | hn) & (vp | vn)) | both;
        I01 = ((hp | hn) & ~(vp | vn)) | both;
      }
      {
        const unsigned Rgs = rows32(Rg, sh);
        const unsigned Lp = a0s & ~G0, Ln = a0s & G0, Rp = Ras & ~Rgs, Rn = Ras & Rgs;
        const unsigned Us = rows32_prev(a1, sh), Ug = rows32_prev(g1, sh), Ds = rows32_next(s1, sh), Dg = rows32_next(g1, sh);
        const unsigned Up = Us & ~Ug, Un = Us & Ug, Dp = Ds & ~Dg, Dn = Ds & Dg;
        const unsigned hp = (Lp & ~Rn) | (Rp & ~Ln), hn = (Ln & ~Rp) | (Rn & ~Lp);
        const unsigned vp = (Up & ~Dn) | (Dp & ~Un), vn = (Un & ~Dp) | (Dn & ~Up);
        const unsigned both = (hp & vp) | (hn & vn);
        y1 = G1 ^ (hn | (~hp & vn));
        I10 = (~(hp | hn) & (vp | vn)) | both;
        I11 = ((hp | hn) & ~(vp | vn)) | both;
      }
      const bool has8 = (n0[3] | n1[3]) != 0;  // some row of this lane has all 8 neighbours significant
      nibble_net(n0); nibble_net(n1);
      unsigned i0[4], i1[4];
      nibble_net2(I00, I01, i0);
      nibble_net2(I10, I11, i1);
#pragma unroll 1
      for (int jj = 0; jj < 2; ++jj) {
        // probabilities of rows 16*jj + 4*t + q: byte t of P?[q] (significance) and S?[q] (sign)
        unsigned P0[4], P1[4], S0[4], S1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned e0 = n0[q] >> (16 * jj), e1 = n1[q] >> (16 * jj);
          P0[q] = lut4(tab.x, tab.y, e0); P1[q] = lut4(tab.x, tab.y, e1);
          S0[q] = lut4(tab.w, 0u, i0[q] >> (16 * jj)); S1[q] = lut4(tab.w, 0u, i1[q] >> (16 * jj));
        }
        if (has8) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            P0[q] = fix8(P0[q], n0[q] >> (16 * jj), d8); P1[q] = fix8(P1[q], n1[q] >> (16 * jj), d8);
          }
        }
        // this 16-row slab's symbol masks: lane codes a 1 / a 0; sign symbol 1 / 0 of samples that become significant
        const unsigned sl = 16 * jj;
        const unsigned ab0 = (A0 & b0) >> sl, an0 = (A0 & ~b0) >> sl, ab1 = (A1 & b1) >> sl, an1 = (A1 & ~b1) >> sl;
        const unsigned gp0 = ab0 & (y0 >> sl), gn0 = ab0 & ~(y0 >> sl), gp1 = ab1 & (y1 >> sl), gn1 = ab1 & ~(y1 >> sl);
        const unsigned anyj = anyact >> sl;
        const unsigned new0 = __reduce_or_sync(kFull, ab0), new1 = __reduce_or_sync(kFull, ab1);  // warp-uniform
#pragma unroll(kSppUnroll)
        for (int t = 0; t < 4; ++t) {
          if (((anyj >> (4 * t)) & 15u) == 0) continue;  // warp-uniform
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const unsigned bitm = 1u << (4 * t + q);
            enc_step<kWords>(ab0, an0, bitm, byte_of(P0[q], t), ac, out1, prec);
            if (new0 & bitm) enc_step<kWords>(gp0, gn0, bitm, byte_of(S0[q], t), ac, out1, prec);
            enc_step<kWords>(ab1, an1, bitm, byte_of(P1[q], t), ac, out1, prec);
            if (new1 & bitm) enc_step<kWords>(gp1, gn1, bitm, byte_of(S1[q], t), ac, out1, prec);
          }
        }
      }
    }
    // ---- magnitude refinement pass: samples significant before this plane (BPCEngine.cu:726-736,1249-1261);
    //      only rows in which some lane refines are visited ----
#pragma unroll 1
    for (int h = 0; h < kHalves; ++h) {
      const int sh = 32 * h;
      const unsigned r0m = (unsigned)(s0 >> sh), r1m = (unsigned)(s1 >> sh), b0 = (unsigned)(B0 >> sh), b1 = (unsigned)(B1 >> sh);
      const unsigned r0b = r0m & b0, r0n = r0m & ~b0, r1b = r1m & b1, r1n = r1m & ~b1;
      unsigned U = __reduce_or_sync(kFull, r0m | r1m);
      while (U) {
        const unsigned bitm = U & (0u - U);  // lowest row in which some lane refines
        U ^= bitm;
        enc_step<kWords>(r0b, r0n, bitm, pref, ac, out1, prec);
        enc_step<kWords>(r1b, r1n, bitm, pref, ac, out1, prec);
      }
    }
    s0 = a0; s1 = a1;
  }
  out1[min(ac.slot, (unsigned)(kWords - 2))] = (unsigned short)ac.lower;  // flush (BPCEngine.cu:1719)
  const int size = (int)min(ac.count, (unsigned)(kWords - 1)) + 1;         // :2010
  __syncwarp();
  if (size == kWords) {
    // expansion fallback (BPCEngine.cu:2015-2024,1904-1912): raw low 16 bits of (|q|<<1 | sign), lane-major
    __threadfence_block();
    for (int i = 0; i < ROWS; ++i) {
      int a, b; load2<T>(src + (size_t)i * p.Wa, a, b);
      out[lane * 2 * ROWS + 2 * i] = (unsigned short)(((umag(a) << 1) | (a < 0 ? 1u : 0u)) & 0xFFFFu);
      out[lane * 2 * ROWS + 2 * i + 1] = (unsigned short)(((umag(b) << 1) | (b < 0 ? 1u : 0u)) & 0xFFFFu);
    }
  } else if (lane == 0) {
    out[0] = (unsigned short)MSB;
  }
  if (lane == 0) sizes[blk] = size;
}

// ---------------------------------------------------------------------------------------------
// decoder
// ---------------------------------------------------------------------------------------------
// sign context table (BPCEngine.cu:252-293) indexed by (sgn(h)+1)*3 + (sgn(v)+1), 3 bits each:
// bit 0 = "flip the decoded symbol", bits 2:1 = probability index
constexpr unsigned kSignTab = 7u | (5u << 3) | (1u << 6) | (3u << 9) | (0u << 12) | (2u << 15) | (0u << 18) | (4u << 21) | (6u << 24);
__device__ __forceinline__ unsigned sign_ctx(int h, int v) {
  const int sh = (h > 0) - (h < 0), sv = (v > 0) - (v < 0);
  return (kSignTab >> (3 * ((sh + 1) * 3 + (sv + 1)))) & 7u;
}
__device__ __forceinline__ int contrib(unsigned sig, unsigned neg) { return sig ? (neg ? -1 : 1) : 0; }
// Entries of the decoder's sign-context byte tables (value: bit 0 = flip, bits 2:1 = probability index).
// Column 0, index bits: [1:0] static horizontal sum as (non-zero, negative), [3:2] down, [5:4] up.
__device__ __forceinline__ unsigned sign_entry0(unsigned i) {
  const int h = contrib(i & 1u, (i >> 1) & 1u);
  const int v = contrib((i >> 4) & 1u, (i >> 5) & 1u) + contrib((i >> 2) & 1u, (i >> 3) & 1u);
  return sign_ctx(h, v);
}
// Column 1, index bits: [1:0] down, [3:2] left, [5:4] right, [7:6] up (each: significant, negative).
__device__ __forceinline__ unsigned sign_entry1(unsigned i) {
  const int h = contrib((i >> 2) & 1u, (i >> 3) & 1u) + contrib((i >> 4) & 1u, (i >> 5) & 1u);
  const int v = contrib((i >> 6) & 1u, (i >> 7) & 1u) + contrib(i & 1u, (i >> 1) & 1u);
  return sign_ctx(h, v);
}

struct DecState { unsigned lower, size, slot, count, code, lt; };
// One call site of the interval decoder for the whole warp (BPCEngine.cu:405-442); returns the decoded
// bit (0 for idle lanes).  `actw` bit `bitm`: this lane decodes a symbol here.  Predicated PTX, see enc_step.
// PREC = 7 (every shipped LUT): the shift is an immediate; PREC = 0: taken from `prec` at run time.
// PCS_DEC_SKIP_REFILL (default 1): no lane starts a codeword at about every other call site (ncu source counters on the
// 4K video workload: 48 % of the significance sites refill); the eleven instructions
// that hand out stream slots and fetch the words are then jumped over with a warp-uniform branch (the ballot is the
// same in every lane).
#ifndef PCS_DEC_SKIP_REFILL
#define PCS_DEC_SKIP_REFILL 1
#endif
#if PCS_DEC_SKIP_REFILL
#define PCS_DEC_SKIP_OPEN "vote.sync.any.pred pz, pneed, 0xffffffff;\n\t@!pz bra.uni LAB_NOREFILL;\n\tvote.sync.ballot.b32 m, pneed, 0xffffffff;\n\t"
#define PCS_DEC_SKIP_CLOSE "LAB_NOREFILL:\n\t"
#else
#define PCS_DEC_SKIP_OPEN "vote.sync.ballot.b32 m, pneed, 0xffffffff;\n\t"
#define PCS_DEC_SKIP_CLOSE
#endif
// MULSHR: r = size * p + 2^prec; t = r >> prec  (= q + 1: the "+ 1" of the interval split rides on the multiply-add)
#define PCS_DEC_STEP_ASM(MULSHR)                                \
      "{\n\t"                                                   \
      ".reg .pred pact, pneed, pb, pnb, pz;\n\t"                 \
      ".reg .b32 t, m, x, pre, pm, r, sl, lim;\n\t"              \
      ".reg .b64 ad;\n\t"                                        \
      "and.b32 t, %6, %7;\n\t"                                   \
      "setp.ne.u32 pact, t, 0;\n\t"                              \
      "setp.eq.and.u32 pneed, %1, 0, pact;\n\t"                  \
      PCS_DEC_SKIP_OPEN                                          \
      "and.b32 x, m, %8;\n\t"                                    \
      "popc.b32 pre, x;\n\t"                                     \
      "popc.b32 pm, m;\n\t"                                      \
      "@pneed add.u32 %2, %3, pre;\n\t"                          \
      "add.u32 %3, %3, pm;\n\t"                                  \
      "min.u32 sl, %2, %12;\n\t"                                 \
      "mad.wide.u32 ad, sl, 2, %11;\n\t"                         \
      "@pneed ld.global.nc.u16 %4, [ad];\n\t"                    \
      "@pneed mov.u32 %1, 0xFFFF;\n\t"                           \
      "@pneed mov.u32 %0, 0;\n\t"                                \
      PCS_DEC_SKIP_CLOSE                                         \
      MULSHR                                                     \
      "add.u32 lim, %0, t;\n\t"                                  \
      "setp.ge.and.u32 pb|pnb, %4, lim, pact;\n\t"               \
      "@pb sub.u32 %1, %1, t;\n\t"                               \
      "@pb mov.u32 %0, lim;\n\t"                                 \
      "@pnb add.u32 %1, t, -1;\n\t"                              \
      "selp.u32 %5, 1, 0, pb;\n\t"                               \
      "}"
template <int PREC, int WORDS>
__device__ __forceinline__ unsigned dec_step(unsigned actw, unsigned bitm, unsigned p, DecState& c,
                                             const unsigned short* in1, unsigned prec) {
  unsigned b;
  if constexpr (PREC == 7) {
    asm volatile(PCS_DEC_STEP_ASM("mad.lo.u32 r, %1, %9, 128;\n\tshr.u32 t, r, 7;\n\t")
                 : "+r"(c.lower), "+r"(c.size), "+r"(c.slot), "+r"(c.count), "+r"(c.code), "=r"(b)
                 : "r"(actw), "r"(bitm), "r"(c.lt), "r"(p), "r"(prec), "l"(in1), "n"(WORDS - 2)
                 : "memory");
    return b;
  }
  asm volatile(PCS_DEC_STEP_ASM("shl.b32 t, 1, %10;\n\tmad.lo.u32 r, %1, %9, t;\n\tshr.u32 t, r, %10;\n\t")
               : "+r"(c.lower), "+r"(c.size), "+r"(c.slot), "+r"(c.count), "+r"(c.code), "=r"(b)
               : "r"(actw), "r"(bitm), "r"(c.lt), "r"(p), "r"(prec), "l"(in1), "n"(WORDS - 2)
               : "memory");
  return b;
}


// The decoder keeps 5 CTAs per SM (96 registers, 8-16 bytes of spill: 0.88 M local loads + 0.47 M local stores in 4.67 G warp
// instructions, 0.03 %): 4 CTAs per SM is spill-free but 1.8 % slower (6.08 vs 5.98 ms per 32 4K frames).
#ifndef PCS_DEC_MINB
#define PCS_DEC_MINB 5
#endif
// The nibble-transposed context words of a 32-row half are pinned where they are computed (PCS_DEC_PIN = 2); the compiler
// otherwise re-derives single words of the transposition network inside the row loop (about ten instructions per use).
// cfg4, 32 frames: 5.14 ms unpinned, 5.04 sign-context words only (1), 4.92 all sixteen (2); profiles/r2s_bpc_variants.json.
#ifndef PCS_DEC_PIN
#define PCS_DEC_PIN 2
#endif
__device__ __forceinline__ unsigned lds_u8(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

template <int PREC, int ROWS>
__global__ void __launch_bounds__(128, PCS_DEC_MINB) bpc_decode_kernel(const unsigned short* __restrict__ cw, const int* __restrict__ sizes,
                                                                      int* __restrict__ q, const BpcArgs p) {
  // decoded magnitude bit planes, same layout as the encoder's: word [32*half + k][lane], k < 16: plane k of
  // column 0, k >= 16: plane k-16 of column 1
  __shared__ unsigned s_pl[4][64][32];
  __shared__ unsigned char s_sign[64 + 256];  // sign-context tables: column 0 at [0, 64), column 1 at [64, 320)
  unsigned char* const s_sign0 = s_sign;
  unsigned char* const s_sign1 = s_sign + 64;
  // probability bytes per table row: [0..8] significance contexts, [9] refinement, [12..15] sign contexts
  extern __shared__ uint4 s_tab[];  // min(nrows, kMaxProbRows) + 1 rows (the launch sizes it): a 4K / wl = 5 table is 4 KB, not 8
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_sign1[i] = (unsigned char)sign_entry1((unsigned)i);
    if (i < 64) s_sign0[i] = (unsigned char)sign_entry0((unsigned)i);
  }
  const int zero_row = min(p.nrows, kMaxProbRows);  // rows outside the table read as zero probabilities
  for (int i = threadIdx.x; i <= zero_row; i += blockDim.x) s_tab[i] = i < zero_row ? __ldg(p.tab4 + i) : make_uint4(0, 0, 0, 0);
  __syncthreads();
  const unsigned char* const s_prob = reinterpret_cast<const unsigned char*>(s_tab);
  // 32-bit shared address of the sign tables, made opaque so that it stays in a register instead of being rebuilt
  // from the CTA's shared window at every use (4 uniform-datapath instructions per sign call site)
  unsigned sgn0 = (unsigned)__cvta_generic_to_shared(s_sign);
  asm volatile("mov.u32 %0, %0;" : "+r"(sgn0));
  // Everything the control flow of this kernel depends on is the same in all lanes of a warp; the broadcasts below tell
  // the compiler so (it otherwise guards every vote / redux of the coder with a divergence check: 30 BRA.DIV + UMOV pairs).
  const int lane = threadIdx.x & 31, wib = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
  const int w = blockIdx.x * (blockDim.x >> 5) + wib;
  if (w >= p.total_cb) return;
  const int frame = (w / p.ncb) * p.pstep + p.pfirst, cb = w % p.ncb;
  constexpr int kHalves = ROWS / 32, kWords = kCB * ROWS;
  const int bx = cb % p.nbx, by = cb / p.nbx;
  int* dst = q + (size_t)frame * p.A + (size_t)(by * ROWS) * p.Wa + bx * kCB + 2 * lane;
  const size_t blk = (size_t)frame * p.ncb + cb;
  const unsigned short* in = cw + blk * kWords;
  const int size = __shfl_sync(kFull, sizes[blk], 0);
  const int MSB = __shfl_sync(kFull, (int)in[0], 0);
  unsigned (*pl)[32] = s_pl[wib];

  if (size == kWords) {  // raw block (BPCEngine.cu:2197-2200,1914-1922,94-111)
    for (int i = 0; i < ROWS; ++i) {
      const unsigned a = in[lane * 2 * ROWS + 2 * i], b = in[lane * 2 * ROWS + 2 * i + 1];
      const int va = (int)(a >> 1), vb = (int)(b >> 1);
      *reinterpret_cast<int2*>(dst + (size_t)i * p.Wa) = make_int2((a & 1u) ? -va : va, (b & 1u) ? -vb : vb);
    }
    return;
  }
  if (MSB >= 23) {  // 32 = all-zero block; 23..31 cannot be produced by the coder (|q| < 2^23, Annex B.2) -> zeros
    for (int i = 0; i < ROWS; ++i) *reinterpret_cast<int2*>(dst + (size_t)i * p.Wa) = make_int2(0, 0);
    return;
  }
  const bool hi_planes = MSB >= 16;  // rare: planes >= 16 are accumulated in the output directly
  if (hi_planes)
    for (int i = 0; i < ROWS; ++i) *reinterpret_cast<int2*>(dst + (size_t)i * p.Wa) = make_int2(0, 0);
#pragma unroll
  for (int k = 0; k < 64; ++k) pl[k][lane] = 0u;

  int row = lane_group(bx * kCB + 2 * lane, by * ROWS, p) * p.nBp + MSB;
  const unsigned short* const in1 = in + 1;
  const unsigned prec = (unsigned)p.prec;
  DecState ac; ac.lower = 0; ac.size = 0; ac.slot = 0; ac.count = 0; ac.code = 0; ac.lt = (1u << lane) - 1u;
  const unsigned not31 = lane == 31 ? 0u : ~0u, not0 = lane == 0 ? 0u : ~0u;  // masks for the shuffled neighbour words
  u64 s0 = 0, s1 = 0, g0 = 0, g1 = 0;  // significant before this plane; sign (1 = negative) of significant samples
  for (int bp = MSB; bp >= 0; --bp, --row) {
    const unsigned char* const prob = s_prob + 16 * ((row >= 0 && row < zero_row) ? row : zero_row);
    const unsigned pref = prob[9];
    const u64 Lb = shfl_up64(s1, lane), Lg = shfl_up64(g1, lane), Rb = shfl_down64(s0, lane);

    u64 B0 = 0, B1 = 0;  // decoded 1-bits of this plane
    u64 N0 = 0, N1 = 0;  // newly significant in this pass
    // carried from the previous row: "became significant in this pass" flags; u0 / u1 = (significant | negative << 1)
    // of row i-1, kept pre-shifted to their position in the sign-table index (<< 4 / << 6)
    unsigned pns = 0, pnl = 0, pnr = 0, u0 = 0, u1 = 0;
#pragma unroll 1
    for (int h = 0; h < kHalves; ++h) {
      const int sh = 32 * h;
      const unsigned S0 = rows32(s0, sh), S1 = rows32(s1, sh), G0 = rows32(g0, sh), G1 = rows32(g1, sh);
      const unsigned A0 = ~S0, A1 = ~S1;
      const unsigned anyact = __reduce_or_sync(kFull, A0 | A1);
      // this half's decoded bits (= new flags) and new signs; rows are shifted in from the top, so row i ends at bit i
      unsigned nb0 = 0, nb1 = 0, ng0 = 0, ng1 = 0;
      if (anyact != 0) {
        // Static part of the contexts of these 32 rows: state BEFORE this pass of all eight neighbours.  Built per half from
        // the 64-row state words, on 32-bit slices, so that only those words stay live across the row loop.
        const unsigned Lbs = rows32(Lb, sh), Rbs = rows32(Rb, sh);
        unsigned n0[4], n1[4];
        count8(rows32_prev(Lb, sh), rows32_prev(s0, sh), rows32_prev(s1, sh), Lbs, S1, rows32_next(Lb, sh), rows32_next(s0, sh),
               rows32_next(s1, sh), n0);
        count8(rows32_prev(s0, sh), rows32_prev(s1, sh), rows32_prev(Rb, sh), S0, Rbs, rows32_next(s0, sh), rows32_next(s1, sh),
               rows32_next(Rb, sh), n1);
        // static sign-context inputs.  Column 0: its horizontal pair (left lane's column 1, own column 1) is not
        // visited before it -> the sign of their sum, presented as a fake "left" neighbour; the sample below is
        // never visited before.
        const unsigned Lgs = rows32(Lg, sh);
        const unsigned Lp = Lbs & ~Lgs, Ln = Lbs & Lgs, Rp = S1 & ~G1, Rn = S1 & G1;
        const unsigned Hp0 = (Lp & ~Rn) | (Rp & ~Ln), Hn0 = (Ln & ~Rp) | (Rn & ~Lp);
        unsigned c0[4] = {Hp0 | Hn0 /* |h| != 0 */, Hn0, rows32_next(s0, sh), rows32_next(s0 & g0, sh)};
        unsigned c1[4];
        nibble_net(n0); nibble_net(n1); nibble_net(c0);
        nibble_net2(rows32_next(s1, sh), rows32_next(s1 & g1, sh), c1);
#if PCS_DEC_PIN >= 1
        asm volatile("" : "+r"(c0[0]), "+r"(c0[1]), "+r"(c0[2]), "+r"(c0[3]), "+r"(c1[0]), "+r"(c1[1]), "+r"(c1[2]), "+r"(c1[3]));
#endif
#if PCS_DEC_PIN >= 2
        asm volatile("" : "+r"(n0[0]), "+r"(n0[1]), "+r"(n0[2]), "+r"(n0[3]), "+r"(n1[0]), "+r"(n1[1]), "+r"(n1[2]), "+r"(n1[3]));
#endif
#pragma unroll 1
        for (int g = 0; g < 8; ++g) {  // rows 4g .. 4g+3
          const int r0 = 4 * g;
          if (((anyact >> r0) & 15u) == 0) {
            // nobody decodes in these rows: nothing becomes significant; carry the state of row 4g+3
            pns = 0; pnl = 0; pnr = 0;
            u0 = (((S0 >> (r0 + 3)) & 1u) | (((G0 >> (r0 + 3)) & 1u) << 1)) << 4;
            u1 = (((S1 >> (r0 + 3)) & 1u) | (((G1 >> (r0 + 3)) & 1u) << 1)) << 6;
            nb0 >>= 4; nb1 >>= 4; ng0 >>= 4; ng1 >>= 4;
            continue;
          }
          unsigned a40 = A0 >> r0, a41 = A1 >> r0, s40 = S0 >> r0, s41 = S1 >> r0, g40 = G0 >> r0, g41 = G1 >> r0;
#if PCS_DEC_PIN >= 3
          asm volatile("" : "+r"(a40), "+r"(a41), "+r"(s40), "+r"(s41), "+r"(g40), "+r"(g41));  // one shift per group, not per use
#endif
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const unsigned bitm = 1u << qq;
            // ---- column 0 ----
            unsigned ctx = (n0[qq] >> r0) + pnl + pns;  // bits above the nibble are masked at the look-up
            const unsigned cn0 = dec_step<PREC, kWords>(a40, bitm, prob[ctx & 15u], ac, in1, prec);
            unsigned neg0 = (g40 >> qq) & 1u;
            if (__any_sync(kFull, cn0)) {
              // horizontal: static (sign of the static sum), up: row i-1 now, down: static
              const unsigned sc = lds_u8(sgn0 + (((c0[qq] >> r0) & 15u) | u0));
              const unsigned sym = dec_step<PREC, kWords>(cn0, 1u, prob[12u + (sc >> 1)], ac, in1, prec);
              if (cn0) neg0 = (sym ^ sc) & 1u;  // BPCEngine.cu:583-589
            }
            nb0 = __funnelshift_r(nb0, cn0, 1); ng0 = __funnelshift_r(ng0, cn0 & neg0, 1);
            const unsigned cur0 = (((s40 >> qq) & 1u) | cn0) | (neg0 << 1);  // significant | negative << 1
            // right neighbour's (i, column 0): state and new flag
            unsigned rv = __shfl_down_sync(kFull, cur0 | (cn0 << 2), 1);
            rv &= not31;  // no right neighbour outside the codeblock
            const unsigned cnr = rv >> 2;
            // ---- column 1 ----
            ctx = (n1[qq] >> r0) + pns + pnr + cn0 + cnr;
            const unsigned cn1 = dec_step<PREC, kWords>(a41, bitm, prob[ctx & 15u], ac, in1, prec);
            unsigned neg1 = (g41 >> qq) & 1u;
            if (__any_sync(kFull, cn1)) {
              const unsigned sc = lds_u8(sgn0 + 64u + (((c1[qq] >> r0) & 3u) | (cur0 << 2) | ((rv & 3u) << 4) | u1));
              const unsigned sym = dec_step<PREC, kWords>(cn1, 1u, prob[12u + (sc >> 1)], ac, in1, prec);
              if (cn1) neg1 = (sym ^ sc) & 1u;
            }
            nb1 = __funnelshift_r(nb1, cn1, 1); ng1 = __funnelshift_r(ng1, cn1 & neg1, 1);
            const unsigned cur1 = (((s41 >> qq) & 1u) | cn1) | (neg1 << 1);
            unsigned lv = __shfl_up_sync(kFull, cn1, 1);
            lv &= not0;
            pns = cn0 + cn1; pnl = lv; pnr = cnr;
            u0 = cur0 << 4; u1 = cur1 << 6;
          }
        }
      } else {
        pns = 0; pnl = 0; pnr = 0;
        u0 = ((S0 >> 31) | ((G0 >> 31) << 1)) << 4;
        u1 = ((S1 >> 31) | ((G1 >> 31) << 1)) << 6;
      }
      B0 |= (u64)nb0 << sh; B1 |= (u64)nb1 << sh;
      N0 |= (u64)nb0 << sh; N1 |= (u64)nb1 << sh;
      g0 |= (u64)ng0 << sh; g1 |= (u64)ng1 << sh;
    }
    // ---- magnitude refinement pass (BPCEngine.cu:742-762,1267-1279): only rows in which some lane refines ----
#pragma unroll 1
    for (int h = 0; h < kHalves; ++h) {
      const int sh = 32 * h;
      const unsigned r0m = (unsigned)(s0 >> sh), r1m = (unsigned)(s1 >> sh);
      unsigned nb0 = 0, nb1 = 0;
      unsigned U = __reduce_or_sync(kFull, r0m | r1m);
      while (U) {
        const unsigned bitm = U & (0u - U);  // lowest row in which some lane refines
        U ^= bitm;
        nb0 |= (0u - dec_step<PREC, kWords>(r0m, bitm, pref, ac, in1, prec)) & bitm;
        nb1 |= (0u - dec_step<PREC, kWords>(r1m, bitm, pref, ac, in1, prec)) & bitm;
      }
      B0 |= (u64)nb0 << sh; B1 |= (u64)nb1 << sh;
    }
    s0 |= N0; s1 |= N1;
    if (bp < 16) {
      pl[bp][lane] = (unsigned)B0; pl[32 + bp][lane] = (unsigned)(B0 >> 32);
      pl[16 + bp][lane] = (unsigned)B1; pl[48 + bp][lane] = (unsigned)(B1 >> 32);
    } else {
      for (int i = 0; i < ROWS; ++i) {
        int2* d = reinterpret_cast<int2*>(dst + (size_t)i * p.Wa);
        int2 v = *d;
        v.x |= (int)((B0 >> i) & 1ull) << bp; v.y |= (int)((B1 >> i) & 1ull) << bp;
        *d = v;
      }
    }
  }
  __syncwarp();
  // assemble magnitudes from the bit planes (transpose back), apply signs and write (BPCEngine.cu:94-111)
#pragma unroll 1
  for (int h = 0; h < kHalves; ++h) {
    unsigned wv[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) wv[k] = pl[32 * h + k][lane];
    transpose32(wv);
    const unsigned sg0 = (unsigned)(g0 >> (32 * h)), sg1 = (unsigned)(g1 >> (32 * h));
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      int2* d = reinterpret_cast<int2*>(dst + (size_t)(32 * h + i) * p.Wa);
      unsigned a = wv[i] & 0xFFFFu, b = wv[i] >> 16;
      if (hi_planes) { const int2 o = *d; a |= (unsigned)o.x; b |= (unsigned)o.y; }
      const int va = (int)a, vb = (int)b;
      *d = make_int2(((sg0 >> i) & 1u) ? -va : va, ((sg1 >> i) & 1u) ? -vb : vb);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Complexity-scalable mode (-k > 0).  The planes below `consec` of a codeblock are coded coefficient by
// coefficient ("bulk" pass, BPC/BPCEngine.cu:1285-1662,1683-1716,1776-1837) with the LUT set min(consec, MSB) of the
// multi-file tables (Engines/Engine.cu:12-100).  This path is a plain lane-per-column-pair kernel with the state words
// of the reference (magnitude << 1 | sign, flags in bits 31 / 29, significance plane in bits 28:24) in shared memory:
// it is the SIMT form of oracle/picsong_oracle.c (pcso_bpc_encode_k / pcso_bpc_decode_k), which is pinned byte for
// byte on reference runs.  Functionally complete, not yet tuned: the bit-sliced kernels above remain the k = 0 path.
// As in the oracle, lane 0's subband decides `consec` for the whole block (exact for every block inside one subband;
// a block that straddles subbands is undefined behaviour in the reference, whose own decoder aborts on it).
// ---------------------------------------------------------------------------------------------
__constant__ float kL2Norm[10][4] = {  // BPC/BPCEngine.cuh:158-169
    {1.965908f, 1.0112865f, 1.0112865f, 0.52021784f}, {4.1224113f, 1.9968134f, 1.9968134f, 0.96721643f},
    {8.416739f, 4.1833673f, 4.1833673f, 2.0792568f},  {16.935543f, 8.534108f, 8.534108f, 4.3004827f},
    {33.924816f, 17.166693f, 17.166693f, 8.686718f},  {67.87687f, 34.385098f, 34.385098f, 17.41882f},
    {135.76744f, 68.7964f, 68.7964f, 34.860676f},     {271.5416f, 137.60588f, 137.60588f, 69.73287f},
    {543.0866f, 275.21814f, 275.21814f, 139.47136f},  {1086.1624f, 550.43286f, 550.43286f, 278.94202f}};

struct KCoder { unsigned lower, size, slot, code, count; };

__device__ __forceinline__ int klut(const BpcKArgs& a, int idx) { return (idx >= 0 && idx < a.lutN) ? __ldg(a.lut + idx) : 0; }
__device__ __forceinline__ void k_subband(int X, int Y, const BpcKArgs& a, int& lvl, int& sb) {  // BPCEngine.cu:143-170
  for (int e = 1; e <= a.wl; ++e) {
    const bool bx = X >= (a.Wa >> e), by = Y >= (a.Ha >> e);
    if (bx || by) { lvl = e - 1; sb = bx ? (by ? 2 : 0) : 1; return; }
  }
  lvl = a.wl; sb = 0;
}
__device__ __forceinline__ int k_sign_hv(int h, int v) {  // BPCEngine.cu:252-293
  if (h == 0) return v == 0 ? 0 : (v > 0 ? 2 : 3);
  if (h > 0) return v == 0 ? 4 : (v > 0 ? 6 : 0);
  return v == 0 ? 5 : (v > 0 ? 1 : 7);
}
__device__ __forceinline__ int k_sgn(unsigned s) { return (s >> 31) ? ((s & 1u) ? -1 : 1) : 0; }
__device__ __forceinline__ int k_sgn_bulk(unsigned s, int cur) {
  return ((s >> 31) && ((s >> 24) & 31u) >= (unsigned)cur) ? ((s & 1u) ? -1 : 1) : 0;
}
// neighbour snapshot of (row i, column c) of this lane: 0 up-left, 1 up, 2 up-right, 3 left, 4 right, 5 down-left,
// 6 down, 7 down-right; outside the codeblock counts as 0 (BPCEngine.cu:446-484)
__device__ __forceinline__ void k_neighbours(const unsigned (*S)[32], int lane, int i, int c, unsigned (&n)[8]) {
  auto at = [&](int l, int row, int col) -> unsigned {
    return (row < 0 || row >= kCB || l < 0 || l > 31) ? 0u : S[2 * row + col][l];
  };
  if (c == 0) {
    n[0] = at(lane - 1, i - 1, 1); n[1] = at(lane, i - 1, 0); n[2] = at(lane, i - 1, 1);
    n[3] = at(lane - 1, i, 1);                                 n[4] = at(lane, i, 1);
    n[5] = at(lane - 1, i + 1, 1); n[6] = at(lane, i + 1, 0); n[7] = at(lane, i + 1, 1);
  } else {
    n[0] = at(lane, i - 1, 0); n[1] = at(lane, i - 1, 1); n[2] = at(lane + 1, i - 1, 0);
    n[3] = at(lane, i, 0);                                 n[4] = at(lane + 1, i, 0);
    n[5] = at(lane, i + 1, 0); n[6] = at(lane, i + 1, 1); n[7] = at(lane + 1, i + 1, 0);
  }
}
__device__ __forceinline__ int k_ctx(const unsigned (&n)[8]) {
  int c = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) c += (int)(n[j] >> 31);
  return c;
}
__device__ __forceinline__ int k_ctx_bulk(const unsigned (&n)[8], int B) {
  int c = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) c += (int)(((n[j] >> 24) & 31u) >= (unsigned)B);
  return c;
}
// one call site of the interval coder for the whole warp (BPCEngine.cu:371-442), slots handed out in lane order
__device__ __forceinline__ void k_enc(KCoder& c, bool act, unsigned bit, int p, unsigned prec, unsigned short* out, unsigned lt) {
  const bool need = act && c.size == 0;
  const unsigned m = __ballot_sync(kFull, need);
  if (need) { c.lower = 0; c.size = 0xFFFFu; c.slot = 1u + min(c.count + (unsigned)__popc(m & lt), 4094u); }
  c.count = min(c.count + (unsigned)__popc(m), 4095u);
  if (act) {
    const unsigned t = ((c.size * (unsigned)p) >> prec) + bit;
    if (bit == 0) c.size = t;
    else { c.size -= t; c.lower += t; }
    if (c.size == 0) out[c.slot] = (unsigned short)c.lower;
  }
}
__device__ __forceinline__ unsigned k_dec(KCoder& c, bool act, int p, unsigned prec, const unsigned short* in, unsigned lt) {
  const bool need = act && c.size == 0;
  const unsigned m = __ballot_sync(kFull, need);
  if (need) { c.lower = 0; c.size = 0xFFFFu; c.slot = 1u + min(c.count + (unsigned)__popc(m & lt), 4094u); c.code = in[c.slot]; }
  c.count = min(c.count + (unsigned)__popc(m), 4095u);
  unsigned sym = 0;
  if (act) {
    const unsigned t = ((c.size * (unsigned)p) >> prec) + 1u, lim = c.lower + t;
    if (c.code >= lim) { c.size -= t; c.lower = lim; sym = 1; }
    else c.size = t - 1u;
  }
  return sym;
}
// planes [0, consec) of a block go to the bulk pass (BPCEngine.cu:1683-1692)
__device__ __forceinline__ int k_consec(const BpcKArgs& a, int X, int Y, int MSB) {
  int lvl, sb;
  k_subband(X, Y, a, lvl, sb);
  const float norm = (a.wl == lvl) ? kL2Norm[max(lvl - 1, 0)][0] : kL2Norm[lvl][3 - sb];
  return max((int)floorf((float)MSB * __fdiv_rn(a.k, norm)), 0);
}
__device__ __forceinline__ void k_pointers(const BpcKArgs& a, int X, int Y, int lane, int MSB, int consec, int& refP, int& sigP, int& sgnP) {
  int lvl, sb;
  k_subband(X + 2 * lane, Y, a, lvl, sb);  // per lane, as the reference (BPCEngine.cu:1975-1990)
  const int R = a.nSub * a.nBp * a.cR * a.wl + a.nBp * a.cR, Sg = a.nSub * a.nBp * a.cS * a.wl + a.nBp * a.cS;
  const int g = (lvl * a.nSub + sb) * a.nBp + MSB, off = min(consec, MSB) * a.per;  // LUT set (BPCEngine.cu:329-350)
  refP = g * a.cR + off; sigP = R + g * a.cS + off; sgnP = R + Sg + g * a.cG + off;
}

constexpr int kKWarps = 2;

template <typename T>
__global__ void __launch_bounds__(32 * kKWarps) bpc_encode_k_kernel(const T* __restrict__ coeff, unsigned short* __restrict__ cw,
                                                                   int* __restrict__ sizes, const BpcKArgs a) {
  __shared__ unsigned s_S[kKWarps][2 * kCB][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int w = blockIdx.x * kKWarps + wib;
  if (w >= a.total_cb) return;
  const int frame = (w / a.ncb) * a.pstep + a.pfirst, cb = w % a.ncb;
  const int X = (cb % a.nbx) * kCB, Y = (cb / a.nbx) * kCB;
  const T* src = coeff + (size_t)frame * a.A + (size_t)Y * a.Wa + X + 2 * lane;
  const size_t blk = (size_t)frame * a.ncb + cb;
  unsigned short* out = cw + blk * kCBWords;
  unsigned (*S)[32] = s_S[wib];
  const unsigned lt = (1u << lane) - 1u, prec = (unsigned)a.prec;

  unsigned ormag = 0;
  for (int i = 0; i < kCB; ++i) {
    int v0, v1;
    load2<T>(src + (size_t)i * a.Wa, v0, v1);
    const unsigned m0 = umag(v0), m1 = umag(v1);
    S[2 * i][lane] = (m0 << 1) + (v0 < 0 ? 1u : 0u);  // BPCEngine.cu:41-63
    S[2 * i + 1][lane] = (m1 << 1) + (v1 < 0 ? 1u : 0u);
    ormag |= m0 | m1;
  }
  ormag = __reduce_or_sync(kFull, ormag);
  const int MSB = ormag ? 31 - __clz(ormag) : 32;
  if (MSB == 32) {
    if (lane == 0) { out[0] = 32; sizes[blk] = 1; }
    return;
  }
  if (lane == 0) out[0] = (unsigned short)MSB;
  __syncwarp();
  const int consec = k_consec(a, X, Y, MSB);
  int refP, sigP, sgnP;
  k_pointers(a, X, Y, lane, MSB, consec, refP, sigP, sgnP);
  KCoder ac; ac.lower = 0; ac.size = 0; ac.slot = 0; ac.code = 0; ac.count = 0;
  unsigned n[8];
  int bp;
  for (bp = MSB; bp >= consec; --bp) {
    // significance propagation pass (BPCEngine.cu:770-843, 490-516)
    for (int i = 0; i < kCB; ++i)
      for (int c = 0; c < 2; ++c) {
        k_neighbours(S, lane, i, c, n);
        unsigned wd = S[2 * i + c][lane];
        __syncwarp();
        const bool act = !(wd >> 31);
        const unsigned bit = (wd >> (bp + 1)) & 1u;
        k_enc(ac, act, bit, klut(a, sigP + k_ctx(n)), prec, out, lt);
        const bool one = act && bit;
        int sc = 0;
        if (one) {
          wd |= 1u << 31; wd |= (unsigned)bp << 24;
          sc = k_sign_hv(k_sgn(n[3]) + k_sgn(n[4]), k_sgn(n[1]) + k_sgn(n[6]));
        }
        k_enc(ac, one, ((wd & 1u) == (unsigned)(sc & 1)) ? 0u : 1u, klut(a, sgnP + (sc >> 1)), prec, out, lt);
        S[2 * i + c][lane] = wd;
        __syncwarp();
      }
    // magnitude refinement pass (BPCEngine.cu:1249-1261, 726-736)
    for (int i = 0; i < kCB; ++i)
      for (int c = 0; c < 2; ++c) {
        unsigned wd = S[2 * i + c][lane];
        const bool act = (wd >> 29) & 1u;
        if (!act && (wd >> 31)) { wd |= 1u << 29; S[2 * i + c][lane] = wd; }
        k_enc(ac, act, (wd >> (bp + 1)) & 1u, klut(a, refP), prec, out, lt);
      }
    __syncwarp();
    refP -= a.cR; sigP -= a.cS; sgnP -= a.cG;
  }
  if (bp >= 0) {
    // bulk pass: planes B..0 of one coefficient at a time; context and neighbour snapshot fixed per coefficient
    const int B = bp;
    for (int i = 0; i < kCB; ++i)
      for (int c = 0; c < 2; ++c) {
        k_neighbours(S, lane, i, c, n);
        unsigned wd = S[2 * i + c][lane];
        __syncwarp();
        const bool any = __any_sync(kFull, !(wd >> 31));
        const int ctx = any ? (B != 0 ? k_ctx_bulk(n, B) : k_ctx(n)) : 0;
        for (int cur = B; cur >= 0; --cur) {
          const int d = B - cur;
          const bool sig = wd >> 31;
          const unsigned bit = (wd >> (cur + 1)) & 1u;
          k_enc(ac, sig, bit, klut(a, refP - a.cR * d), prec, out, lt);
          k_enc(ac, !sig, bit, klut(a, sigP + ctx - a.cS * d), prec, out, lt);
          const bool one = !sig && bit;
          int sc = 0;
          if (one) {
            wd |= 1u << 31; wd |= (unsigned)cur << 24;
            sc = k_sign_hv(k_sgn_bulk(n[3], cur) + k_sgn_bulk(n[4], cur), k_sgn_bulk(n[1], cur) + k_sgn_bulk(n[6], cur));
          }
          k_enc(ac, one, ((wd & 1u) == (unsigned)(sc & 1)) ? 0u : 1u, klut(a, sgnP + (sc >> 1) - a.cG * d), prec, out, lt);
        }
        S[2 * i + c][lane] = wd;
        __syncwarp();
      }
  }
  out[ac.slot] = (unsigned short)ac.lower;  // flush (BPCEngine.cu:1719)
  const int size = (int)ac.count + 1;
  __syncwarp();
  if (size == kCBWords) {  // expansion fallback (BPCEngine.cu:2015-2024,1904-1912)
    for (int kk = 0; kk < 2 * kCB; ++kk) out[lane * 2 * kCB + kk] = (unsigned short)(S[kk][lane] & 0xFFFFu);
  }
  if (lane == 0) sizes[blk] = size;
}

__global__ void __launch_bounds__(32 * kKWarps) bpc_decode_k_kernel(const unsigned short* __restrict__ cw, const int* __restrict__ sizes,
                                                                   int* __restrict__ q, const BpcKArgs a) {
  __shared__ unsigned s_S[kKWarps][2 * kCB][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int w = blockIdx.x * kKWarps + wib;
  if (w >= a.total_cb) return;
  const int frame = (w / a.ncb) * a.pstep + a.pfirst, cb = w % a.ncb;
  const int X = (cb % a.nbx) * kCB, Y = (cb / a.nbx) * kCB;
  int* dst = q + (size_t)frame * a.A + (size_t)Y * a.Wa + X + 2 * lane;
  const size_t blk = (size_t)frame * a.ncb + cb;
  const unsigned short* in = cw + blk * kCBWords;
  unsigned (*S)[32] = s_S[wib];
  const unsigned lt = (1u << lane) - 1u, prec = (unsigned)a.prec;
  const int size = sizes[blk];
  const int MSB = in[0];
  for (int kk = 0; kk < 2 * kCB; ++kk) S[kk][lane] = 0u;
  __syncwarp();
  if (size == kCBWords) {  // raw block (BPCEngine.cu:2197-2200,1914-1922)
    for (int kk = 0; kk < 2 * kCB; ++kk) S[kk][lane] = in[lane * 2 * kCB + kk];
  } else if (MSB < 23) {   // 32 = all-zero block; 23..31 cannot be produced by the coder
    const int consec = k_consec(a, X, Y, MSB);
    int refP, sigP, sgnP;
    k_pointers(a, X, Y, lane, MSB, consec, refP, sigP, sgnP);
    KCoder ac; ac.lower = 0; ac.size = 0; ac.slot = 0; ac.code = 0; ac.count = 0;
    unsigned mask = 0x3u << MSB;  // decoded bit + mid-point approximation bit (BPCEngine.cu:1791-1824)
    if (MSB == 0) mask &= 0x2u;
    unsigned n[8];
    int bp;
    for (bp = MSB; bp >= consec; --bp) {
      for (int i = 0; i < kCB; ++i)
        for (int c = 0; c < 2; ++c) {
          k_neighbours(S, lane, i, c, n);
          unsigned wd = S[2 * i + c][lane];
          __syncwarp();
          const bool act = !(wd >> 31);
          const bool one = k_dec(ac, act, klut(a, sigP + k_ctx(n)), prec, in, lt) != 0;
          int sc = 0;
          if (one) {
            wd |= mask; wd |= 1u << 31; wd |= (unsigned)bp << 24;
            sc = k_sign_hv(k_sgn(n[3]) + k_sgn(n[4]), k_sgn(n[1]) + k_sgn(n[6]));
          }
          const unsigned sym = k_dec(ac, one, klut(a, sgnP + (sc >> 1)), prec, in, lt);
          if (one) wd |= ((sym & 1u) == (unsigned)(sc & 1)) ? 0u : 1u;
          S[2 * i + c][lane] = wd;
          __syncwarp();
        }
      for (int i = 0; i < kCB; ++i)
        for (int c = 0; c < 2; ++c) {
          unsigned wd = S[2 * i + c][lane];
          const bool act = (wd >> 29) & 1u;
          if (!act && (wd >> 31)) wd |= 1u << 29;
          const unsigned sym = k_dec(ac, act, klut(a, refP), prec, in, lt);
          if (act) { wd &= ~mask; wd |= mask & (((sym << 1) + 1u) << bp); }
          S[2 * i + c][lane] = wd;
        }
      __syncwarp();
      mask >>= 1;
      if (bp == 1) mask = 0x2u;
      refP -= a.cR; sigP -= a.cS; sgnP -= a.cG;
    }
    if (bp >= 0) {
      const int B = bp;
      const unsigned mask0 = mask;
      for (int i = 0; i < kCB; ++i)
        for (int c = 0; c < 2; ++c) {
          k_neighbours(S, lane, i, c, n);
          unsigned wd = S[2 * i + c][lane];
          __syncwarp();
          const bool any = __any_sync(kFull, !(wd >> 31));
          const int ctx = any ? (B != 0 ? k_ctx_bulk(n, B) : k_ctx(n)) : 0;
          unsigned m = mask0;
          for (int cur = B; cur >= 0; --cur) {
            const int d = B - cur;
            const bool sig = wd >> 31;
            const unsigned rs = k_dec(ac, sig, klut(a, refP - a.cR * d), prec, in, lt);
            if (sig) { wd &= ~m; wd |= m & (((rs << 1) + 1u) << cur); }
            const bool one = k_dec(ac, !sig, klut(a, sigP + ctx - a.cS * d), prec, in, lt) != 0;
            int sc = 0;
            if (one) {
              wd |= m; wd |= 1u << 31; wd |= (unsigned)cur << 24;
              sc = k_sign_hv(k_sgn_bulk(n[3], cur) + k_sgn_bulk(n[4], cur), k_sgn_bulk(n[1], cur) + k_sgn_bulk(n[6], cur));
            }
            const unsigned sym = k_dec(ac, one, klut(a, sgnP + (sc >> 1) - a.cG * d), prec, in, lt);
            if (one) wd |= ((sym & 1u) == (unsigned)(sc & 1)) ? 0u : 1u;
            m >>= 1;
            if (cur == 1) m = 0x2u;
          }
          S[2 * i + c][lane] = wd;
          __syncwarp();
        }
    }
  }
  __syncwarp();
  for (int i = 0; i < kCB; ++i) {  // BPCEngine.cu:94-111
    const unsigned w0 = S[2 * i][lane], w1 = S[2 * i + 1][lane];
    const int v0 = (int)((w0 & 0xFFFFFFu) >> 1), v1 = (int)((w1 & 0xFFFFFFu) >> 1);
    *reinterpret_cast<int2*>(dst + (size_t)i * a.Wa) = make_int2((w0 & 1u) ? -v0 : v0, (w1 & 1u) ? -v1 : v1);
  }
}

// Debug/parity export: the reference's int32 sparse `DCodeStreamValues` layout, -1 filled.
__global__ void export_ref_cs_kernel(const unsigned short* cw, const int* sizes, int* ref, long long n, int cbw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cb = (int)(i / cbw), j = (int)(i % cbw);
  ref[i] = j < sizes[cb] ? (int)cw[i] : -1;
}

}  // namespace

int bpc_encode(const void* d_coeff, bool is_float, uint16_t* d_cw, int32_t* d_sizes, const BpcArgs& a, cudaStream_t st) {
  const int blocks = (a.total_cb + 3) / 4;
  if (a.cb_rows == 32) {
    if (is_float) bpc_encode_kernel<float, 32><<<blocks, 128, 0, st>>>(static_cast<const float*>(d_coeff), d_cw, d_sizes, a);
    else bpc_encode_kernel<int, 32><<<blocks, 128, 0, st>>>(static_cast<const int*>(d_coeff), d_cw, d_sizes, a);
  } else {
    if (is_float) bpc_encode_kernel<float, 64><<<blocks, 128, 0, st>>>(static_cast<const float*>(d_coeff), d_cw, d_sizes, a);
    else bpc_encode_kernel<int, 64><<<blocks, 128, 0, st>>>(static_cast<const int*>(d_coeff), d_cw, d_sizes, a);
  }
  ++g_launches;
  return (int)cudaGetLastError();
}

int bpc_decode(const uint16_t* d_cw, const int32_t* d_sizes, int32_t* d_q, const BpcArgs& a, cudaStream_t st) {
  const int blocks = (a.total_cb + 3) / 4;
  const size_t tab_bytes = ((size_t)std::min(a.nrows, kMaxProbRows) + 1) * sizeof(uint4);  // + the all-zero row
  if (a.cb_rows == 32) {
    if (a.prec == 7) bpc_decode_kernel<7, 32><<<blocks, 128, tab_bytes, st>>>(d_cw, d_sizes, d_q, a);
    else bpc_decode_kernel<0, 32><<<blocks, 128, tab_bytes, st>>>(d_cw, d_sizes, d_q, a);
  } else {
    if (a.prec == 7) bpc_decode_kernel<7, 64><<<blocks, 128, tab_bytes, st>>>(d_cw, d_sizes, d_q, a);
    else bpc_decode_kernel<0, 64><<<blocks, 128, tab_bytes, st>>>(d_cw, d_sizes, d_q, a);
  }
  ++g_launches;
  return (int)cudaGetLastError();
}

int bpc_encode_k(const void* d_coeff, bool is_float, uint16_t* d_cw, int32_t* d_sizes, const BpcKArgs& a, cudaStream_t st) {
  const int blocks = (a.total_cb + kKWarps - 1) / kKWarps;
  if (is_float) bpc_encode_k_kernel<float><<<blocks, 32 * kKWarps, 0, st>>>(static_cast<const float*>(d_coeff), d_cw, d_sizes, a);
  else bpc_encode_k_kernel<int><<<blocks, 32 * kKWarps, 0, st>>>(static_cast<const int*>(d_coeff), d_cw, d_sizes, a);
  ++g_launches;
  return (int)cudaGetLastError();
}

int bpc_decode_k(const uint16_t* d_cw, const int32_t* d_sizes, int32_t* d_q, const BpcKArgs& a, cudaStream_t st) {
  const int blocks = (a.total_cb + kKWarps - 1) / kKWarps;
  bpc_decode_k_kernel<<<blocks, 32 * kKWarps, 0, st>>>(d_cw, d_sizes, d_q, a);
  ++g_launches;
  return (int)cudaGetLastError();
}

int export_ref_codestream(const uint16_t* d_cw, const int32_t* d_sizes, int32_t* d_ref, long long n, int cbw, cudaStream_t st) {
  export_ref_cs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_cw, d_sizes, d_ref, n, cbw);
  ++g_launches;
  return (int)cudaGetLastError();
}

}  // namespace pcs
