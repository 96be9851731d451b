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
#include "pcs_common.cuh"

namespace pcs {

namespace {

typedef unsigned long long u64;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void add3(u64 a, u64 b, u64 c, u64& s, u64& cy) {
  s = a ^ b ^ c;
  cy = (a & b) | (c & (a ^ b));
}
// bit-sliced population count of 8 one-bit row masks -> 4 bit planes (value 0..8 per row)
__device__ __forceinline__ void count8(u64 m0, u64 m1, u64 m2, u64 m3, u64 m4, u64 m5, u64 m6, u64 m7, u64 k[4]) {
  u64 s1, c1, s2, c2, b0, c4, s5, c5;
  add3(m0, m1, m2, s1, c1);
  add3(m3, m4, m5, s2, c2);
  const u64 s3 = m6 ^ m7, c3 = m6 & m7;
  add3(s1, s2, s3, b0, c4);
  add3(c1, c2, c3, s5, c5);
  const u64 c6 = s5 & c4;
  k[0] = b0; k[1] = s5 ^ c4; k[2] = c5 ^ c6; k[3] = c5 & c6;
}

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

struct Probs { u64 sig; unsigned sgn, ref; };
__device__ __forceinline__ Probs load_probs(const BpcArgs& p, int row) {
  Probs r;
  const bool ok = row >= 0 && row < p.nrows;
  r.sig = ok ? __ldg(p.sigp + row) : 0ull;
  r.sgn = ok ? __ldg(p.sgnp + row) : 0u;
  r.ref = ok ? (unsigned)__ldg(p.refp + row) : 0u;
  return r;
}

// Per-lane interval coder + per-warp codeword counter (BPCEngine.cu:371-442).
struct Coder {
  unsigned lower, size, slot, count, lt, prec;
  __device__ __forceinline__ void init(int lane, int prec_) { lower = 0; size = 0; slot = 0; count = 0; lt = (1u << lane) - 1u; prec = prec_; }
  // lanes in `need` restart and reserve slots in ascending lane order; returns whether this lane did
  __device__ __forceinline__ bool reserve(bool need) {
    const unsigned m = __ballot_sync(kFull, need);
    if (m) {
      if (need) { slot = 1u + min(count + __popc(m & lt), 4094u); lower = 0; size = 0xFFFFu; }
      count = min(count + __popc(m), 4095u);
    }
    return need;
  }
  __device__ __forceinline__ void encode(bool act, unsigned b, unsigned p, unsigned short* out) {
    reserve(act && size == 0);
    if (act) {
      const unsigned t = ((size * p) >> prec) + b;
      if (b) { size -= t; lower += t; } else { size = t; }
      if (size == 0) out[slot] = (unsigned short)lower;
    }
  }
  __device__ __forceinline__ unsigned decode(bool act, unsigned p, unsigned& code, const unsigned short* in) {
    if (reserve(act && size == 0)) code = in[slot];
    unsigned b = 0;
    if (act) {
      const unsigned t = ((size * p) >> prec) + 1u;
      const unsigned lim = lower + t;
      if (code >= lim) { b = 1; size -= t; lower = lim; } else { size = t - 1u; }
    }
    return b;
  }
};

template <typename T> __device__ __forceinline__ void load2(const T* p, int& a, int& b);
template <> __device__ __forceinline__ void load2<int>(const int* p, int& a, int& b) {
  const int2 v = __ldg(reinterpret_cast<const int2*>(p)); a = v.x; b = v.y;
}
template <> __device__ __forceinline__ void load2<float>(const float* p, int& a, int& b) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  a = __float2int_rz(v.x); b = __float2int_rz(v.y);  // (int)c, BPCEngine.cu:49,55
}
__device__ __forceinline__ unsigned umag(int v) { return v < 0 ? 0u - (unsigned)v : (unsigned)v; }

__device__ __forceinline__ unsigned prob_sig(u64 pack, unsigned ctx) { return (unsigned)(pack >> (7u * ctx)) & 127u; }
__device__ __forceinline__ unsigned prob_sgn(unsigned pack, unsigned idx) { return (pack >> (7u * idx)) & 127u; }

// ---------------------------------------------------------------------------------------------
// encoder
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) bpc_encode_kernel(const T* __restrict__ coeff, unsigned short* __restrict__ cw,
                                                        int* __restrict__ sizes, const BpcArgs p) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= p.total_cb) return;
  const int frame = w / p.ncb, cb = w - frame * p.ncb;
  const int bx = cb % p.nbx, by = cb / p.nbx;
  const T* src = coeff + (size_t)frame * p.A + (size_t)(by * kCB) * p.Wa + bx * kCB + 2 * lane;
  unsigned short* out = cw + (size_t)w * kCBWords;

  // pass 0: MSB of the block and the sign masks (BPCEngine.cu:41-63,176-192)
  u64 g0 = 0, g1 = 0;
  unsigned ormag = 0;
#pragma unroll 8
  for (int i = 0; i < kCB; ++i) {
    int a, b; load2<T>(src + (size_t)i * p.Wa, a, b);
    ormag |= umag(a) | umag(b);
    g0 |= (u64)(a < 0) << i; g1 |= (u64)(b < 0) << i;
  }
  ormag = __reduce_or_sync(kFull, ormag);
  const int MSB = ormag ? 31 - __clz(ormag) : 32;
  if (MSB == 32) {  // nothing is coded; sizeArray = 1 (BPCEngine.cu:1998-2010)
    if (lane == 0) { out[0] = 32; sizes[w] = 1; }
    return;
  }
  int row = lane_group(bx * kCB + 2 * lane, by * kCB, p) * p.nBp + MSB;  // initializeLUTPointers :329-350
  const u64 Lg = shfl_up64(g1, lane), Rg = shfl_down64(g0, lane);       // neighbour lanes' signs

  Coder ac; ac.init(lane, p.prec);
  u64 s0 = 0, s1 = 0;  // significant BEFORE the current plane
  for (int bp = MSB; bp >= 0; --bp, --row) {
    u64 B0 = 0, B1 = 0;
#pragma unroll 8
    for (int i = 0; i < kCB; ++i) {
      int a, b; load2<T>(src + (size_t)i * p.Wa, a, b);
      B0 |= (u64)((umag(a) >> bp) & 1u) << i; B1 |= (u64)((umag(b) >> bp) & 1u) << i;
    }
    const Probs pr = load_probs(p, row);
    const u64 a0 = s0 | B0, a1 = s1 | B1;  // significant AFTER this plane's SPP visit
    const u64 Lb = shfl_up64(s1, lane), La = shfl_up64(a1, lane);
    const u64 Rb = shfl_down64(s0, lane), Ra = shfl_down64(a0, lane);
    // significance contexts of all 64 rows (BPCEngine.cu:222-230 with the visit-order rule of :770-843):
    // rows above and same-row samples visited earlier use the "after" state, everything else "before".
    u64 K0[4], K1[4];
    count8(La << 1, a0 << 1, a1 << 1, Lb, s1, Lb >> 1, s0 >> 1, s1 >> 1, K0);
    count8(a0 << 1, a1 << 1, Ra << 1, a0, Ra, s0 >> 1, s1 >> 1, Rb >> 1, K1);
    // sign contexts (BPCEngine.cu:252-308) bit-sliced: pos/neg contribution of L,R (h) and U,D (v)
    u64 SY0, I00, I01, SY1, I10, I11;
    {
      const u64 Lp = Lb & ~Lg, Ln = Lb & Lg, Rp = s1 & ~g1, Rn = s1 & g1;
      const u64 Us = a0 << 1, Ug = g0 << 1, Ds = s0 >> 1, Dg = g0 >> 1;
      const u64 Up = Us & ~Ug, Un = Us & Ug, Dp = Ds & ~Dg, Dn = Ds & Dg;
      const u64 hp = (Lp & ~Rn) | (Rp & ~Ln), hn = (Ln & ~Rp) | (Rn & ~Lp);
      const u64 vp = (Up & ~Dn) | (Dp & ~Un), vn = (Un & ~Dp) | (Dn & ~Up);
      const u64 both = (hp & vp) | (hn & vn);
      SY0 = g0 ^ (hn | (~hp & vn));
      I00 = (~(hp | hn) & (vp | vn)) | both;
      I01 = ((hp | hn) & ~(vp | vn)) | both;
    }
    {
      const u64 Lp = a0 & ~g0, Ln = a0 & g0, Rp = Ra & ~Rg, Rn = Ra & Rg;
      const u64 Us = a1 << 1, Ug = g1 << 1, Ds = s1 >> 1, Dg = g1 >> 1;
      const u64 Up = Us & ~Ug, Un = Us & Ug, Dp = Ds & ~Dg, Dn = Ds & Dg;
      const u64 hp = (Lp & ~Rn) | (Rp & ~Ln), hn = (Ln & ~Rp) | (Rn & ~Lp);
      const u64 vp = (Up & ~Dn) | (Dp & ~Un), vn = (Un & ~Dp) | (Dn & ~Up);
      const u64 both = (hp & vp) | (hn & vn);
      SY1 = g1 ^ (hn | (~hp & vn));
      I10 = (~(hp | hn) & (vp | vn)) | both;
      I11 = ((hp | hn) & ~(vp | vn)) | both;
    }

    // ---- significance propagation pass: rows in order, column 0 then column 1 (BPCEngine.cu:770-843) ----
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int sh = 32 * h;
      unsigned A0 = ~(unsigned)(s0 >> sh), A1 = ~(unsigned)(s1 >> sh), b0 = (unsigned)(B0 >> sh), b1 = (unsigned)(B1 >> sh);
      unsigned k00 = (unsigned)(K0[0] >> sh), k01 = (unsigned)(K0[1] >> sh), k02 = (unsigned)(K0[2] >> sh), k03 = (unsigned)(K0[3] >> sh);
      unsigned k10 = (unsigned)(K1[0] >> sh), k11 = (unsigned)(K1[1] >> sh), k12 = (unsigned)(K1[2] >> sh), k13 = (unsigned)(K1[3] >> sh);
      unsigned y0 = (unsigned)(SY0 >> sh), i00 = (unsigned)(I00 >> sh), i01 = (unsigned)(I01 >> sh);
      unsigned y1 = (unsigned)(SY1 >> sh), i10 = (unsigned)(I10 >> sh), i11 = (unsigned)(I11 >> sh);
#pragma unroll 2
      for (int j = 0; j < 32; ++j) {
        {
          const bool act = A0 & 1u;
          const unsigned b = b0 & 1u;
          const unsigned ctx = (k00 & 1u) | ((k01 & 1u) << 1) | ((k02 & 1u) << 2) | ((k03 & 1u) << 3);
          ac.encode(act, b, prob_sig(pr.sig, ctx), out);
          const bool sg = act && b;
          if (__any_sync(kFull, sg)) ac.encode(sg, y0 & 1u, prob_sgn(pr.sgn, (i00 & 1u) | ((i01 & 1u) << 1)), out);
        }
        {
          const bool act = A1 & 1u;
          const unsigned b = b1 & 1u;
          const unsigned ctx = (k10 & 1u) | ((k11 & 1u) << 1) | ((k12 & 1u) << 2) | ((k13 & 1u) << 3);
          ac.encode(act, b, prob_sig(pr.sig, ctx), out);
          const bool sg = act && b;
          if (__any_sync(kFull, sg)) ac.encode(sg, y1 & 1u, prob_sgn(pr.sgn, (i10 & 1u) | ((i11 & 1u) << 1)), out);
        }
        A0 >>= 1; A1 >>= 1; b0 >>= 1; b1 >>= 1;
        k00 >>= 1; k01 >>= 1; k02 >>= 1; k03 >>= 1; k10 >>= 1; k11 >>= 1; k12 >>= 1; k13 >>= 1;
        y0 >>= 1; i00 >>= 1; i01 >>= 1; y1 >>= 1; i10 >>= 1; i11 >>= 1;
      }
    }
    // ---- magnitude refinement pass: samples significant before this plane (BPCEngine.cu:726-736,1249-1261) ----
    if (__any_sync(kFull, (s0 | s1) != 0)) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int sh = 32 * h;
        unsigned r0 = (unsigned)(s0 >> sh), r1 = (unsigned)(s1 >> sh), b0 = (unsigned)(B0 >> sh), b1 = (unsigned)(B1 >> sh);
#pragma unroll 2
        for (int j = 0; j < 32; ++j) {
          ac.encode(r0 & 1u, b0 & 1u, pr.ref, out);
          ac.encode(r1 & 1u, b1 & 1u, pr.ref, out);
          r0 >>= 1; r1 >>= 1; b0 >>= 1; b1 >>= 1;
        }
      }
    }
    s0 = a0; s1 = a1;
  }
  out[ac.slot] = (unsigned short)ac.lower;  // flush (BPCEngine.cu:1719)
  const int size = (int)ac.count + 1;       // :2010
  __syncwarp();
  if (size == kCBWords) {
    // expansion fallback (BPCEngine.cu:2015-2024,1904-1912): raw low 16 bits of (|q|<<1 | sign), lane-major
    __threadfence_block();
    for (int i = 0; i < kCB; ++i) {
      int a, b; load2<T>(src + (size_t)i * p.Wa, a, b);
      out[lane * 2 * kCB + 2 * i] = (unsigned short)(((umag(a) << 1) | (a < 0 ? 1u : 0u)) & 0xFFFFu);
      out[lane * 2 * kCB + 2 * i + 1] = (unsigned short)(((umag(b) << 1) | (b < 0 ? 1u : 0u)) & 0xFFFFu);
    }
  } else if (lane == 0) {
    out[0] = (unsigned short)MSB;
  }
  if (lane == 0) sizes[w] = size;
}

// ---------------------------------------------------------------------------------------------
// decoder
// ---------------------------------------------------------------------------------------------
constexpr int kSmemPlanes = 16;
// sign context table (BPCEngine.cu:252-293) indexed by (sgn(h)+1)*3 + (sgn(v)+1), 3 bits each
constexpr unsigned kSignTab = 7u | (5u << 3) | (1u << 6) | (3u << 9) | (0u << 12) | (2u << 15) | (0u << 18) | (4u << 21) | (6u << 24);
__device__ __forceinline__ unsigned sign_ctx(int h, int v) {
  const int sh = (h > 0) - (h < 0), sv = (v > 0) - (v < 0);
  return (kSignTab >> (3 * ((sh + 1) * 3 + (sv + 1)))) & 7u;
}
__device__ __forceinline__ int contrib(unsigned sig, unsigned neg) { return sig ? (neg ? -1 : 1) : 0; }

__global__ void __launch_bounds__(128) bpc_decode_kernel(const unsigned short* __restrict__ cw, const int* __restrict__ sizes,
                                                        int* __restrict__ q, const BpcArgs p) {
  __shared__ uint4 planes[4][kSmemPlanes][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int w = blockIdx.x * (blockDim.x >> 5) + wib;
  if (w >= p.total_cb) return;
  const int frame = w / p.ncb, cb = w - frame * p.ncb;
  const int bx = cb % p.nbx, by = cb / p.nbx;
  int* dst = q + (size_t)frame * p.A + (size_t)(by * kCB) * p.Wa + bx * kCB + 2 * lane;
  const unsigned short* in = cw + (size_t)w * kCBWords;
  const int size = sizes[w];
  const int MSB = in[0];

  if (size == kCBWords) {  // raw block (BPCEngine.cu:2197-2200,1914-1922,94-111)
    for (int i = 0; i < kCB; ++i) {
      const unsigned a = in[lane * 2 * kCB + 2 * i], b = in[lane * 2 * kCB + 2 * i + 1];
      const int va = (int)(a >> 1), vb = (int)(b >> 1);
      *reinterpret_cast<int2*>(dst + (size_t)i * p.Wa) = make_int2((a & 1u) ? -va : va, (b & 1u) ? -vb : vb);
    }
    return;
  }
  if (MSB >= 23) {  // 32 = all-zero block; 23..31 cannot be produced by the coder (|q| < 2^23, Annex B.2) -> zeros
    for (int i = 0; i < kCB; ++i) *reinterpret_cast<int2*>(dst + (size_t)i * p.Wa) = make_int2(0, 0);
    return;
  }
  const bool hi_planes = MSB >= kSmemPlanes;  // rare: planes >= 16 are accumulated in the output directly
  if (hi_planes)
    for (int i = 0; i < kCB; ++i) *reinterpret_cast<int2*>(dst + (size_t)i * p.Wa) = make_int2(0, 0);

  int row = lane_group(bx * kCB + 2 * lane, by * kCB, p) * p.nBp + MSB;
  Coder ac; ac.init(lane, p.prec);
  unsigned code = 0;
  u64 s0 = 0, s1 = 0, g0 = 0, g1 = 0;  // significant before this plane; sign (1 = negative) of significant samples
  for (int bp = MSB; bp >= 0; --bp, --row) {
    const Probs pr = load_probs(p, row);
    const u64 Lb = shfl_up64(s1, lane), Lg = shfl_up64(g1, lane), Rb = shfl_down64(s0, lane);
    // static part of the contexts: state BEFORE this pass of all eight neighbours
    u64 K0[4], K1[4];
    count8(Lb << 1, s0 << 1, s1 << 1, Lb, s1, Lb >> 1, s0 >> 1, s1 >> 1, K0);
    count8(s0 << 1, s1 << 1, Rb << 1, s0, Rb, s0 >> 1, s1 >> 1, Rb >> 1, K1);
    // static sign-context inputs: column 0's horizontal pair (left lane col 1, own col 1) is not visited
    // before it, the sample below is never visited before
    const u64 Lp = Lb & ~Lg, Ln = Lb & Lg, Rp = s1 & ~g1, Rn = s1 & g1;
    const u64 Hp0 = (Lp & ~Rn) | (Rp & ~Ln), Hn0 = (Ln & ~Rp) | (Rn & ~Lp);
    const u64 Dp0 = (s0 & ~g0) >> 1, Dn0 = (s0 & g0) >> 1, Dp1 = (s1 & ~g1) >> 1, Dn1 = (s1 & g1) >> 1;

    u64 B0 = 0, B1 = 0;  // decoded 1-bits of this plane
    u64 N0 = 0, N1 = 0;  // newly significant in this pass
    // carried from the previous row: "became significant in this pass" flags and current state of (i-1)
    unsigned pn0 = 0, pn1 = 0, pnl = 0, pnr = 0, u0s = 0, u0n = 0, u1s = 0, u1n = 0;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int sh = 32 * h;
      unsigned S0 = (unsigned)(s0 >> sh), S1 = (unsigned)(s1 >> sh), G0 = (unsigned)(g0 >> sh), G1 = (unsigned)(g1 >> sh);
      unsigned k00 = (unsigned)(K0[0] >> sh), k01 = (unsigned)(K0[1] >> sh), k02 = (unsigned)(K0[2] >> sh), k03 = (unsigned)(K0[3] >> sh);
      unsigned k10 = (unsigned)(K1[0] >> sh), k11 = (unsigned)(K1[1] >> sh), k12 = (unsigned)(K1[2] >> sh), k13 = (unsigned)(K1[3] >> sh);
      unsigned hp0 = (unsigned)(Hp0 >> sh), hn0 = (unsigned)(Hn0 >> sh);
      unsigned dp0 = (unsigned)(Dp0 >> sh), dn0 = (unsigned)(Dn0 >> sh), dp1 = (unsigned)(Dp1 >> sh), dn1 = (unsigned)(Dn1 >> sh);
      unsigned nb0 = 0, nb1 = 0, nn0 = 0, nn1 = 0, ng0 = 0, ng1 = 0;  // this half's decoded bits / new flags / new signs
#pragma unroll 1
      for (int j = 0; j < 32; ++j) {
        // ---- column 0 ----
        unsigned cn0 = 0, neg0 = G0 & 1u;
        {
          const bool act = !(S0 & 1u);
          const unsigned ctx = ((k00 & 1u) | ((k01 & 1u) << 1) | ((k02 & 1u) << 2) | ((k03 & 1u) << 3)) + pnl + pn0 + pn1;
          const unsigned b = ac.decode(act, prob_sig(pr.sig, ctx), code, in);
          if (__any_sync(kFull, b)) {
            const int hh = (int)(hp0 & 1u) - (int)(hn0 & 1u);
            const int vv = contrib(u0s, u0n) + (int)(dp0 & 1u) - (int)(dn0 & 1u);
            const unsigned sc = sign_ctx(hh, vv);
            const unsigned sym = ac.decode(b != 0, prob_sgn(pr.sgn, sc >> 1), code, in);
            if (b) { cn0 = 1; neg0 = sym ^ (sc & 1u); }  // BPCEngine.cu:583-589
          }
          nb0 |= cn0 << j; nn0 |= cn0 << j; ng0 |= (cn0 & neg0) << j;
        }
        const unsigned cur0 = (S0 & 1u) | cn0;
        // right neighbour's (i, col 0): current significance, sign, new flag
        unsigned rv = __shfl_down_sync(kFull, cur0 | (neg0 << 1) | (cn0 << 2), 1);
        if (lane == 31) rv = 0;
        const unsigned cnr = (rv >> 2) & 1u;
        // ---- column 1 ----
        unsigned cn1 = 0, neg1 = G1 & 1u;
        {
          const bool act = !(S1 & 1u);
          const unsigned ctx = ((k10 & 1u) | ((k11 & 1u) << 1) | ((k12 & 1u) << 2) | ((k13 & 1u) << 3)) + pn0 + pn1 + pnr + cn0 + cnr;
          const unsigned b = ac.decode(act, prob_sig(pr.sig, ctx), code, in);
          if (__any_sync(kFull, b)) {
            const int hh = contrib(cur0, neg0) + contrib(rv & 1u, (rv >> 1) & 1u);
            const int vv = contrib(u1s, u1n) + (int)(dp1 & 1u) - (int)(dn1 & 1u);
            const unsigned sc = sign_ctx(hh, vv);
            const unsigned sym = ac.decode(b != 0, prob_sgn(pr.sgn, sc >> 1), code, in);
            if (b) { cn1 = 1; neg1 = sym ^ (sc & 1u); }
          }
          nb1 |= cn1 << j; nn1 |= cn1 << j; ng1 |= (cn1 & neg1) << j;
        }
        const unsigned cur1 = (S1 & 1u) | cn1;
        unsigned lv = __shfl_up_sync(kFull, cn1, 1);
        if (lane == 0) lv = 0;
        pn0 = cn0; pn1 = cn1; pnl = lv; pnr = cnr;
        u0s = cur0; u0n = neg0; u1s = cur1; u1n = neg1;
        S0 >>= 1; S1 >>= 1; G0 >>= 1; G1 >>= 1;
        k00 >>= 1; k01 >>= 1; k02 >>= 1; k03 >>= 1; k10 >>= 1; k11 >>= 1; k12 >>= 1; k13 >>= 1;
        hp0 >>= 1; hn0 >>= 1; dp0 >>= 1; dn0 >>= 1; dp1 >>= 1; dn1 >>= 1;
      }
      B0 |= (u64)nb0 << sh; B1 |= (u64)nb1 << sh;
      N0 |= (u64)nn0 << sh; N1 |= (u64)nn1 << sh;
      g0 |= (u64)ng0 << sh; g1 |= (u64)ng1 << sh;
    }
    // ---- magnitude refinement pass (BPCEngine.cu:742-762,1267-1279) ----
    if (__any_sync(kFull, (s0 | s1) != 0)) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int sh = 32 * h;
        unsigned r0 = (unsigned)(s0 >> sh), r1 = (unsigned)(s1 >> sh), nb0 = 0, nb1 = 0;
#pragma unroll 2
        for (int j = 0; j < 32; ++j) {
          nb0 |= ac.decode(r0 & 1u, pr.ref, code, in) << j;
          nb1 |= ac.decode(r1 & 1u, pr.ref, code, in) << j;
          r0 >>= 1; r1 >>= 1;
        }
        B0 |= (u64)nb0 << sh; B1 |= (u64)nb1 << sh;
      }
    }
    s0 |= N0; s1 |= N1;
    if (bp < kSmemPlanes) {
      planes[wib][bp][lane] = make_uint4((unsigned)B0, (unsigned)(B0 >> 32), (unsigned)B1, (unsigned)(B1 >> 32));
    } else {
      for (int i = 0; i < kCB; ++i) {
        int2* d = reinterpret_cast<int2*>(dst + (size_t)i * p.Wa);
        int2 v = *d;
        v.x |= (int)((B0 >> i) & 1ull) << bp; v.y |= (int)((B1 >> i) & 1ull) << bp;
        *d = v;
      }
    }
  }
  __syncwarp();
  // assemble magnitudes from the bit planes and write (BPCEngine.cu:94-111)
  const int top = min(MSB, kSmemPlanes - 1);
  for (int h = 0; h < 2; ++h) {
    for (int j = 0; j < 32; ++j) {
      unsigned a = 0, b = 0;
      for (int bp = 0; bp <= top; ++bp) {
        const uint4 v = planes[wib][bp][lane];
        a |= (((h ? v.y : v.x) >> j) & 1u) << bp;
        b |= (((h ? v.w : v.z) >> j) & 1u) << bp;
      }
      const int i = 32 * h + j;
      int2* d = reinterpret_cast<int2*>(dst + (size_t)i * p.Wa);
      if (hi_planes) { const int2 o = *d; a |= (unsigned)o.x; b |= (unsigned)o.y; }
      const int va = (int)a, vb = (int)b;
      *d = make_int2(((g0 >> i) & 1ull) ? -va : va, ((g1 >> i) & 1ull) ? -vb : vb);
    }
  }
}

// Debug/parity export: the reference's int32 sparse `DCodeStreamValues` layout, -1 filled.
__global__ void export_ref_cs_kernel(const unsigned short* cw, const int* sizes, int* ref, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cb = (int)(i / kCBWords), j = (int)(i % kCBWords);
  ref[i] = j < sizes[cb] ? (int)cw[i] : -1;
}

}  // namespace

int bpc_encode(const void* d_coeff, bool is_float, uint16_t* d_cw, int32_t* d_sizes, const BpcArgs& a, cudaStream_t st) {
  const int blocks = (a.total_cb + 3) / 4;
  if (is_float) bpc_encode_kernel<float><<<blocks, 128, 0, st>>>(static_cast<const float*>(d_coeff), d_cw, d_sizes, a);
  else bpc_encode_kernel<int><<<blocks, 128, 0, st>>>(static_cast<const int*>(d_coeff), d_cw, d_sizes, a);
  ++g_launches;
  return (int)cudaGetLastError();
}

int bpc_decode(const uint16_t* d_cw, const int32_t* d_sizes, int32_t* d_q, const BpcArgs& a, cudaStream_t st) {
  const int blocks = (a.total_cb + 3) / 4;
  bpc_decode_kernel<<<blocks, 128, 0, st>>>(d_cw, d_sizes, d_q, a);
  ++g_launches;
  return (int)cudaGetLastError();
}

int export_ref_codestream(const uint16_t* d_cw, const int32_t* d_sizes, int32_t* d_ref, long long n, cudaStream_t st) {
  export_ref_cs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_cw, d_sizes, d_ref, n);
  ++g_launches;
  return (int)cudaGetLastError();
}

}  // namespace pcs
