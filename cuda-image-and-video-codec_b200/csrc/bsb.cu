// bsb.cu -- BitStreamBuilder for sm_100a: codeblock-length prefix scan + compaction, and its inverse.
//
// Replaces BitStreamBuilder::createBitStream / ::createCodeStream
// (BitStreamBuilder/BitStreamBuilder.cpp:100-153) and the device steps under them:
// cub::DeviceScan::InclusiveSum (BitStreamBuilder.cu:290-323), the 256-entry search index
// binarySearchLUTBSB (:198-265), the two memsets (:270-285) and the per-codeword binary-search copy
// buildBitStreamLUTBS / buildCodeStreamLUTBS (:33-171) -- four kernels, two memsets, a host-side
// size parse and four host synchronisations per frame in the reference.
//
// Here each direction is ONE single-pass kernel: a tile of 32 codeblocks per CTA, a decoupled
// look-back scan over tile aggregates (status word = flag|value; no CUB), then each warp copies its
// codeblocks' codewords to their final position  9 + 2n + (P[cb-1] - cb)  -- the closed form of the
// reference's search (SURVEY.md Annex C.2; oracle/picsong_oracle.c shows closed form == literal).
// The pack scan runs across ALL frames of a batch, so a batch is packed back to back by one launch,
// and every output word (header, per-block MSB/size pairs, codewords, trailing 0xFFFF) is written by
// the kernel itself: no memset, no host round trip for the total.
#include "pcs_common.cuh"

namespace pcs {

namespace {

constexpr int kTile = 32;  // codeblocks per CTA
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagInc = 2ull << 62, kValMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Decoupled look-back (warp 0 of the CTA).  Returns the sum of the aggregates of tiles [first, tile).
// status[t] bits 63:62: 0 = not yet published, 1 = aggregate only, 2 = inclusive prefix.  Tile ids are
// handed out by an atomic ticket, so a tile only ever waits for tiles that have already started.
__device__ __forceinline__ long long lookback(unsigned long long* status, int tile, int first, long long aggregate, int lane) {
  if (lane == 0) st_release(status + tile, (tile == first ? kFlagInc : kFlagAgg) | (unsigned long long)aggregate);
  long long excl = 0;
  if (tile > first) {
    int look = tile - 1;
    while (true) {
      const int t = look - lane;  // each lane inspects one predecessor
      unsigned long long s = kFlagInc;  // tiles before `first` contribute nothing and end the search
      if (t >= first) { do { s = ld_acquire(status + t); } while ((s >> 62) == 0); }
      const unsigned inc = __ballot_sync(0xffffffffu, (s >> 62) == 2);
      const int stop = inc ? __ffs(inc) - 1 : 32;  // nearest predecessor holding an inclusive prefix
      long long v = (lane <= stop && t >= first) ? (long long)(s & kValMask) : 0;
      for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      excl += v;
      if (inc) break;
      look -= 32;
    }
    if (lane == 0) st_release(status + tile, kFlagInc | (unsigned long long)(excl + aggregate));
  }
  return excl;
}

// Warp-cooperative copy of n uint16 between two arrays of arbitrary (2-byte) alignment.  The destination is written
// in 16-byte vectors (512 B per warp instruction instead of 64 B); the source is read as aligned 32-bit words and, when
// source and destination disagree in their alignment modulo 4, re-aligned with one PRMT per output word.  Reads stay
// inside [src, src + n) rounded outwards to 32-bit words only where the extra half word belongs to the same array:
// the caller guarantees src[n] is readable when it passes tail_ok (it is another element of the same buffer).
__device__ __forceinline__ void warp_copy_u16(unsigned short* __restrict__ dst, const unsigned short* __restrict__ src, int n, bool tail_ok,
                                              int lane) {
  const int head = min(n, (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) >> 1));
  for (int j = lane; j < head; j += 32) dst[j] = src[j];
  const unsigned short* const s = src + head;
  const int odd = (int)((reinterpret_cast<uintptr_t>(s) >> 1) & 1u);
  // an odd source reads one half word past the last element of a vector
  const int nvec = max(0, n - head - ((odd && !tail_ok) ? 1 : 0)) >> 3;
  const unsigned sel = odd ? 0x5432u : 0x3210u;
  const unsigned* const s32 = reinterpret_cast<const unsigned*>(s - odd);
  uint4* const d4 = reinterpret_cast<uint4*>(dst + head);
  for (int v = lane; v < nvec; v += 32) {
    const unsigned* const p = s32 + 4 * v;
    const unsigned w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3], w4 = odd ? p[4] : 0u;
    d4[v] = make_uint4(__byte_perm(w0, w1, sel), __byte_perm(w1, w2, sel), __byte_perm(w2, w3, sel), __byte_perm(w3, w4, sel));
  }
  for (int j = head + 8 * nvec + lane; j < n; j += 32) dst[j] = src[j];
}

// sparse uint16 [batch][ncb][4096] + sizes  ->  frames back to back (Annex C.2 layout per frame)
__global__ void __launch_bounds__(256) bsb_pack_kernel(const unsigned short* __restrict__ cw, const int* __restrict__ sizes,
                                                      unsigned short* __restrict__ out, long long* __restrict__ frame_off,
                                                      unsigned long long* status, unsigned* ticket, const BsbArgs a) {
  __shared__ int s_tile;
  __shared__ long long s_excl[kTile];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int n = a.ncb;
  const long long total_cb = (long long)n * a.batch;
  const long long cb0 = (long long)tile * kTile;
  if (warp == 0) {
    const long long g = cb0 + lane;
    const int sz = g < total_cb ? sizes[g] : 0;
    int incl = sz;
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    const long long base = lookback(status, tile, 0, __shfl_sync(0xffffffffu, incl, 31), lane);
    s_excl[lane] = base + incl - sz;  // E_g: sizes of all earlier codeblocks of the batch
  }
  __syncthreads();
  // codewords: word j (1 <= j < size) of block (f, cb) goes to  S_f + 9 + 2n + (P_f[cb-1] - cb) + j - 1  with
  // S_f = E_{f*n} + f*(n+10) and P_f[cb-1] = E_g - E_{f*n}; the frame start cancels:
  for (int k = warp; k < kTile; k += nwarps) {
    const long long g = cb0 + k;
    if (g >= total_cb) break;
    const int f = (int)(g / n), cb = (int)(g - (long long)f * n);
    const int sz = sizes[g];
    const long long data = s_excl[k] + (long long)f * (n + 10) + 9 + 2LL * n - cb;
    const unsigned short* src = cw + (size_t)g * a.cbw;
    // src[sz] is readable: sz < cbw leaves it inside the block, and a block of cbw words ends on a 32-bit boundary
    warp_copy_u16(out + data, src + 1, sz - 1, sz < a.cbw, lane);
  }
  // The tile holding block 0 of frame f knows S_f: it writes that frame's 9-word header, all n
  // (MSB, size) pairs (BitStreamBuilder.cu:122-127) and the previous frame's trailing 0xFFFF.
  for (int k = 0; k < kTile; ++k) {
    const long long g = cb0 + k;
    if (g >= total_cb) break;
    if (g % n) continue;
    const int f = (int)(g / n);
    const long long S = s_excl[k] + (long long)f * (n + 10);
    if (threadIdx.x == 0) { frame_off[f] = S; if (f > 0) out[S - 1] = 0xFFFFu; }
    if (threadIdx.x < 9) out[S + threadIdx.x] = (a.first_frame_index + f < a.hdr_count) ? a.hdr[threadIdx.x] : (unsigned short)0xFFFFu;
    for (int cb = threadIdx.x; cb < n; cb += blockDim.x) {
      out[S + 9 + 2 * cb] = cw[(size_t)(g + cb) * a.cbw];
      out[S + 10 + 2 * cb] = (unsigned short)sizes[g + cb];
    }
  }
  // the tile holding the last block knows the total
  if (threadIdx.x == 0 && total_cb - 1 >= cb0 && total_cb - 1 < cb0 + kTile) {
    const int k = (int)(total_cb - 1 - cb0);
    const long long total = s_excl[k] + sizes[total_cb - 1] + (long long)a.batch * (n + 10);
    frame_off[a.batch] = total;
    out[total - 1] = 0xFFFFu;  // one trailing 0xFFFF per frame (HTotalBSSize, BitStreamBuilder.cu:304-305)
  }
}

// frames at in + frame_off[f]  ->  sparse uint16 [batch][ncb][4096] + sizes (read from the stream itself;
// the reference parses them on the host, BitStreamBuilder.cpp:119-129).  Tiles never straddle frames.
__global__ void __launch_bounds__(256) bsb_unpack_kernel(const unsigned short* __restrict__ in, const long long* __restrict__ frame_off,
                                                        unsigned short* __restrict__ cw, int* __restrict__ sizes,
                                                        unsigned long long* status, unsigned* ticket, const BsbArgs a) {
  __shared__ int s_tile;
  __shared__ long long s_excl[kTile];
  __shared__ int s_sz[kTile];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int n = a.ncb, tpf = (n + kTile - 1) / kTile;
  const int tile = s_tile, f = tile / tpf, t = tile - f * tpf;
  const long long S = frame_off[f], E = frame_off[f + 1];
  const unsigned short* fin = in + S;
  const int cb0 = t * kTile;
  if (warp == 0) {
    const int cb = cb0 + lane;
    int sz = cb < n ? (int)fin[10 + 2 * cb] : 0;
    if (cb < n) sz = min(max(sz, 1), a.cbw);  // malformed streams cannot index outside a block
    int incl = sz;
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    const long long base = lookback(status, tile, f * tpf, __shfl_sync(0xffffffffu, incl, 31), lane);
    s_excl[lane] = base + incl - sz;  // P_f[cb-1]
    s_sz[lane] = sz;
  }
  __syncthreads();
  for (int k = warp; k < kTile; k += nwarps) {
    const int cb = cb0 + k;
    if (cb >= n) break;
    const int sz = s_sz[k];
    const long long g = (long long)f * n + cb;
    const long long data = 9 + 2LL * n + s_excl[k] - cb;
    unsigned short* dst = cw + (size_t)g * a.cbw;
    if (lane == 0) { dst[0] = fin[9 + 2 * cb]; sizes[g] = sz; }
    // words the (possibly malformed) frame really holds; the rest of the block reads as 0xFFFF
    const int have = (int)min((long long)(sz - 1), max(0LL, E - S - data));
    warp_copy_u16(dst + 1, fin + data, have, false, lane);
    for (int j = 1 + have + lane; j < sz; j += 32) dst[j] = (unsigned short)0xFFFFu;
  }
}

}  // namespace

int bsb_pack(const uint16_t* d_cw, const int32_t* d_sizes, uint16_t* d_out, long long* d_frame_off, void* d_status,
             const BsbArgs& a, cudaStream_t st) {
  const long long total_cb = (long long)a.ncb * a.batch;
  const int tiles = (int)((total_cb + kTile - 1) / kTile);
  unsigned long long* status = static_cast<unsigned long long*>(d_status);
  cudaError_t e = cudaMemsetAsync(d_status, 0, (size_t)(tiles + 1) * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return (int)e;
  bsb_pack_kernel<<<tiles, 256, 0, st>>>(d_cw, d_sizes, d_out, d_frame_off, status + 1, reinterpret_cast<unsigned*>(status), a);
  ++g_launches;
  return (int)cudaGetLastError();
}

int bsb_unpack(const uint16_t* d_in, const long long* d_frame_off, uint16_t* d_cw, int32_t* d_sizes, void* d_status,
               const BsbArgs& a, cudaStream_t st) {
  const int tpf = (a.ncb + kTile - 1) / kTile, tiles = tpf * a.batch;
  unsigned long long* status = static_cast<unsigned long long*>(d_status);
  cudaError_t e = cudaMemsetAsync(d_status, 0, (size_t)(tiles + 1) * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return (int)e;
  bsb_unpack_kernel<<<tiles, 256, 0, st>>>(d_in, d_frame_off, d_cw, d_sizes, status + 1, reinterpret_cast<unsigned*>(status), a);
  ++g_launches;
  return (int)cudaGetLastError();
}

size_t bsb_status_bytes(int ncb, int batch) {
  const long long tiles = ((long long)ncb + kTile - 1) / kTile * batch + 2;
  return (size_t)tiles * sizeof(unsigned long long);
}

}  // namespace pcs
