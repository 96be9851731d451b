// api.cu -- context, workspace and the extern "C" entry points of libpicsong_b200.so.
//
// Replaces the orchestration the reference does in Engines/{Engine,CodingEngine,DecodingEngine}.cu:
// LUT upload (Engine.cu:8-185), buffer allocation (CodingEngine.cu:119-199, DecodingEngine.cu:140-235)
// and the per-frame stage order of runImage/runVideo (CodingEngine.cu:634-674,758-983;
// DecodingEngine.cu:770-859,866-1043) -- minus its per-frame new/delete of stage objects, the
// memsets, and the host synchronisation after every kernel.  A batch of frames goes through
// DWT -> BPC -> bit-stream builder as 7 launches on one stream with no host round trip.
//
// RGB (CodingEngine.cu:598-630,680-710,764-810; DecodingEngine.cu:738-770): a frame is three
// colour-transformed component planes, each coded as an independent "frame" with its own probability table.
// Everything below works on PLANES = frames x components; plane 3f+c is component c of frame f.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pcs_common.cuh"

namespace pcs {

unsigned long long g_launches = 0;

size_t dwt_scratch_elems(int Wa, int Ha, int wl);
int dwt_forward(const uint8_t*, void*, void*, int, int, int, int, int, bool, float, bool, int, cudaStream_t, int offset);
int dwt_forward_planes(const void*, void*, void*, int, int, int, bool, float, bool, int, cudaStream_t);
bool dwt_forward_rgb_ok(const uint8_t*, const void*, const void*, int, int, int, int, int);
int dwt_forward_rgb(const uint8_t*, void*, void*, int, int, int, int, int, bool, float, bool, int, cudaStream_t);
int dwt_inverse_planes(const int32_t*, void*, void*, int, int, int, bool, float, int, cudaStream_t);
int rgb_forward(const uint8_t*, void*, int, int, int, int, bool, int, cudaStream_t);
int rgb_inverse(const void*, uint8_t*, int, int, int, int, bool, bool, int, cudaStream_t);
bool dwt_inverse_rgb_ok(const int32_t*, const uint8_t*, const void*, int, int, int, int, int);
int dwt_inverse_rgb(const int32_t*, uint8_t*, void*, int, int, int, int, int, bool, float, int, cudaStream_t);
int dwt_inverse(const int32_t*, uint8_t*, void*, int, int, int, int, int, bool, float, bool, int, cudaStream_t, int offset);
int bpc_encode(const void*, bool, uint16_t*, int32_t*, const BpcArgs&, cudaStream_t);
int bpc_decode(const uint16_t*, const int32_t*, int32_t*, const BpcArgs&, cudaStream_t);
int bpc_encode_k(const void*, bool, uint16_t*, int32_t*, const BpcKArgs&, cudaStream_t);
int bpc_decode_k(const uint16_t*, const int32_t*, int32_t*, const BpcKArgs&, cudaStream_t);
int export_ref_codestream(const uint16_t*, const int32_t*, int32_t*, long long, int, cudaStream_t);
int bsb_pack(const uint16_t*, const int32_t*, uint16_t*, long long*, void*, const BsbArgs&, cudaStream_t);
int bsb_unpack(const uint16_t*, const long long*, uint16_t*, int32_t*, void*, const BsbArgs&, cudaStream_t);
size_t bsb_status_bytes(int ncb, int batch);

}  // namespace pcs

using namespace pcs;

struct pcs_ctx {
  pcs_params p;
  int Wa, Ha, ncb, ncomp;          // ncomp: planes per frame (1 grey, 3 RGB)
  int cb_rows = 64;                // codeblock length: 64, or 32 with PCS_F_CB_LENGTH_32
  int offset = 128;                // level shift of the grey path: 2^(bit_depth - 1), 0 for signed samples
  size_t A, scratch_elems, max_words, frame_px;  // max_words: per PLANE; frame_px = ncomp * W * H
  std::vector<int32_t> lut;
  uint4* d_tab4 = nullptr; int nrows = 0;  // device: probability rows, [ncomp][nrows]
  // device: workspace, sized for max_batch frames = max_batch * ncomp planes
  void* d_scratch = nullptr;       // DWT LL planes, [planes][scratch_elems] x 4 bytes
  void* d_status = nullptr;        // look-back status words
  uint8_t* d_px = nullptr;         // [max_batch][ncomp][W*H]
  void* d_tplanes = nullptr;       // RGB only: colour-transformed 32-bit planes [planes][A]
  int32_t* d_coeff = nullptr;      // [planes][A]
  uint16_t* d_cw = nullptr;        // [planes][ncb][4096]
  int32_t* d_sizes = nullptr;      // [planes][ncb]
  uint16_t* d_stream = nullptr;    // [planes * max_words]
  long long* d_off = nullptr;      // [planes + 1]
  long long* h_off = nullptr;      // pinned, [kSlots][planes + 1]
  // host-buffer pipeline (pcs_encode_frames / pcs_decode_frames): the workspace is cut into kSlots slots,
  // each with its own stream, so that the H2D copy of one chunk, the kernels of another and the D2H copy of a
  // third overlap (PCIe is full duplex; copy engines run beside the SMs)
  static constexpr int kSlots = 4;
  cudaStream_t sstream[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t sev[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t kev[kSlots] = {nullptr, nullptr, nullptr, nullptr};  // "kernels of the slot's chunk are done"
  long long* d_soff = nullptr;     // [kSlots][planes + 1]
  void* d_sstatus = nullptr;       // [kSlots] look-back status areas
  size_t status_bytes = 0;
  int32_t* d_lutk = nullptr;       // k > 0: flat tables of all bit-plane files, [ncomp][files][per]
  int lut_files = 1;               // table sets per component in `lut` (1 for k == 0)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool timing = false;
  std::string err;
};

namespace {

int fail(pcs_ctx* c, int code, const char* msg) {
  if (c) c->err = msg;
  return code;
}

BpcArgs bpc_args(const pcs_ctx* c, int frames, int comp) {
  BpcArgs a{};
  a.Wa = c->Wa; a.Ha = c->Ha; a.nbx = c->Wa / kCB; a.ncb = c->ncb; a.wl = c->p.wavelet_levels;
  a.nBp = c->p.lut_header[0]; a.nSub = c->p.lut_header[1]; a.nrows = c->nrows; a.prec = c->p.lut_header[5];
  a.total_cb = c->ncb * frames; a.A = (long long)c->A;
  a.pstep = c->ncomp; a.pfirst = comp; a.cb_rows = c->cb_rows;
  a.tab4 = c->d_tab4 + (size_t)comp * c->nrows;
  return a;
}

// `planes` consecutive planes starting at global plane index `first_plane`
BsbArgs bsb_args(const pcs_ctx* c, int planes, int first_plane) {
  BsbArgs a{};
  a.ncb = c->ncb; a.batch = planes; a.first_frame_index = first_plane; a.cbw = kCB * c->cb_rows;
  // stills: component i is coded with iter = i, video: every component of frame f with iter = f
  // (CodingEngine.cu:620,700 vs :790,902) -> the three planes of video frame 0 all carry the header
  // (PCS_F_STILLS: the caller codes stills although `frames`, a header-only value, is > 0)
  a.hdr_count = (c->ncomp == 3 && c->p.frames > 0 && !(c->p.flags & PCS_F_STILLS)) ? 3 : 1;
  uint16_t h[9];
  pcs_header_pack(&c->p, h);
  for (int i = 0; i < 9; ++i) a.hdr[i] = h[i];
  return a;
}

// Packs the flat int32 tables (ref | sig | sign sections, row = (lvl*nSub+sb)*nBp + bp) of every component
// into 16-byte rows: bytes 0..8 significance contexts, 9 refinement, 12..15 sign contexts.  Rows past the
// table keep the reference's unguarded flat indexing (BPCEngine.cu:329-350: MSB >= nBp walks into the next
// group / section); beyond the table -> 0.
int build_prob_tables(pcs_ctx* c) {
  const int32_t* h = c->p.lut_header;
  const int nBp = h[0], nSub = h[1], cR = h[2], cG = h[3], cS = h[4], wl = c->p.wavelet_levels;
  if (cR != 1 || cS < 1 || cS > 9 || cG < 1 || cG > 4 || nBp < 1 || nBp > 32 || nSub != 3 || h[5] < 1 || h[5] > 15)
    return PCS_E_LUT_RANGE;
  const int R = (nSub * wl + 1) * nBp * cR, S = (nSub * wl + 1) * nBp * cS;
  const int n = R + S + (nSub * wl + 1) * nBp * cG;  // entries per component
  if ((long long)c->lut.size() < (long long)n * c->lut_files * c->ncomp) return PCS_E_LUT;
  const int nrows = (nSub * wl + 1) * nBp + 32;
  if (nrows > kMaxProbRows) return PCS_E_LUT_RANGE;  // e.g. nBp = 16 at wl = 10: more rows than the decoder stages
  std::vector<uint4> tab4((size_t)nrows * c->ncomp, make_uint4(0, 0, 0, 0));
  for (int comp = 0; comp < c->ncomp; ++comp) {
    const int32_t* lut = c->lut.data() + (size_t)comp * n * c->lut_files;  // set 0 of the component (the k == 0 tables)
    auto at = [&](long long i) -> int { return (i >= 0 && i < n) ? lut[(size_t)i] : 0; };
    for (int r = 0; r < nrows; ++r) {
      uint4& t = tab4[(size_t)comp * nrows + r];
      for (int k = 0; k < 9; ++k) {
        const int v = k < cS ? at((long long)R + (long long)r * cS + k) : 0;
        if (v < 0 || v > 127) return PCS_E_LUT_RANGE;
        if (k < 4) t.x |= (unsigned)v << (8 * k);
        else if (k < 8) t.y |= (unsigned)v << (8 * (k - 4));
        else t.z |= (unsigned)v;
      }
      for (int k = 0; k < 4; ++k) {
        const int v = k < cG ? at((long long)R + S + (long long)r * cG + k) : 0;
        if (v < 0 || v > 127) return PCS_E_LUT_RANGE;
        t.w |= (unsigned)v << (8 * k);
      }
      const int v = at((long long)r * cR);
      if (v < 0 || v > 127) return PCS_E_LUT_RANGE;
      t.z |= (unsigned)v << 8;
    }
  }
  c->nrows = nrows;
  PCS_CUDA_TRY(cudaMalloc(&c->d_tab4, tab4.size() * sizeof(uint4)));
  PCS_CUDA_TRY(cudaMemcpy(c->d_tab4, tab4.data(), tab4.size() * sizeof(uint4), cudaMemcpyHostToDevice));
  if (c->p.k_factor > 0) {  // the complexity-scalable kernels index the flat tables the way the reference does
    const size_t total = (size_t)n * c->lut_files * c->ncomp;
    PCS_CUDA_TRY(cudaMalloc(&c->d_lutk, total * sizeof(int32_t)));
    PCS_CUDA_TRY(cudaMemcpy(c->d_lutk, c->lut.data(), total * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  return PCS_OK;
}

// ---- the stage chain on explicit workspace pointers (used by the device entry points and by the slots) ----
struct Ws {
  void* scratch; void* status; void* tplanes; int32_t* coeff; uint16_t* cw; int32_t* sizes;
};

int run_dwt_forward(pcs_ctx* c, const Ws& w, const uint8_t* d_px, void* d_coeff, bool out_int, int frames, cudaStream_t st) {
  const pcs_params& p = c->p;
  if (c->ncomp == 1)
    return dwt_forward(d_px, d_coeff, w.scratch, p.width, p.height, c->Wa, c->Ha, p.wavelet_levels, p.lossy != 0, p.quant_size, out_int,
                       frames, st, c->offset);
  // colour transform fused into the level-1 load where the streaming kernel applies ...
  static const bool fuse_rgb = [] { const char* e = std::getenv("PCS_RGB_FUSE"); return !e || std::atoi(e) != 0; }();
  if (fuse_rgb && dwt_forward_rgb_ok(d_px, d_coeff, w.scratch, p.width, p.height, c->Wa, c->Ha, p.wavelet_levels))
    return dwt_forward_rgb(d_px, d_coeff, w.scratch, p.width, p.height, c->Wa, c->Ha, p.wavelet_levels, p.lossy != 0, p.quant_size,
                           out_int, frames * 3, st);
  // ... else as a pass of its own into 32-bit planes
  int rc = rgb_forward(d_px, w.tplanes, p.width, p.height, c->Wa, c->Ha, p.lossy != 0, frames, st);
  if (rc) return rc;
  return dwt_forward_planes(w.tplanes, d_coeff, w.scratch, c->Wa, c->Ha, p.wavelet_levels, p.lossy != 0, p.quant_size, out_int,
                            frames * 3, st);
}

int run_dwt_inverse(pcs_ctx* c, const Ws& w, const int32_t* d_q, uint8_t* d_px, int frames, cudaStream_t st) {
  const pcs_params& p = c->p;
  const bool flat = (p.flags & PCS_F_REF_FLAT_CROP) != 0;
  if (c->ncomp == 1)
    return dwt_inverse(d_q, d_px, w.scratch, p.width, p.height, c->Wa, c->Ha, p.wavelet_levels, p.lossy != 0, p.quant_size, flat, frames,
                       st, c->offset);
  static const bool fuse_rgb = [] { const char* e = std::getenv("PCS_RGB_FUSE"); return !e || std::atoi(e) != 0; }();
  if (fuse_rgb && dwt_inverse_rgb_ok(d_q, d_px, w.scratch, p.width, p.height, c->Wa, c->Ha, p.wavelet_levels))
    return dwt_inverse_rgb(d_q, d_px, w.scratch, p.width, p.height, c->Wa, c->Ha, p.wavelet_levels, p.lossy != 0, p.quant_size, frames * 3,
                           st);
  int rc = dwt_inverse_planes(d_q, w.tplanes, w.scratch, c->Wa, c->Ha, p.wavelet_levels, p.lossy != 0, p.quant_size, frames * 3, st);
  if (rc) return rc;
  return rgb_inverse(w.tplanes, d_px, p.width, p.height, c->Wa, c->Ha, p.lossy != 0, flat, frames, st);
}

BpcKArgs bpc_k_args(const pcs_ctx* c, int frames, int comp) {
  BpcKArgs a{};
  const int32_t* h = c->p.lut_header;
  a.Wa = c->Wa; a.Ha = c->Ha; a.nbx = c->Wa / kCB; a.ncb = c->ncb; a.wl = c->p.wavelet_levels;
  a.nBp = h[0]; a.nSub = h[1]; a.cR = h[2]; a.cG = h[3]; a.cS = h[4]; a.prec = h[5];
  a.total_cb = c->ncb * frames; a.A = (long long)c->A;
  a.pstep = c->ncomp; a.pfirst = comp;
  a.per = pcs_lut_size(h, a.wl); a.lutN = a.per * c->lut_files;
  a.k = c->p.k_factor;
  a.lut = c->d_lutk + (size_t)comp * a.lutN;
  return a;
}

int run_bpc_encode(pcs_ctx* c, const void* d_coeff, bool is_float, uint16_t* d_cw, int32_t* d_sizes, int frames, cudaStream_t st) {
  if (c->p.k_factor > 0) {  // complexity-scalable mode: the lane-per-column-pair kernels of bpc.cu
    for (int comp = 0; comp < c->ncomp; ++comp) {
      int rc = bpc_encode_k(d_coeff, is_float, d_cw, d_sizes, bpc_k_args(c, frames, comp), st);
      if (rc) return rc;
    }
    return PCS_OK;
  }
  for (int comp = 0; comp < c->ncomp; ++comp) {  // one launch per component: each has its own probability table
    int rc = bpc_encode(d_coeff, is_float, d_cw, d_sizes, bpc_args(c, frames, comp), st);
    if (rc) return rc;
  }
  return PCS_OK;
}

int run_bpc_decode(pcs_ctx* c, const uint16_t* d_cw, const int32_t* d_sizes, int32_t* d_q, int frames, cudaStream_t st) {
  if (c->p.k_factor > 0) {
    for (int comp = 0; comp < c->ncomp; ++comp) {
      int rc = bpc_decode_k(d_cw, d_sizes, d_q, bpc_k_args(c, frames, comp), st);
      if (rc) return rc;
    }
    return PCS_OK;
  }
  for (int comp = 0; comp < c->ncomp; ++comp) {
    int rc = bpc_decode(d_cw, d_sizes, d_q, bpc_args(c, frames, comp), st);
    if (rc) return rc;
  }
  return PCS_OK;
}

}  // namespace

extern "C" {

int32_t pcs_create(const pcs_params* params, pcs_ctx** out) {
  if (!params || !out) return PCS_E_ARG;
  *out = nullptr;
  const pcs_params& p = *params;
  // Launcher.cu:132: "Incorrect parameters. Please choose valid values."
  if (p.width <= 0 || p.height <= 0 || p.wavelet_levels < 1 || p.wavelet_levels > 10 || p.quant_size < 0 ||
      p.quant_size > 1 || p.cb_width % 64 != 0 || p.cb_height < 18 || p.cb_height > 20 || p.coding_passes < 2 ||
      p.coding_passes > 3 || p.k_factor < 0 || p.k_factor > 65.535f || p.max_batch < 1)
    return PCS_E_ARG;
  if (p.lossy && !(p.quant_size > 0)) return PCS_E_ARG;
  if (p.coding_passes != 2) return PCS_E_UNSUPPORTED;  // the 3-pass variant needs LUT files the reference does not ship
  // -bps / -signedOrUnsigned: samples stay BYTES (IO/IOManager.ipp:72-112 reads unsigned char whatever -bps
  // says), only the level shift changes: 2^(bps-1), or 0 for signed input (Engines/CodingEngine.cu:581-588,644-656); the
  // decoder clamps to [0, 255] either way (Engines/DecodingEngine.cu:706-729).  Grey only; bounded so that the shifted
  // samples stay exact in fp32 and the LL indices inside the de-quantiser's 2^21 range.
  if (p.bit_depth < 1 || p.bit_depth > 16) return PCS_E_ARG;
  const int offset = p.is_signed ? 0 : 1 << (p.bit_depth - 1);
  if (offset != 128 && p.is_rgb) return PCS_E_UNSUPPORTED;
  if (p.lossy && (double)(offset + 128) * 1.1 * (double)(1 << p.wavelet_levels) >= 2097152.0) return PCS_E_UNSUPPORTED;
  // RGB: `-isRGB 1` with the default `-components 1` is what the reference's own examples use (README.md:105-115);
  // `-components 3` is accepted as well.  Grey: one component.
  if (p.is_rgb ? (p.components != 1 && p.components != 3) : p.components != 1) return PCS_E_UNSUPPORTED;
  if (!p.lut || p.lut_len <= 0) return PCS_E_LUT;
  // codeblock length 32 = the reference rebuilt with -DCBLOCK_LENGTH=32 (BPC/BPCEngine.cuh:33-36): the frame is padded to a
  // multiple of 32 rows instead of 64 (fixImageProportions is called with CBLOCK_LENGTH).  k = 0 coder only.
  const int cb_rows = (p.flags & PCS_F_CB_LENGTH_32) ? 32 : 64;
  if (cb_rows == 32 && p.k_factor > 0) return PCS_E_UNSUPPORTED;
  int Wa, Ha;
  pcs_adapted_dims_cb(p.width, p.height, cb_rows, &Wa, &Ha);
  // every level input must have even size
  if (((Wa >> (p.wavelet_levels - 1)) & 1) || ((Ha >> (p.wavelet_levels - 1)) & 1) || (Wa >> p.wavelet_levels) < 1 ||
      (Ha >> p.wavelet_levels) < 1)
    return PCS_E_ARG;
  pcs_ctx* c = new (std::nothrow) pcs_ctx();
  if (!c) return PCS_E_NOMEM;
  c->p = p;
  c->lut.assign(p.lut, p.lut + p.lut_len);
  c->p.lut = c->lut.data();
  c->ncomp = p.is_rgb ? 3 : 1;
  c->offset = offset;
  c->cb_rows = cb_rows;
  c->Wa = Wa; c->Ha = Ha; c->A = (size_t)Wa * Ha; c->ncb = (Wa / kCB) * (Ha / cb_rows);
  c->max_words = c->A + c->ncb + 10;
  c->frame_px = (size_t)c->ncomp * p.width * p.height;
  c->scratch_elems = dwt_scratch_elems(Wa, Ha, p.wavelet_levels);
  const char* t = std::getenv("PCS_STAGE_TIMING");
  c->timing = t && t[0] == '1';
  // k > 0: `lut` holds the tables of ALL bit-plane files of every component (pcs_lut_load_k), header[7] sets each
  c->lut_files = p.k_factor > 0 ? p.lut_header[7] : 1;
  if (c->lut_files < 1 || c->lut_files > 32) { delete c; return PCS_E_LUT_RANGE; }
  int rc = build_prob_tables(c);
  if (rc != PCS_OK) { pcs_destroy(c); return rc; }
  const size_t B = (size_t)p.max_batch, PL = B * c->ncomp;  // frames, planes
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 16); };
  alloc(&c->d_scratch, PL * (c->scratch_elems + 4) * 4);
  alloc(&c->d_status, bsb_status_bytes(c->ncb, (int)PL));
  alloc((void**)&c->d_px, B * c->frame_px + 16);
  if (c->ncomp == 3) alloc(&c->d_tplanes, PL * c->A * 4);
  alloc((void**)&c->d_coeff, PL * c->A * 4);
  alloc((void**)&c->d_cw, PL * c->A * 2);
  alloc((void**)&c->d_sizes, PL * c->ncb * 4);
  alloc((void**)&c->d_stream, PL * c->max_words * 2);
  alloc((void**)&c->d_off, (PL + 1) * sizeof(long long));
  c->status_bytes = (bsb_status_bytes(c->ncb, (int)PL) + 255) / 256 * 256;
  alloc(&c->d_sstatus, pcs_ctx::kSlots * c->status_bytes);
  alloc((void**)&c->d_soff, pcs_ctx::kSlots * (PL + 1) * sizeof(long long));
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&c->h_off, pcs_ctx::kSlots * (PL + 1) * sizeof(long long), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  for (int i = 0; i < pcs_ctx::kSlots && e == cudaSuccess; ++i) {
    e = cudaStreamCreateWithFlags(&c->sstream[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->sev[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->kev[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
  if (e != cudaSuccess) { pcs_destroy(c); return (int)e; }
  *out = c;
  return PCS_OK;
}

void pcs_destroy(pcs_ctx* c) {
  if (!c) return;
  cudaFree(c->d_tab4);
  cudaFree(c->d_scratch); cudaFree(c->d_status); cudaFree(c->d_px); cudaFree(c->d_tplanes); cudaFree(c->d_coeff); cudaFree(c->d_cw);
  cudaFree(c->d_sizes); cudaFree(c->d_stream); cudaFree(c->d_off);
  if (c->h_off) cudaFreeHost(c->h_off);
  cudaFree(c->d_sstatus); cudaFree(c->d_soff); cudaFree(c->d_lutk);
  for (auto& s2 : c->sstream) if (s2) cudaStreamDestroy(s2);
  for (auto& e2 : c->sev) if (e2) cudaEventDestroy(e2);
  for (auto& e2 : c->kev) if (e2) cudaEventDestroy(e2);
  if (c->stream) cudaStreamDestroy(c->stream);
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  delete c;
}

// Pinned host memory for callers that stream frames from / to disk (the reference allocates its frame buffers
// with cudaHostAlloc as well: Engines/CodingEngine.cu:129-131,289): copies from pinned buffers are truly
// asynchronous, which is what lets pcs_encode_frames / pcs_decode_frames overlap PCIe with the kernels.
void* pcs_host_alloc(size_t bytes) {
  void* p = nullptr;
  return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
void pcs_host_free(void* p) { if (p) cudaFreeHost(p); }

const char* pcs_last_error(const pcs_ctx* c) { return c ? c->err.c_str() : "no context"; }
size_t pcs_frame_pixels(const pcs_ctx* c) { return c->A; }
int32_t pcs_frame_codeblocks(const pcs_ctx* c) { return c->ncb; }
size_t pcs_frame_max_stream_words(const pcs_ctx* c) { return c->max_words * c->ncomp; }
int32_t pcs_frame_planes(const pcs_ctx* c) { return c->ncomp; }
uint64_t pcs_kernel_launch_count(void) { return g_launches; }

#define CHECK_BATCH(c, batch) \
  if (!(c)) return PCS_E_ARG; \
  if ((batch) < 1 || (batch) > (c)->p.max_batch) return fail((c), PCS_E_BATCH, "batch exceeds max_batch")
#define FULL_WS(c) Ws{(c)->d_scratch, (c)->d_status, (c)->d_tplanes, (c)->d_coeff, (c)->d_cw, (c)->d_sizes}

int32_t pcs_dwt_forward(pcs_ctx* c, const uint8_t* d_px, void* d_coeff, int32_t trunc, int32_t batch, void* st) {
  CHECK_BATCH(c, batch);
  return run_dwt_forward(c, FULL_WS(c), d_px, d_coeff, trunc != 0, batch, (cudaStream_t)st);
}

int32_t pcs_dwt_inverse(pcs_ctx* c, const int32_t* d_q, uint8_t* d_px, int32_t batch, void* st) {
  CHECK_BATCH(c, batch);
  return run_dwt_inverse(c, FULL_WS(c), d_q, d_px, batch, (cudaStream_t)st);
}

int32_t pcs_bpc_encode(pcs_ctx* c, const void* d_coeff, int32_t is_float, uint16_t* d_cw, int32_t* d_sizes, int32_t batch, void* st) {
  CHECK_BATCH(c, batch);
  return run_bpc_encode(c, d_coeff, is_float != 0, d_cw, d_sizes, batch, (cudaStream_t)st);
}

int32_t pcs_bpc_decode(pcs_ctx* c, const uint16_t* d_cw, const int32_t* d_sizes, int32_t* d_q, int32_t batch, void* st) {
  CHECK_BATCH(c, batch);
  return run_bpc_decode(c, d_cw, d_sizes, d_q, batch, (cudaStream_t)st);
}

int32_t pcs_bsb_pack(pcs_ctx* c, const uint16_t* d_cw, const int32_t* d_sizes, int32_t first, uint16_t* d_out, int64_t* d_off,
                     int32_t batch, void* st) {
  CHECK_BATCH(c, batch);
  return bsb_pack(d_cw, d_sizes, d_out, reinterpret_cast<long long*>(d_off), c->d_status, bsb_args(c, batch * c->ncomp, first * c->ncomp),
                  (cudaStream_t)st);
}

int32_t pcs_bsb_unpack(pcs_ctx* c, const uint16_t* d_in, const int64_t* d_off, uint16_t* d_cw, int32_t* d_sizes, int32_t batch, void* st) {
  CHECK_BATCH(c, batch);
  return bsb_unpack(d_in, reinterpret_cast<const long long*>(d_off), d_cw, d_sizes, c->d_status, bsb_args(c, batch * c->ncomp, 1),
                    (cudaStream_t)st);
}

int32_t pcs_export_ref_codestream(pcs_ctx* c, const uint16_t* d_cw, const int32_t* d_sizes, int32_t* d_ref, int32_t batch, void* st) {
  CHECK_BATCH(c, batch);
  return export_ref_codestream(d_cw, d_sizes, d_ref, (long long)c->A * batch * c->ncomp, kCB * c->cb_rows, (cudaStream_t)st);
}

int32_t pcs_encode_device(pcs_ctx* c, const uint8_t* d_px, int32_t batch, int32_t first, uint16_t* d_stream, int64_t* d_off, void* stv) {
  CHECK_BATCH(c, batch);
  cudaStream_t st = (cudaStream_t)stv;
  if (c->timing) cudaEventRecord(c->ev[0], st);
  int rc = pcs_dwt_forward(c, d_px, c->d_coeff, 1, batch, st);  // 9/7: quantisation indices (int)c straight from the DWT
  if (rc) return rc;
  if (c->timing) cudaEventRecord(c->ev[1], st);
  rc = pcs_bpc_encode(c, c->d_coeff, 0, c->d_cw, c->d_sizes, batch, st);
  if (rc) return rc;
  if (c->timing) cudaEventRecord(c->ev[2], st);
  rc = pcs_bsb_pack(c, c->d_cw, c->d_sizes, first, d_stream, d_off, batch, st);
  if (c->timing) cudaEventRecord(c->ev[3], st);
  return rc;
}

int32_t pcs_decode_device(pcs_ctx* c, const uint16_t* d_stream, const int64_t* d_off, int32_t batch, uint8_t* d_px, void* stv) {
  CHECK_BATCH(c, batch);
  cudaStream_t st = (cudaStream_t)stv;
  if (c->timing) cudaEventRecord(c->ev[0], st);
  int rc = pcs_bsb_unpack(c, d_stream, d_off, c->d_cw, c->d_sizes, batch, st);
  if (rc) return rc;
  if (c->timing) cudaEventRecord(c->ev[1], st);
  rc = pcs_bpc_decode(c, c->d_cw, c->d_sizes, c->d_coeff, batch, st);
  if (rc) return rc;
  if (c->timing) cudaEventRecord(c->ev[2], st);
  rc = pcs_dwt_inverse(c, c->d_coeff, d_px, batch, st);
  if (c->timing) cudaEventRecord(c->ev[3], st);
  return rc;
}

int32_t pcs_last_stage_ms(pcs_ctx* c, float ms[3]) {
  if (!c || !c->timing) return PCS_E_ARG;
  PCS_CUDA_TRY(cudaEventSynchronize(c->ev[3]));
  for (int i = 0; i < 3; ++i) PCS_CUDA_TRY(cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
  return PCS_OK;
}

namespace {

// One slot of the workspace: frames [k*sb, (k+1)*sb) of every per-frame buffer.
struct Slot {
  Ws w; uint8_t* px; uint16_t* stream; long long* d_off; long long* h_off; cudaStream_t st; cudaEvent_t ev;
};
Slot make_slot(pcs_ctx* c, int k, int sb) {
  const size_t f0 = (size_t)k * sb, p0 = f0 * c->ncomp, PL1 = (size_t)c->p.max_batch * c->ncomp + 1;
  Slot s;
  s.w.scratch = static_cast<char*>(c->d_scratch) + p0 * c->scratch_elems * 4;
  s.w.status = static_cast<char*>(c->d_sstatus) + (size_t)k * c->status_bytes;
  s.w.tplanes = c->d_tplanes ? static_cast<char*>(c->d_tplanes) + p0 * c->A * 4 : nullptr;
  s.w.coeff = c->d_coeff + p0 * c->A;
  s.w.cw = c->d_cw + p0 * c->A;
  s.w.sizes = c->d_sizes + p0 * c->ncb;
  s.px = c->d_px + f0 * c->frame_px;
  s.stream = c->d_stream + p0 * c->max_words;
  s.d_off = c->d_soff + (size_t)k * PL1;
  s.h_off = c->h_off + (size_t)k * PL1;
  s.st = c->sstream[k]; s.ev = c->sev[k];
  return s;
}
// How many chunks may have kernels on the GPU at once.  The slot streams would otherwise share the SMs evenly, all
// chunks in flight would finish together and their downloads would queue up behind each other at the end of the
// call; with a depth of 2 a chunk still fills the partial last waves of its predecessor, but chunks complete in
// order, one download apart, and uploads run ahead of the kernels (they are enqueued before the wait).
// A call of no more chunks than slots has no such tail to speak of and runs them all at once (measured: RGB stills,
// 4 chunks: +5 %).
int kernel_depth(int nslots, int nchunks) {
  static const int env = [] { const char* e = std::getenv("PCS_PIPE_DEPTH"); return e ? std::atoi(e) : 0; }();
  const int d = env > 0 ? env : (nchunks <= nslots ? nslots : 2);
  return std::max(1, std::min(d, nslots));
}
int slots_for(const pcs_ctx* c, int* sb) {
  // chunks of max_batch / (2 * slots) frames: fine enough that the first upload and the last download (the
  // parts of a call that cannot overlap anything) are short, coarse enough that a launch still fills the GPU.
  // Small frames: a chunk carries at least 8 MB of pixels (a 512x512 batch goes through as one chunk -- ten
  // launches and two copies per chunk cost more than overlapping a 4 MB upload could ever gain).
  const size_t min_chunk = 8u << 20;
  const int B = c->p.max_batch;
  const int sb_min = (int)std::min<size_t>((size_t)B, (min_chunk + c->frame_px - 1) / c->frame_px);
  int n = std::max(1, std::min<int>(pcs_ctx::kSlots, B / sb_min));
  // ... and no more than ~32 MB: measured on 32 x 16.7 MB frames, chunks of 1 / 2 / 4 / 8 frames give 19.4 / 20.1 / 19.3 /
  // 16.6 Gpix/s (longer chunks lengthen the un-overlapped first upload and last download, shorter ones pay more launches)
  const int sb_big = (int)std::max<size_t>(1, ((32u << 20) + c->frame_px - 1) / c->frame_px);
  *sb = std::min(std::max(sb_min, std::min(sb_big, std::max(1, B / (2 * n)))), B / n);
  static const int env = [] { const char* e = std::getenv("PCS_PIPE_CHUNK"); return e ? std::atoi(e) : 0; }();
  if (env > 0) { n = std::min<int>(pcs_ctx::kSlots, B); *sb = std::max(1, std::min(env, B / n)); }
  return n;
}

// Early error returns must not leave work queued on the slot streams (the next call would reuse the slots under it).
struct SlotDrain {
  pcs_ctx* c; int n; bool armed = true;
  ~SlotDrain() { if (armed) for (int k = 0; k < n; ++k) cudaStreamSynchronize(c->sstream[k]); }
};

// The chunks of a call: `sb` frames each, except that the first two and the last two are halved -- the upload of the
// first chunk and the download of the last one are the parts of a call nothing else can overlap.
struct Chunk { int f0, b; };
std::vector<Chunk> chunk_plan(int nframes, int sb) {
  static const int taper = [] { const char* e = std::getenv("PCS_PIPE_TAPER"); return e ? std::atoi(e) : 1; }();
  std::vector<Chunk> plan;
  const int small = std::max(1, sb / 2);
  const bool tap = taper && sb >= 2 && nframes >= 4 * sb;
  int f = 0;
  while (f < nframes) {
    const int left = nframes - f;
    const bool edge = tap && (f < 2 * small || left <= 2 * small);
    const int b = std::min(edge ? small : sb, left);
    // never let a full-size chunk eat into the tapered tail
    const int bb = (tap && !edge && left - b < 2 * small && left - b > 0) ? left - 2 * small : b;
    plan.push_back({f, std::max(1, bb)});
    f += plan.back().b;
  }
  return plan;
}

}  // namespace

// h_frame_len receives one length per PLANE (nframes entries for grey, 3 * nframes for RGB: the values the
// reference appends to `<out>_SIZE`, IO/IOManager.ipp:176-189).
int32_t pcs_encode_frames(pcs_ctx* c, const uint8_t* h_px, int32_t nframes, int32_t first, uint16_t* h_stream, size_t cap,
                          int32_t* h_len, size_t* total_words) {
  if (!c || !h_px || !h_stream || nframes < 1) return PCS_E_ARG;
  const size_t fpx = c->frame_px;
  const int nc = c->ncomp;
  int sb;
  const int nslots = slots_for(c, &sb);
  const std::vector<Chunk> plan = chunk_plan(nframes, sb);
  const int nchunks = (int)plan.size();
  SlotDrain drain{c, nslots};
  size_t written = 0;
  // second half of a chunk's work: its plane offsets are on the host -> copy its packed planes out
  auto finish = [&](int j) -> int {
    Slot s = make_slot(c, j % nslots, sb);
    const int f0 = plan[j].f0, b = plan[j].b;
    PCS_CUDA_TRY(cudaEventSynchronize(s.ev));
    const size_t words = (size_t)s.h_off[b * nc];
    if (written + words > cap) return fail(c, PCS_E_NOMEM, "output capacity too small");
    PCS_CUDA_TRY(cudaMemcpyAsync(h_stream + written, s.stream, words * 2, cudaMemcpyDefault, s.st));
    if (h_len) for (int f = 0; f < b * nc; ++f) h_len[(size_t)f0 * nc + f] = (int32_t)(s.h_off[f + 1] - s.h_off[f]);
    written += words;
    return PCS_OK;
  };
  // A chunk's packed frames can only be copied out once its length is on the host.  Finishing chunk i - lag right
  // after enqueueing chunk i keeps `lag` chunks of work queued on the GPU while the host waits, and leaves only
  // `lag` downloads for the end of the call.
  static const int lag_env = [] { const char* e = std::getenv("PCS_PIPE_LAG"); return e ? std::atoi(e) : 2; }();
  const int lag = std::max(1, std::min(lag_env, nslots));
  const int depth = kernel_depth(nslots, nchunks);
  int finished = 0;
  for (int i = 0; i < nchunks; ++i) {
    while (finished <= i - nslots) { int rc = finish(finished++); if (rc) return rc; }  // frees the slot (same stream: ordered)
    Slot s = make_slot(c, i % nslots, sb);
    const int f0 = plan[i].f0, b = plan[i].b;
    PCS_CUDA_TRY(cudaMemcpyAsync(s.px, h_px + (size_t)f0 * fpx, (size_t)b * fpx, cudaMemcpyDefault, s.st));
    if (i >= depth && depth < nslots) PCS_CUDA_TRY(cudaStreamWaitEvent(s.st, c->kev[(i - depth) % nslots], 0));
    int rc = run_dwt_forward(c, s.w, s.px, s.w.coeff, true, b, s.st);
    if (rc) return rc;
    rc = run_bpc_encode(c, s.w.coeff, false, s.w.cw, s.w.sizes, b, s.st);
    if (rc) return rc;
    rc = bsb_pack(s.w.cw, s.w.sizes, s.stream, s.d_off, s.w.status, bsb_args(c, b * nc, (first + f0) * nc), s.st);
    if (rc) return rc;
    PCS_CUDA_TRY(cudaEventRecord(c->kev[i % nslots], s.st));
    PCS_CUDA_TRY(cudaMemcpyAsync(s.h_off, s.d_off, (size_t)(b * nc + 1) * sizeof(long long), cudaMemcpyDeviceToHost, s.st));
    PCS_CUDA_TRY(cudaEventRecord(s.ev, s.st));
    while (finished <= i - lag) { int rc2 = finish(finished++); if (rc2) return rc2; }
  }
  while (finished < nchunks) { int rc = finish(finished++); if (rc) return rc; }
  for (int k = 0; k < nslots; ++k) PCS_CUDA_TRY(cudaStreamSynchronize(c->sstream[k]));
  drain.armed = false;
  if (total_words) *total_words = written;
  return PCS_OK;
}

int32_t pcs_decode_frames(pcs_ctx* c, const uint16_t* h_stream, const int32_t* h_len, int32_t nframes, uint8_t* h_px) {
  if (!c || !h_stream || !h_len || !h_px || nframes < 1) return PCS_E_ARG;
  const size_t fpx = c->frame_px;
  const int nc = c->ncomp;
  int sb;
  const int nslots = slots_for(c, &sb);
  const std::vector<Chunk> plan = chunk_plan(nframes, sb);
  const int nchunks = (int)plan.size();
  SlotDrain drain{c, nslots};
  size_t consumed = 0;
  const int depth = kernel_depth(nslots, nchunks);
  for (int i = 0; i < nchunks; ++i) {
    Slot s = make_slot(c, i % nslots, sb);
    const int f0 = plan[i].f0, b = plan[i].b;
    if (i >= nslots) PCS_CUDA_TRY(cudaEventSynchronize(s.ev));  // the slot's pinned offsets have been uploaded
    long long off = 0;
    for (int f = 0; f < b * nc; ++f) {
      const int len = h_len[(size_t)f0 * nc + f];
      if (len < 2 * c->ncb + 10 || (size_t)len > c->max_words) return fail(c, PCS_E_STREAM, "frame length out of range");
      s.h_off[f] = off; off += len;
    }
    s.h_off[b * nc] = off;
    PCS_CUDA_TRY(cudaMemcpyAsync(s.stream, h_stream + consumed, (size_t)off * 2, cudaMemcpyDefault, s.st));
    PCS_CUDA_TRY(cudaMemcpyAsync(s.d_off, s.h_off, (size_t)(b * nc + 1) * sizeof(long long), cudaMemcpyHostToDevice, s.st));
    PCS_CUDA_TRY(cudaEventRecord(s.ev, s.st));
    if (i >= depth && depth < nslots) PCS_CUDA_TRY(cudaStreamWaitEvent(s.st, c->kev[(i - depth) % nslots], 0));
    int rc = bsb_unpack(s.stream, s.d_off, s.w.cw, s.w.sizes, s.w.status, bsb_args(c, b * nc, 1), s.st);
    if (rc) return rc;
    rc = run_bpc_decode(c, s.w.cw, s.w.sizes, s.w.coeff, b, s.st);
    if (rc) return rc;
    rc = run_dwt_inverse(c, s.w, s.w.coeff, s.px, b, s.st);
    if (rc) return rc;
    PCS_CUDA_TRY(cudaEventRecord(c->kev[i % nslots], s.st));
    PCS_CUDA_TRY(cudaMemcpyAsync(h_px + (size_t)f0 * fpx, s.px, (size_t)b * fpx, cudaMemcpyDefault, s.st));
    consumed += (size_t)off;
  }
  for (int k = 0; k < nslots; ++k) PCS_CUDA_TRY(cudaStreamSynchronize(c->sstream[k]));
  drain.armed = false;
  return PCS_OK;
}

}  // extern "C"
