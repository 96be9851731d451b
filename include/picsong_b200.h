/*
 * picsong_b200.h -- C ABI of the B200-native PICSONG hot path (libpicsong_b200.so).
 *
 * The reference (13Karl/CUDA-Image-and-Video-codec) has no FFI layer; its boundary is the set of
 * host-blocking C++ stage calls that Engines/CodingEngine.cu and Engines/DecodingEngine.cu make on
 * device buffers.  Each entry point below names the reference interface it replaces
 * (file:line relative to /root/reference/CUDA_ImCod/).  Conventions:
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = default)
 *   - every function returns 0 on success, a negative PCS_E_* code, or a positive cudaError_t
 *   - device-pointer entry points enqueue on `stream` and do NOT synchronise the host
 *   - no allocation after pcs_create(); there is no CPU fallback: without a CUDA device
 *     pcs_create() fails with the CUDA error
 *   - batched layouts: `batch` frames are processed per call (video frames / row bands); the
 *     reference processes one frame per call
 */
#ifndef PICSONG_B200_H
#define PICSONG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCS_CB 64          /* BPC codeblock is 64x64 (BPC/BPCEngine.cuh:29-36) */
#define PCS_CB_WORDS 4096  /* uint16 words reserved per codeblock: [0]=MSB, [1+j]=codeword j */

enum {
  PCS_OK = 0,
  PCS_E_ARG = -1,        /* invalid parameter ("Incorrect parameters", Launcher.cu:132-136) */
  PCS_E_LUT = -2,        /* LUT folder unreadable / malformed */
  PCS_E_LUT_RANGE = -3,  /* LUT shape or probability outside what the coder supports */
  PCS_E_UNSUPPORTED = -4,/* cp==3, signed / 16-bit input: out of scope */
  PCS_E_NOMEM = -5,
  PCS_E_STREAM = -6,     /* malformed codestream */
  PCS_E_BATCH = -7       /* batch larger than pcs_params.max_batch */
};

/* pcs_params.flags */
#define PCS_F_REF_FLAT_CROP 1u /* decode output: reproduce the reference's "first W*H elements of the
                                  W_a-stride plane" quirk (IO/IOManager.ipp:222-225,332-335) instead of
                                  a proper W x H crop.  Identical whenever W is a multiple of 64. */

#define PCS_F_STILLS 2u        /* the context codes stills although `frames` (a header-only value) is > 0:
                                  RGB component i is coded with iter = i (Engines/CodingEngine.cu:620,700), so only
                                  plane 0 carries the header; without the flag frames > 0 means video, where the three
                                  planes of frame 0 all do (CodingEngine.cu:790,902). */

#define PCS_F_CB_LENGTH_32 4u  /* 64 x 32 codeblocks: the reference rebuilt with -DCBLOCK_LENGTH=32 (BPC/BPCEngine.cuh:33-36;
                                  BASELINE configs[2]'s sweep point).  The frame is padded to a multiple of 32 rows, a block
                                  owns 2048 words.  k = 0 only.  (The reference build keeps its 4094 / 4095 codeword clamps,
                                  so it is only defined while no block needs more than 2046 codewords; this library clamps
                                  at 2046 / 2047 and falls back to the raw dump at 2048 like the 64-row coder does at 4096.) */

typedef struct pcs_ctx pcs_ctx;

/* Mirrors the Engine setters (Engines/Engine.cuh:28-44) + Image fields (Image/Image.hpp:52-62). */
typedef struct pcs_params {
  int32_t width, height;      /* original frame size (-xSize/-ySize) */
  int32_t wavelet_levels;     /* setWaveletLevels, 1..10 (header can carry <= 7) */
  int32_t lossy;              /* setWType: 0 = reversible 5/3, 1 = irreversible 9/7 */
  float quant_size;           /* setQSizes (-qs), (0,1] */
  int32_t coding_passes;      /* setCodingPasses: 2 (3 is out of scope) */
  float k_factor;             /* setKFactor (-k): 0, or > 0 = complexity-scalable mode (needs the pcs_lut_load_k tables) */
  int32_t cb_width, cb_height;/* setCBWidth/setCBHeight: header-only values (64, 18..20) */
  int32_t frames;             /* setNumberOfFrames: value written to the header (0 for stills) */
  int32_t bit_depth, bits_per_sample, endianess, is_signed, components, is_rgb;
  int32_t max_batch;          /* frames processed per call (workspace is sized for this) */
  uint32_t flags;
  int32_t lut_header[8];      /* header.txt values (Engines/Engine.cu:190-210) */
  const int32_t* lut;         /* host table as produced by pcs_lut_load (Engine.cu:124-139) */
  int32_t lut_len;
} pcs_params;

/* ---- host utilities (no GPU needed) -------------------------------------------------------- */

/* SupportFunctions/AuxiliarFunctions.cpp:22-26 fixImageProportions */
void pcs_adapted_dims(int32_t width, int32_t height, int32_t* adapted_w, int32_t* adapted_h);
/* ... with the codeblock length (64 or 32) the height is padded to (see PCS_F_CB_LENGTH_32) */
void pcs_adapted_dims_cb(int32_t width, int32_t height, int32_t cb_rows, int32_t* adapted_w, int32_t* adapted_h);

/* Engine::initLUT grey/k==0/cp==2 branch (Engines/Engine.cu:101-141) + IOManager::loadLUTHeaders
 * (IO/IOManager.ipp:363-386) + loadLUTUpgraded (:404-612).  `folder` is used as the reference
 * uses -LUTFolder: "<folder>/header.txt" and "<folder>ref[R].txt_0" etc.  Returns the number of
 * int32 entries written (>0) or PCS_E_LUT.  Entries the files do not define are 0. */
int32_t pcs_lut_load(const char* folder, int32_t wavelet_levels, int32_t header[8], int32_t* table,
                     int32_t capacity);
int32_t pcs_lut_size(const int32_t header[8], int32_t wavelet_levels);
/* RGB path: the table set of component 0/1/2 = files "R"/"G"/"B" (Engines/Engine.cu:124-139).  For an RGB
 * context pcs_params.lut is the three component tables back to back (lut_len = 3 * pcs_lut_size). */
int32_t pcs_lut_load_component(const char* folder, int32_t wavelet_levels, int32_t component, int32_t header[8],
                               int32_t* table, int32_t capacity);
/* Complexity-scalable mode (-k > 0, Engines/Engine.cu:12-100): the tables of ALL bit-plane files `*_0 .. *_(n-1)`
 * (n = header[7], AMOUNT_OF_BITPLANE_FILES) of one component back to back, pcs_lut_size() ints each.  A context
 * created with k_factor > 0 expects these tables (lut_len = components * n * pcs_lut_size()) instead of the k = 0
 * ones.  Returns the number of ints written or a negative PCS_E_* code. */
int32_t pcs_lut_load_k(const char* folder, int32_t wavelet_levels, int32_t component, int32_t header[8],
                       int32_t* table, int32_t capacity);

/* BitStreamBuilder::setExtraInformation (BitStreamBuilder/BitStreamBuilder.cpp:35-94) */
void pcs_header_pack(const pcs_params* p, uint16_t out[9]);
/* DecodingEngine::getExtraInformation (Engines/DecodingEngine.cu:567-585); fills the fields a
 * decoder needs (width, height, wavelet_levels, lossy, quant_size = int(qs*10000)/10000, ...) */
int32_t pcs_header_unpack(const uint16_t in[9], pcs_params* p);

/* Sizes a caller needs to allocate buffers (all per frame). */
size_t pcs_frame_pixels(const pcs_ctx*);        /* W_a*H_a = coefficient count A */
int32_t pcs_frame_codeblocks(const pcs_ctx*);   /* nCB = A/4096 */
size_t pcs_frame_max_stream_words(const pcs_ctx*); /* planes * (A + nCB + 10) uint16: upper bound of one frame */
/* RGB (pcs_params.is_rgb = 1; Engines/CodingEngine.cu:598-630,680-710,764-810, Engines/DecodingEngine.cu:738-770):
 * a frame is u8 planar [3][height][width]; after the colour transform (RCT for 5/3, ICT for 9/7) its three
 * component planes are coded as three independent "frames" with the R/G/B probability tables.  Every batched
 * layout below then holds batch * 3 PLANES where it says [batch] (plane 3f+c = component c of frame f), frame
 * offset / length arrays have one entry per plane, and only plane 0 of a still (planes 0..2 of video frame 0)
 * carries the 9-word header.  1 for greyscale. */
int32_t pcs_frame_planes(const pcs_ctx*);

/* ---- context ------------------------------------------------------------------------------- */

/* Replaces Engine::initLUT + CodingEngine/DecodingEngine::initMemory (Engines/Engine.cu:8-185,
 * Engines/CodingEngine.cu:119-199, Engines/DecodingEngine.cu:140-235): uploads the probability
 * tables and allocates every device buffer once. */
int32_t pcs_create(const pcs_params* params, pcs_ctx** out);
void pcs_destroy(pcs_ctx* ctx);
const char* pcs_last_error(const pcs_ctx* ctx);

/* ---- stage calls on device buffers --------------------------------------------------------- */

/* offsetImage (Engines/CodingEngine.cu:581-588) + mirror padding (IO/IOManager.ipp:72-112, done on
 * the host in the reference) + DWT<T,Y>::DWTEncode (DWT/DWTGenerator.ipp:43-55 ->
 * DWTEngine::DWTForward DWT/DWTGenerator.cu:1268-1342).
 * d_pixels: uint8 [batch][height][width] (unpadded).  d_coeff: [batch][A] Mallat plane, int32 for
 * 5/3; for 9/7 float (the reference's layout) or, if trunc_to_int != 0, int32 = (int)c, which is
 * what the coder consumes (BPC/BPCEngine.cu:49,55). */
int32_t pcs_dwt_forward(pcs_ctx* ctx, const uint8_t* d_pixels, void* d_coeff, int32_t trunc_to_int,
                        int32_t batch, void* stream);

/* DWT<T,Y>::DWTDecode (DWT/DWTGenerator.ipp:60-72 -> DWTEngine::DWTReverse DWTGenerator.cu:1349-1424)
 * + removeOffsetAndApplyMaxMin[Lossy] (Engines/DecodingEngine.cu:706-729) + the u8 narrowing and
 * crop the reference does on the host (IO/IOManager.ipp:214-231,332-335).
 * d_q: int32 [batch][A]; d_pixels: uint8 [batch][height][width]. */
int32_t pcs_dwt_inverse(pcs_ctx* ctx, const int32_t* d_q, uint8_t* d_pixels, int32_t batch, void* stream);

/* BPCEngine<T>::kernelLauncher(CODE) + deviceMemoryAllocator (BPC/BPCEngine.cu:2307-2441), the
 * first half of BPCCuda<T>::Code (BPC/BPCEngine.ipp:24-37).
 * d_coeff: int32 or float [batch][A] (coeff_is_float selects); d_codewords: uint16
 * [batch][nCB][4096] (word 0 = MSB, word 1+j = codeword j; only the first d_sizes[cb] words of a
 * block are defined); d_sizes: int32 [batch][nCB] = 1 + codewords (4096 = expansion fallback).
 * Domain: |coefficient| < 2^23 (what 8..16-bit samples through the DWT can produce; Annex B.2 of SURVEY.md): the decoder
 * treats MSB values 23..31 as "cannot occur" and writes zeros for such a block. */
int32_t pcs_bpc_encode(pcs_ctx* ctx, const void* d_coeff, int32_t coeff_is_float, uint16_t* d_codewords,
                       int32_t* d_sizes, int32_t batch, void* stream);

/* BPCEngine<int>::kernelLauncher(DECODE), second half of BPCCuda<T>::Decode (BPCEngine.ipp:42-58). */
int32_t pcs_bpc_decode(pcs_ctx* ctx, const uint16_t* d_codewords, const int32_t* d_sizes, int32_t* d_q,
                       int32_t batch, void* stream);

/* BitStreamBuilder::createBitStream (BitStreamBuilder/BitStreamBuilder.cpp:100-114): inclusive scan
 * of sizes (cub::DeviceScan, BitStreamBuilder.cu:290-323), search index (:198-265) and compaction
 * (:106-137), as ONE single-pass decoupled-look-back kernel.  Frames are written back to back:
 * frame f starts at d_out + d_frame_offset[f] (uint16 units) and is
 * d_frame_offset[f+1] - d_frame_offset[f] long (= sum(sizes) + nCB + 10).  The frame whose global index (first_frame_index + f) is 0 carries the
 * 9-word header, the others 0xFFFF x 9 (BitStreamBuilder.cu:276-278).
 * d_frame_offset: int64 [batch+1] (written by the kernel). */
int32_t pcs_bsb_pack(pcs_ctx* ctx, const uint16_t* d_codewords, const int32_t* d_sizes,
                     int32_t first_frame_index, uint16_t* d_out, int64_t* d_frame_offset,
                     int32_t batch, void* stream);

/* BitStreamBuilder::createCodeStream (BitStreamBuilder.cpp:119-153, BitStreamBuilder.cu:142-171):
 * per-block sizes are read from the stream on the device (the reference parses them on the host). */
int32_t pcs_bsb_unpack(pcs_ctx* ctx, const uint16_t* d_in, const int64_t* d_frame_offset,
                       uint16_t* d_codewords, int32_t* d_sizes, int32_t batch, void* stream);

/* Debug/parity exports: the reference's exact intermediate layout (int32 sparse codestream
 * `DCodeStreamValues`, -1 filled; BPC/BPCEngine.cu:2431-2440, :1996-2013). */
int32_t pcs_export_ref_codestream(pcs_ctx* ctx, const uint16_t* d_codewords, const int32_t* d_sizes,
                                  int32_t* d_ref_cs, int32_t batch, void* stream);

/* ---- fused frame calls with HOST buffers (what CodingEngine::runImage/runVideo and
 * DecodingEngine::runImage/runVideo do per frame: Engines/CodingEngine.cu:593-983,
 * Engines/DecodingEngine.cu:765-1043).  Pixels may be pageable or pinned.  Blocking. ----------- */

/* h_pixels: uint8 [nframes][height][width].  h_stream receives the frames back to back;
 * h_frame_len[f] = length of frame f in uint16 units (the `_SIZE` sidecar values).  Returns the
 * total number of uint16 written through *total_words.  nframes is not limited by max_batch (the
 * call pipelines chunks of frames through the workspace).  The pixel and stream pointers of both
 * calls may also be DEVICE pointers (copies are issued with cudaMemcpyDefault): the multi-rank
 * path keeps a rank's packed stream in HBM until the all-gather of the lengths has told it where
 * in the concatenated host stream its bytes go (picsong_b200/sharding.py). */
int32_t pcs_encode_frames(pcs_ctx* ctx, const uint8_t* h_pixels, int32_t nframes, int32_t first_frame_index,
                          uint16_t* h_stream, size_t capacity_words, int32_t* h_frame_len,
                          size_t* total_words);
int32_t pcs_decode_frames(pcs_ctx* ctx, const uint16_t* h_stream, const int32_t* h_frame_len, int32_t nframes,
                          uint8_t* h_pixels);

/* Pinned host buffers for the two calls above (the reference's frame buffers are cudaHostAlloc'd too,
 * Engines/CodingEngine.cu:129-131,289).  NULL on failure. */
void* pcs_host_alloc(size_t bytes);
void pcs_host_free(void* p);

/* Same, device-resident in and out (no PCIe), for throughput measurement and for callers that
 * keep frames on the GPU.  d_stream needs max_batch * pcs_frame_max_stream_words() capacity. */
int32_t pcs_encode_device(pcs_ctx* ctx, const uint8_t* d_pixels, int32_t batch, int32_t first_frame_index,
                          uint16_t* d_stream, int64_t* d_frame_offset, void* stream);
int32_t pcs_decode_device(pcs_ctx* ctx, const uint16_t* d_stream, const int64_t* d_frame_offset, int32_t batch,
                          uint8_t* d_pixels, void* stream);

/* Per-stage device time (ms) of the last pcs_encode_device/pcs_decode_device call when
 * PCS_STAGE_TIMING=1 was set at pcs_create: [0] dwt [1] bpc [2] bsb.  Synchronises. */
int32_t pcs_last_stage_ms(pcs_ctx* ctx, float ms[3]);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t pcs_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PICSONG_B200_H */
