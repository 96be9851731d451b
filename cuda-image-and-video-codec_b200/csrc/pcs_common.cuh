// pcs_common.cuh -- shared device/host helpers of libpicsong_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/picsong_b200.h"

namespace pcs {

constexpr int kCB = 64;          // codeblock edge (BPC/BPCEngine.cuh:29-36 of the reference)
constexpr int kCBWords = 4096;   // words reserved per codeblock
constexpr int kMaxLevels = 10;   // Launcher.cu:132 accepts wl in [1,10]
// Probability rows the k = 0 decoder stages in shared memory: (nSub * wl + 1) * nBp + 32 for the shipped table shape
// (nSub = 3, nBp = 15) at the deepest wl.  pcs_create refuses tables with more rows (PCS_E_LUT_RANGE) so that the
// encoder (global reads) and the decoder (shared copy) can never disagree on a row.
constexpr int kMaxProbRows = (3 * kMaxLevels + 1) * 15 + 32;

// Whole-sample symmetric extension: fold an index into [0, n), period 2(n-1).
// (x[-i] = x[i], x[n-1+i] = x[n-1-i]; DWT/DWTGenerator.cu:148-150,285,290 mirror rule.)
__host__ __device__ __forceinline__ int fold_ws(int i, int n) {
  if (n == 1) return 0;
  const int m = 2 * (n - 1);
  i %= m;
  if (i < 0) i += m;
  return i < n ? i : m - i;
}

// Edge-inclusive mirror padding of the ORIGINAL frame to the adapted (x64) size
// (IO/IOManager.ipp:99-110: column W+j <- W-1-j, row H+r <- H-1-r).
__host__ __device__ __forceinline__ int pad_src(int i, int n) { return i < n ? i : 2 * n - 1 - i; }

extern unsigned long long g_launches;  // host-side counter of kernel launches (pcs_kernel_launch_count)

// Arguments of the bit-plane coder kernels (bpc.cu)
struct BpcArgs {
  int Wa, Ha, nbx, ncb, wl, nBp, nSub, nrows, prec, total_cb;
  int pstep, pfirst;               // plane of the launch's frame f = f * pstep + pfirst (RGB: one launch per component, pstep = 3)
  int cb_rows;                     // codeblock length: 64, or 32 (PCS_F_CB_LENGTH_32)
  long long A;
  const uint4* tab4;               // [nrows] probability rows of this launch's component, as bytes: x = sig ctx 0..3, y = sig ctx 4..7, z = sig ctx 8 | ref << 8, w = sign ctx 0..3
};
// Arguments of the bit-stream builder kernels (bsb.cu)
// complexity-scalable mode (k > 0): flat int32 tables of all bit-plane files, reference indexing
struct BpcKArgs {
  int Wa, Ha, nbx, ncb, wl, nBp, nSub, cR, cS, cG, prec, total_cb;
  int pstep, pfirst;
  int per, lutN;                   // ints per LUT set (file), ints in `lut`
  long long A;
  float k;
  const int* lut;                  // this launch's component: [files][per]
};

struct BsbArgs {
  int ncb;                // codeblocks per frame
  int batch;              // frames
  int first_frame_index;  // global index of frame 0 of this batch
  int hdr_count;          // global frames 0 .. hdr_count-1 carry the header (1; RGB video: 3, BitStreamBuilder.cpp:104 with iter = frame)
  int cbw;                // words a codeblock owns in the sparse layout: 64 x codeblock length
  unsigned short hdr[9];
};

}  // namespace pcs

#define PCS_CUDA_TRY(expr)                     \
  do {                                         \
    cudaError_t e__ = (expr);                  \
    if (e__ != cudaSuccess) return (int)e__;   \
  } while (0)
