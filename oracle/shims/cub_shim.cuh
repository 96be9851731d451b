/* Build shim for oracle/_ref (test infrastructure, not product code).
 *
 * The reference's only third-party arithmetic is
 *   cub::DeviceScan::InclusiveSum(temp, temp_bytes, d_in, d_out, n[, stream])
 * (BitStreamBuilder/BitStreamBuilder.cu:297,312), taken from a vendored CUB 1.x
 * snapshot that does not compile under CUDA 12.9.  An inclusive int32 prefix sum
 * has exactly one correct answer, so this shim supplies that entry point with a
 * plain single-block kernel.  It is force-included (-include) with -DCUB_DATA so
 * the reference's own `#ifndef CUB_DATA / #include "../cub/cub.cuh"` is skipped.
 * The reference passes a deliberately small temp buffer; this shim needs none.
 */
#pragma once
#include <cuda_runtime.h>

namespace cub {
namespace shim_detail {
template <typename T>
__global__ void inclusive_sum_1block(const T* in, T* out, int n) {
  __shared__ T warp_tot[32];
  __shared__ T carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + tid;
    T v = (i < n) ? in[i] : T(0);
    for (int d = 1; d < 32; d <<= 1) {
      T o = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += o;
    }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    if (wid == 0) {
      T w = (lane < (blockDim.x >> 5)) ? warp_tot[lane] : T(0);
      for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += o;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    T add = carry_s + (wid ? warp_tot[wid - 1] : T(0));
    if (i < n) out[i] = v + add;
    __syncthreads();
    if (tid == 0) carry_s += warp_tot[(blockDim.x >> 5) - 1];
    __syncthreads();
  }
}
}  // namespace shim_detail

struct DeviceScan {
  template <typename SizeT, typename T>
  static cudaError_t InclusiveSum(void* /*d_temp*/, SizeT& /*temp_bytes*/, T* d_in, T* d_out,
                                  int n, cudaStream_t stream = 0) {
    if (n <= 0) return cudaSuccess;
    shim_detail::inclusive_sum_1block<T><<<1, 1024, 0, stream>>>(d_in, d_out, n);
    return cudaGetLastError();
  }
};
}  // namespace cub
