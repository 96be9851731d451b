// color.cu -- colour transforms of the RGB path for sm_100a.
//
// Replaces RGBTransformLossless / RGBTransformLossy of the encoder (Engines/CodingEngine.cu:357-405, matrix
// CodingEngine.cuh:25) and of the decoder (Engines/DecodingEngine.cu:599-650, matrix DecodingEngine.cuh:41),
// plus the host-side mirror padding of each component plane (IO/IOManager.ipp:72-112) and the host-side crop.
// Plane order after the forward transform: 0 = Y, 1 = B-G (Cb), 2 = R-G (Cr); component c of frame f is
// plane 3f+c everywhere downstream (it is coded as an independent "frame" with probability table c).
// The float operation shapes are the reference binary's (t = G*m1; t = fma(R, m0, t); t = fma(B, m2, t)), so
// lossy RGB codestreams match the reference byte for byte.
#include "pcs_common.cuh"

namespace pcs {

namespace {

__device__ __forceinline__ int floordiv4(int v) { return v >> 2; }  // arithmetic shift == floor division by 4

// u8 planar [frames][3][H][W] -> T planes [frames*3][Ha][Wa] (mirror padded)
template <bool LOSSY>
__global__ void __launch_bounds__(256) rgb_forward_kernel(const unsigned char* __restrict__ px, void* __restrict__ planes, int W, int H,
                                                         int Wa, int Ha) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
  if (x >= Wa) return;
  const size_t P = (size_t)W * H, A = (size_t)Wa * Ha;
  const size_t src = (size_t)pad_src(y, H) * W + pad_src(x, W);
  const unsigned char* p = px + (size_t)f * 3 * P + src;
  const int r = (int)__ldg(p) - 128, g = (int)__ldg(p + P) - 128, b = (int)__ldg(p + 2 * P) - 128;
  const size_t dst = (size_t)f * 3 * A + (size_t)y * Wa + x;
  if (LOSSY) {
    float* o = static_cast<float*>(planes) + dst;
    const float R = (float)r, G = (float)g, B = (float)b;
    o[0] = __fmaf_rn(B, 0.114f, __fmaf_rn(R, 0.299f, __fmul_rn(G, 0.587f)));
    o[A] = __fmaf_rn(B, 0.5f, __fmaf_rn(R, -0.168736f, __fmul_rn(G, -0.331264f)));
    o[2 * A] = __fmaf_rn(B, -0.081312f, __fmaf_rn(R, 0.5f, __fmul_rn(G, -0.418688f)));
  } else {
    int* o = static_cast<int*>(planes) + dst;
    o[0] = floordiv4(r + 2 * g + b);
    o[A] = b - g;
    o[2 * A] = r - g;
  }
}

// T planes [frames*3][Ha][Wa] -> u8 planar [frames][3][H][W]; flat_crop reproduces the reference's
// "first W*H elements of the W_a-stride plane" output for widths that are not multiples of 64.
template <bool LOSSY>
__global__ void __launch_bounds__(256) rgb_inverse_kernel(const void* __restrict__ planes, unsigned char* __restrict__ px, int W, int H,
                                                         int Wa, int Ha, int flat_crop) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
  if (x >= Wa) return;
  const size_t P = (size_t)W * H, A = (size_t)Wa * Ha;
  const size_t src = (size_t)f * 3 * A + (size_t)y * Wa + x;
  int R, G, B;
  if (LOSSY) {
    const float* i = static_cast<const float*>(planes) + src;
    const float Y = __ldg(i), Cb = __ldg(i + A), Cr = __ldg(i + 2 * A);
    R = __float2int_rn(__fadd_rn(__fmaf_rn(Cr, 1.402f, Y), 0.01f));
    G = __float2int_rn(__fadd_rn(__fmaf_rn(Cr, -0.714136f, __fmaf_rn(Cb, -0.344136f, Y)), 0.01f));
    B = __float2int_rn(__fadd_rn(__fmaf_rn(Cb, 1.772f, Y), 0.01f));
  } else {
    const int* i = static_cast<const int*>(planes) + src;
    const int Y = __ldg(i), Cb = __ldg(i + A), Cr = __ldg(i + 2 * A);
    G = Y - floordiv4(Cb + Cr); R = Cr + G; B = Cb + G;
  }
  size_t dst;
  if (flat_crop) {
    dst = (size_t)y * Wa + x;
    if (dst >= P) return;
  } else {
    if (x >= W || y >= H) return;
    dst = (size_t)y * W + x;
  }
  unsigned char* o = px + (size_t)f * 3 * P + dst;
  o[0] = (unsigned char)min(max(R + 128, 0), 255);
  o[P] = (unsigned char)min(max(G + 128, 0), 255);
  o[2 * P] = (unsigned char)min(max(B + 128, 0), 255);
}

}  // namespace

int rgb_forward(const uint8_t* d_px, void* d_planes, int W, int H, int Wa, int Ha, bool lossy, int frames, cudaStream_t st) {
  dim3 grid((Wa + 255) / 256, Ha, frames);
  if (lossy) rgb_forward_kernel<true><<<grid, 256, 0, st>>>(d_px, d_planes, W, H, Wa, Ha);
  else rgb_forward_kernel<false><<<grid, 256, 0, st>>>(d_px, d_planes, W, H, Wa, Ha);
  ++g_launches;
  return (int)cudaGetLastError();
}

int rgb_inverse(const void* d_planes, uint8_t* d_px, int W, int H, int Wa, int Ha, bool lossy, bool flat_crop, int frames, cudaStream_t st) {
  dim3 grid((Wa + 255) / 256, Ha, frames);
  if (lossy) rgb_inverse_kernel<true><<<grid, 256, 0, st>>>(d_planes, d_px, W, H, Wa, Ha, flat_crop ? 1 : 0);
  else rgb_inverse_kernel<false><<<grid, 256, 0, st>>>(d_planes, d_px, W, H, Wa, Ha, flat_crop ? 1 : 0);
  ++g_launches;
  return (int)cudaGetLastError();
}

}  // namespace pcs
