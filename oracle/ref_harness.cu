/*
 * ref_harness.cu -- TEST / BASELINE INFRASTRUCTURE, not product code.
 *
 * A thin driver around the UNMODIFIED reference stage wrappers.  It is compiled by
 * oracle/build_ref.sh against the reference's own headers where they lie under /root/reference and
 * linked with the reference's own objects (everything except Launcher.cu's main), and it calls
 *   DWT<T,Y>::DWTEncode / ::DWTDecode      (DWT/DWTGenerator.ipp:43-72)
 *   BPCCuda<T>::Code / ::Decode            (BPC/BPCEngine.ipp:24-58)
 * on device buffers allocated the way CodingEngine::initMemory / DecodingEngine::initMemory do
 * (Engines/CodingEngine.cu:119-199, Engines/DecodingEngine.cu:140-235), in the order
 * CodingEngine::runImage / DecodingEngine::runImage use (Engines/CodingEngine.cu:634-674,
 * Engines/DecodingEngine.cu:770-859).  Two uses:
 *   1. stage dumps (wavelet coefficients, sparse codestream, packed stream) that pin the oracle;
 *   2. the reference arm of bench.py: stage-level device time of the reference's own CUDA path on
 *      the same B200, with pixels already on the device (T_dev) and from host buffers (T_e2e).
 * The level-shift / clamp kernels here restate offsetImage (CodingEngine.cu:581-588) and
 * removeOffsetAndApplyMaxMin[Lossy] (DecodingEngine.cu:706-729), which the reference defines as
 * file-local kernels; they are elementwise and not part of what is being measured against.
 *
 * usage: ref_harness W H wl type qs LUTFolder/ input.raw frames iters dump_prefix|- [json_out] [coef=plane.i32]
 *   coef=<file>: coder-only mode (type 0).  The int32 plane in <file> (W_a*H_a values, Mallat layout) is fed to
 *   BPCCuda<int>::Code in place of the DWT output and the result decoded with ::Decode; dumps <prefix>.stream and
 *   <prefix>.q (the decoder's int32 plane).  Pins the rare coder branches (expansionFix, planes >= 16).
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "Engines/Engine.cuh"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

template <class T>
__global__ void h_offset(const unsigned char* in, T* out, int n) {
  int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (i < n) out[i] = (T)in[i] - (T)128;
}
__global__ void h_clamp_i(int* d, int n) {
  int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (i < n) d[i] = max(min(d[i] + 128, 255), 0);
}
__global__ void h_clamp_f(float* d, int n) {
  int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (i < n) d[i] = fmaxf(fminf(__float2int_rn(d[i] + 128 + 0.01f), 255.0f), 0.0f);
}

static void dump(const std::string& path, const void* p, size_t bytes) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); exit(3); }
  fwrite(p, 1, bytes, f);
  fclose(f);
}

struct Stat { double sum = 0; double best = 1e30; int n = 0; void add(double v) { sum += v; if (v < best) best = v; n++; } };

template <class T, class Y>
static int run(int W, int H, int wl, bool lossy, float qs, const std::string& lutFolder, const std::string& input,
               int frames, int iters, const std::string& prefix, const std::string& jsonOut, const std::string& coefFile = "") {
  Image img;
  img.setBitDepth(8); img.setBitsPerSample(8); img.setComponents(1); img.setEndianess(0);
  img.setHeight(H); img.setWidth(W); img.setIsRGB(false); img.setName(input); img.setSignedOrUnsigned(0);
  SupportFunctions::fixImageProportions(&img, CBLOCK_LENGTH, CBLOCK_WIDTH);
  const int Wa = img.getAdaptedWidth(), Ha = img.getAdaptedHeight();
  const int A = Wa * Ha;
  const int nCB = (int)ceil(Wa / (float)CBLOCK_WIDTH) * (int)ceil(Ha / (float)CBLOCK_LENGTH);

  /* LUT: Engine::initLUT k==0/cp==2/3-file branch, component "R" (Engine.cu:124-139) */
  IOManager<int, int2> iom;
  iom.setInputFile(lutFolder + "/header.txt");
  int* hv = iom.loadLUTHeaders();
  const int nBp = hv[0], nSub = hv[1], cR = hv[2], cG = hv[3], cS = hv[4], prec = hv[5], nFiles = hv[6];
  iom.setInputFile("");
  iom.setInputFolder(lutFolder);
  size_t lutInts = (size_t)(nSub * nBp * cR * wl + nBp * cR) + (nSub * nBp * cS * wl + nBp * cS) + (nSub * nBp * cG * wl + nBp * cG);
  int* hLUT = (int*)calloc(lutInts * 16, sizeof(int));
  iom.loadLUTUpgraded(nSub, cR, cS, cG, 2, wl, nBp, hLUT, nFiles == 1 ? 0 : 1, qs, 0);
  int* dLUT;
  CK(cudaMalloc(&dLUT, lutInts * 16 * sizeof(int)));
  CK(cudaMemcpy(dLUT, hLUT, lutInts * sizeof(int), cudaMemcpyHostToDevice));

  int extra = 0;
  for (int i = 1; i < wl; ++i) extra += (Wa / (2 << (i - 1))) * (Ha / (2 << (i - 1)));

  unsigned char* dPx; T* dIn; T* dCoef; int* dCS; int* dSize; int* dPref; int* dTemp; int* dLUTBS; unsigned short* dBS;
  CK(cudaMalloc(&dPx, A));
  CK(cudaMalloc(&dIn, (size_t)(A + extra) * sizeof(T)));
  CK(cudaMalloc(&dCoef, (size_t)(A + extra) * sizeof(T)));
  CK(cudaMalloc(&dCS, (size_t)A * sizeof(int)));
  CK(cudaMalloc(&dSize, (size_t)nCB * sizeof(int)));
  CK(cudaMalloc(&dPref, (size_t)(nCB + 1) * sizeof(int)));
  int storage = A / (CBLOCK_WIDTH * CBLOCK_LENGTH * 2);
  if (storage < 1000) storage = 1000;
  CK(cudaMalloc(&dTemp, storage));
  CK(cudaMalloc(&dLUTBS, 257 * sizeof(int)));
  CK(cudaMalloc(&dBS, (size_t)A * sizeof(unsigned short)));
  int* dQ; CK(cudaMalloc(&dQ, (size_t)A * sizeof(int)));
  unsigned char* hPxPin; CK(cudaHostAlloc(&hPxPin, A, 0));
  unsigned short* hBS; CK(cudaHostAlloc(&hBS, (size_t)A * sizeof(unsigned short), 0));
  T* hOut; CK(cudaHostAlloc(&hOut, (size_t)A * sizeof(T), 0));
  unsigned short hExtra[9];
  int* hTotal = (int*)malloc(sizeof(int));
  int* hSizeArr = (int*)malloc((size_t)nCB * sizeof(int));
  int hBasic[16] = {0};

  IOManager<unsigned char, uchar2> iomc;
  Stat sDwtE, sBpcE, sEnc, sEncE2E, sBpcD, sDwtD, sDec, sDecE2E;
  cudaEvent_t e0, e1, e2, e3;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
  std::vector<int> frameLens;
  const int nf = frames > 0 ? frames : 1;
  int bad = 0;
  for (int it = 0; it < iters; ++it) {
    for (int f = 0; f < nf; ++f) {
      if (coefFile.empty()) iomc.loadFrameCAdaptedSizes(&img, hPxPin, f); /* host mirror padding, IOManager.ipp:72-112 */
      auto w0 = std::chrono::steady_clock::now();
      CK(cudaMemcpy(dPx, hPxPin, A, cudaMemcpyHostToDevice));
      CK(cudaEventRecord(e0));
      if (!coefFile.empty()) { /* coder-only mode: the plane comes from the file */
        std::vector<int> plane(A);
        FILE* cf = fopen(coefFile.c_str(), "rb");
        if (!cf || fread(plane.data(), sizeof(int), A, cf) != (size_t)A) { fprintf(stderr, "cannot read %s\n", coefFile.c_str()); exit(4); }
        fclose(cf);
        CK(cudaMemcpy(dCoef, plane.data(), (size_t)A * sizeof(int), cudaMemcpyHostToDevice));
      } else {
        h_offset<T><<<(A + 255) / 256, 256>>>(dPx, dIn, A);
        DWT<T, Y> dwt(&img, lossy, wl, 64, 18, qs);
        dwt.DWTEncode(dIn, dCoef, cudaStreamDefault);
      }
      CK(cudaEventRecord(e1));
      double tb = 0;
      {
        BPCCuda<T> bpc(&img, dCoef, wl, 64, 18, 2, lossy, qs, 0.0f, 15);
        bpc.Code(nBp, nSub, cR, cG, cS, prec, dLUT, dCS, dPref, dTemp, dSize, hExtra, dBS, hTotal, dLUTBS, 256,
                 frames > 0 ? f : 0, cudaStreamDefault, frames, &tb);
      }
      CK(cudaEventRecord(e2));
      CK(cudaEventSynchronize(e2));
      CK(cudaMemcpy(hBS, dBS, (size_t)hTotal[0] * sizeof(unsigned short), cudaMemcpyDeviceToHost));
      auto w1 = std::chrono::steady_clock::now();
      float a, b, c;
      CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2)); CK(cudaEventElapsedTime(&c, e0, e2));
      if (it > 0 || iters == 1) {
        sDwtE.add(a); sBpcE.add(b); sEnc.add(c);
        sEncE2E.add(std::chrono::duration<double, std::milli>(w1 - w0).count());
      }
      if (it == 0) {
        frameLens.push_back(hTotal[0]);
        if (prefix != "-") {
          std::string p = prefix + (frames > 0 ? ".f" + std::to_string(f) : "");
          CK(cudaMemcpy(hOut, dCoef, (size_t)A * sizeof(T), cudaMemcpyDeviceToHost));
          dump(p + ".coef", hOut, (size_t)A * sizeof(T));
          dump(p + ".stream", hBS, (size_t)hTotal[0] * sizeof(unsigned short));
        }
      }
      /* ---- decode the frame just produced (DecodingEngine::runImage order) ---- */
      int total = hTotal[0];
      auto w2 = std::chrono::steady_clock::now();
      CK(cudaMemcpy(dBS, hBS, (size_t)total * sizeof(unsigned short), cudaMemcpyHostToDevice));
      CK(cudaEventRecord(e0));
      hBasic[1] = 2; hBasic[2] = 18; hBasic[3] = 64; hBasic[4] = wl;
      double td = 0;
      {
        BPCCuda<unsigned short> bpcd(&img, hBS, wl, 64, 18, 2, lossy, qs, 0.0f, 15);
        bpcd.Decode(A, nBp, nSub, cR, cG, cS, prec, dLUT, dPref, dSize, hBasic, dTemp, dBS, dCS, hSizeArr, hTotal, dQ,
                    cudaStreamDefault, 256, dLUTBS, &td);
      }
      CK(cudaEventRecord(e1));
      if (!coefFile.empty()) {
        CK(cudaDeviceSynchronize());
        std::vector<int> qh(A);
        CK(cudaMemcpy(qh.data(), dQ, (size_t)A * sizeof(int), cudaMemcpyDeviceToHost));
        if (prefix != "-") dump(prefix + ".q", qh.data(), (size_t)A * sizeof(int));
        printf("{\"coef_mode\":1,\"stream_words\":%d}\n", total);
        return 0;
      }
      {
        DWT<T, Y> dwtd(&img, lossy, wl, 64, 18, qs);
        dwtd.DWTDecode(dQ, dIn, cudaStreamDefault);
      }
      if (lossy) h_clamp_f<<<(A + 255) / 256, 256>>>((float*)(dIn + extra), A);
      else h_clamp_i<<<(A + 255) / 256, 256>>>((int*)(dIn + extra), A);
      CK(cudaEventRecord(e2));
      CK(cudaEventSynchronize(e2));
      CK(cudaMemcpy(hOut, dIn + extra, (size_t)A * sizeof(T), cudaMemcpyDeviceToHost));
      auto w3 = std::chrono::steady_clock::now();
      CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2)); CK(cudaEventElapsedTime(&c, e0, e2));
      if (it > 0 || iters == 1) {
        sBpcD.add(a); sDwtD.add(b); sDec.add(c);
        sDecE2E.add(std::chrono::duration<double, std::milli>(w3 - w2).count());
      }
      if (it == 0) {
        /* round-trip check against the padded input */
        for (int i = 0; i < A; ++i) {
          int v = (int)hOut[i];
          if (!lossy && v != hPxPin[i]) bad++;
          if (lossy && abs(v - (int)hPxPin[i]) > 64) bad++;
        }
        if (prefix != "-") {
          std::vector<unsigned char> px(A);
          for (int i = 0; i < A; ++i) px[i] = (unsigned char)hOut[i];
          dump(prefix + (frames > 0 ? ".f" + std::to_string(f) : "") + ".dec", px.data(), A);
        }
      }
    }
  }
  const double mpix = (double)W * H / 1e6;
  char buf[4096];
  std::string lens;
  for (size_t i = 0; i < frameLens.size(); ++i) lens += (i ? "," : "") + std::to_string(frameLens[i]);
  snprintf(buf, sizeof buf,
           "{\"W\":%d,\"H\":%d,\"wl\":%d,\"lossy\":%d,\"qs\":%g,\"frames\":%d,\"iters\":%d,\"samples\":%d,"
           "\"roundtrip_bad\":%d,\"frame_len\":[%s],"
           "\"enc_dev_ms\":%.4f,\"enc_dev_ms_best\":%.4f,\"enc_dwt_ms\":%.4f,\"enc_bpc_bsb_ms\":%.4f,\"enc_e2e_ms\":%.4f,"
           "\"dec_dev_ms\":%.4f,\"dec_dev_ms_best\":%.4f,\"dec_bpc_bsb_ms\":%.4f,\"dec_dwt_ms\":%.4f,\"dec_e2e_ms\":%.4f,"
           "\"enc_dev_mpix_s\":%.2f,\"dec_dev_mpix_s\":%.2f,\"enc_e2e_mpix_s\":%.2f,\"dec_e2e_mpix_s\":%.2f}",
           W, H, wl, (int)lossy, qs, frames, iters, sEnc.n, bad, lens.c_str(),
           sEnc.sum / sEnc.n, sEnc.best, sDwtE.sum / sDwtE.n, sBpcE.sum / sBpcE.n, sEncE2E.sum / sEncE2E.n,
           sDec.sum / sDec.n, sDec.best, sBpcD.sum / sBpcD.n, sDwtD.sum / sDwtD.n, sDecE2E.sum / sDecE2E.n,
           mpix / (sEnc.sum / sEnc.n) * 1e3, mpix / (sDec.sum / sDec.n) * 1e3,
           mpix / (sEncE2E.sum / sEncE2E.n) * 1e3, mpix / (sDecE2E.sum / sDecE2E.n) * 1e3);
  printf("%s\n", buf);
  if (!jsonOut.empty()) { FILE* f = fopen(jsonOut.c_str(), "w"); if (f) { fprintf(f, "%s\n", buf); fclose(f); } }
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc < 11) {
    fprintf(stderr, "usage: %s W H wl type qs LUTFolder/ input.raw frames iters dump_prefix|- [json_out]\n", argv[0]);
    return 64;
  }
  int W = atoi(argv[1]), H = atoi(argv[2]), wl = atoi(argv[3]), type = atoi(argv[4]);
  float qs = (float)atof(argv[5]);
  std::string lut = argv[6], input = argv[7];
  int frames = atoi(argv[8]), iters = atoi(argv[9]);
  std::string prefix = argv[10], jsonOut = argc > 11 ? argv[11] : "", coefFile;
  if (argc > 12 && !strncmp(argv[12], "coef=", 5)) coefFile = argv[12] + 5;
  cudaDeviceSetCacheConfig(cudaFuncCachePreferL1); /* Launcher.cu:188 */
  if (type) return run<float, float2>(W, H, wl, true, qs, lut, input, frames, iters, prefix, jsonOut);
  return run<int, int2>(W, H, wl, false, qs, lut, input, frames, iters, prefix, jsonOut, coefFile);
}
