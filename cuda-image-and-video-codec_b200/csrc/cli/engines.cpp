// engines.cpp -- CodingEngine / DecodingEngine over the C ABI.  Host work only: RAW frame reads
// (IO/IOManager.ipp:72-112; the mirror padding itself happens on the GPU inside pcs_dwt_forward),
// codestream + `_SIZE` sidecar (IOManager.ipp:176-208,615-620), PGM / RAW output (:214-231,267-344).
#include "engines.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <thread>

namespace picsong {

void Engine::die(const std::string& msg, int code) {
  std::cout << msg << std::endl;
  std::exit(code);
}

pcs_ctx* Engine::makeContext(int maxBatch, unsigned flags) {
  pcs_params p{};
  p.width = _img->width; p.height = _img->height; p.wavelet_levels = _waveletLevels; p.lossy = _lossy ? 1 : 0;
  p.quant_size = _qs; p.coding_passes = _codingPasses; p.k_factor = _k; p.cb_width = _cbWidth; p.cb_height = _cbHeight;
  p.frames = _numberOfFrames; p.bit_depth = _img->bitDepth; p.bits_per_sample = _img->bitsPerSample;
  p.endianess = _img->endianess; p.is_signed = _img->signedOrUnsigned; p.components = _img->components; p.is_rgb = _img->isRGB;
  p.max_batch = maxBatch; p.flags = flags;
  if (const char* e = std::getenv("PICSONG_CB_LENGTH")) if (std::atoi(e) == 32) p.flags |= PCS_F_CB_LENGTH_32;  // the reference: -DCBLOCK_LENGTH=32 at build time
  // greyscale codes with the "R" table set; the RGB path codes component i with set i (Engines/Engine.cu:124-139);
  // k > 0 loads every bit-plane file of a component (Engines/Engine.cu:12-100)
  auto t0 = std::chrono::steady_clock::now();
  std::vector<int32_t> table(_k > 0 ? (3 << 20) : (3 << 16));
  int n = 0;
  for (int comp = 0; comp < (_img->isRGB ? 3 : 1); ++comp) {
    const int m = _k > 0 ? pcs_lut_load_k(_lutPath.c_str(), _waveletLevels, comp, p.lut_header, table.data() + n, (int)table.size() - n)
                         : pcs_lut_load_component(_lutPath.c_str(), _waveletLevels, comp, p.lut_header, table.data() + n, (int)table.size() - n);
    if (m <= 0) die("Error opening LUT files in " + _lutPath);
    n += m;
  }
  p.lut = table.data(); p.lut_len = n;
  _tLut = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  t0 = std::chrono::steady_clock::now();
  pcs_host_free(pcs_host_alloc(16));  // first CUDA call of the process: creates the context (timed apart from our own set-up)
  _tCuda = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  t0 = std::chrono::steady_clock::now();
  pcs_ctx* ctx = nullptr;
  const int rc = pcs_create(&p, &ctx);
  _tCreate = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc == PCS_E_ARG) die("Incorrect parameters. Please choose valid values.", -1);
  if (rc == PCS_E_UNSUPPORTED) die("This build covers the 2-coding-pass paths: grey (any -bps up to 16, signed or unsigned byte samples) and 8-bit RGB, k >= 0 (-cp 3 is out of scope: the reference ships no LUT files for it).", -1);
  if (rc != 0) die("pcs_create failed with code " + std::to_string(rc) + (rc > 0 ? " (CUDA error: is a GPU present? there is no CPU fallback)" : ""));
  return ctx;
}

static size_t batchFor(const Image* img, int frames) {
  // frames per launch: enough codeblocks to fill a B200 (>= ~9500 resident warps), bounded by memory
  const size_t px = (size_t)img->width * img->height * (img->isRGB ? 3 : 1);
  size_t b = std::max<size_t>(1, (size_t)(128u << 20) / std::max<size_t>(px, 1));
  if (const char* e = std::getenv("PICSONG_BATCH")) b = (size_t)std::max(1, std::atoi(e));  // tests: force many small batches
  return std::min<size_t>(std::max<size_t>(b, 1), (size_t)std::max(frames, 1));
}

static double secondsSince(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- host pipeline (replaces readFileParallel / writeFileParallel and their flag-polled double
// buffers, Engines/CodingEngine.cu:212-262,463-498, Engines/DecodingEngine.cu:241-298,441-520) ----------------
// A reader thread fills pinned buffers from disk, the caller's thread runs the GPU calls, a writer thread drains
// pinned output buffers to disk; kRing buffers of one batch each circulate through two bounded queues, so disk
// reads of batch k+1, the GPU work of batch k and the disk write of batch k-1 overlap.
namespace {

constexpr int kRing = 3;

template <class T>
class Channel {  // bounded multi-producer/multi-consumer queue of small tokens
 public:
  void push(T v) { { std::lock_guard<std::mutex> l(m_); q_.push_back(v); } cv_.notify_all(); }
  T pop() { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return !q_.empty(); }); T v = q_.front(); q_.pop_front(); return v; }
 private:
  std::mutex m_; std::condition_variable cv_; std::deque<T> q_;
};

struct Batch { int slot, f0, n; size_t words; };  // n == 0: end of stream; n < 0: the producer failed

// Page-locking memory costs ~0.3 ms per MB, more than coding the frames that go through it, so a buffer is allocated the
// first time its slot is used (by the one thread that owns the slot at that moment): a short video never pays for slots
// it does not reach, and the allocation of slot k+1 overlaps the GPU work on slot k.
// Short jobs skip page-locking altogether: a job that moves less than ~1.5 GB is done sooner through pageable buffers (the
// driver stages the copies at 10-20 GB/s) than the 0.3-0.7 s it takes to pin the ring.
struct PinnedRing {
  std::vector<uint8_t*> in; std::vector<uint16_t*> out; std::vector<std::vector<int32_t>> lens;
  size_t inBytes, outBytes;
  bool pin;
  std::atomic<long long> allocMicros{0};
  PinnedRing(size_t inBytes_, size_t outWords, size_t nlens, bool pin_)
      : in(kRing, nullptr), out(kRing, nullptr), inBytes(inBytes_), outBytes(outWords * 2), pin(pin_) {
    for (int i = 0; i < kRing; ++i) lens.emplace_back(nlens);
  }
  template <class P> P* get(std::vector<P*>& v, int i, size_t bytes) {
    if (!v[i]) {
      auto t0 = std::chrono::steady_clock::now();
      v[i] = static_cast<P*>(pin ? pcs_host_alloc(bytes) : std::aligned_alloc(4096, (bytes + 4095) / 4096 * 4096));
      allocMicros += (long long)(secondsSince(t0) * 1e6);
    }
    return v[i];
  }
  uint8_t* inBuf(int i) { return get(in, i, inBytes); }
  uint16_t* outBuf(int i) { return get(out, i, outBytes); }
  ~PinnedRing() {
    for (auto p : in) { if (pin) pcs_host_free(p); else std::free(p); }
    for (auto p : out) { if (pin) pcs_host_free(p); else std::free(p); }
  }
};
static bool worthPinning(size_t totalPixelBytes) {
  if (const char* e = std::getenv("PICSONG_PIN")) return std::atoi(e) != 0;
  return totalPixelBytes >= (size_t)1536 << 20;
}

}  // namespace

void CodingEngine::engineManager(int cType) {
  const bool video = cType == 1;
  const int frames = video ? _numberOfFrames : 1;
  if (video && frames <= 0) die("Incorrect parameters. Please choose valid values.", -1);
  const int planes = _img->isRGB ? 3 : 1;  // RGB input: planar R, G, B planes per frame (IOManager.ipp:72-112 with iter = 3f+c)
  const size_t fpx = (size_t)_img->width * _img->height * planes;
  const int batch = (int)batchFor(_img, frames);
  auto tStart = std::chrono::steady_clock::now();
  pcs_ctx* ctx = makeContext(batch, video ? 0u : PCS_F_STILLS);  // stills: only plane 0 carries the header, whatever -frames says
  const double tContext = secondsSince(tStart);
  std::ifstream in(_img->name, std::ios::binary);
  if (!in) die("Error opening input file " + _img->name);
  std::remove(_outputFile.c_str());                       // replaceExistingFile, IOManager.ipp:158-170
  std::remove((_outputFile + "_SIZE").c_str());
  std::ofstream out(_outputFile, std::ios::binary | std::ios::trunc);
  if (!out) die("Error opening output file " + _outputFile);
  const size_t maxWords = (size_t)batch * pcs_frame_max_stream_words(ctx);
  PinnedRing ring((size_t)batch * fpx, maxWords, (size_t)batch * planes, worthPinning((size_t)frames * fpx));
  Channel<int> freeIn, freeOut;
  Channel<Batch> filled, coded;
  for (int i = 0; i < kRing; ++i) { freeIn.push(i); freeOut.push(i); }
  std::string readErr;
  std::thread reader([&] {
    for (int f0 = 0; f0 < frames; f0 += batch) {
      const int b = std::min(batch, frames - f0), slot = freeIn.pop();
      if (!ring.inBuf(slot)) { readErr = "pinned host allocation failed"; filled.push({slot, f0, -1, 0}); return; }
      in.read(reinterpret_cast<char*>(ring.inBuf(slot)), (std::streamsize)((size_t)b * fpx));
      if ((size_t)in.gcount() != (size_t)b * fpx) { readErr = "Input file is shorter than xSize*ySize*frames"; filled.push({slot, f0, -1, 0}); return; }
      filled.push({slot, f0, b, 0});
    }
    filled.push({0, 0, 0, 0});
  });
  std::ostringstream sizes;
  std::thread writer([&] {
    for (;;) {
      const Batch w = coded.pop();
      if (w.n <= 0) return;
      out.write(reinterpret_cast<const char*>(ring.outBuf(w.slot)), (std::streamsize)(w.words * 2));
      for (int i = 0; i < w.n * planes; ++i) sizes << (w.f0 + i ? "," : "") << ring.lens[w.slot][i];
      freeOut.push(w.slot);
    }
  });
  double gpuSeconds = 0;
  std::string err;
  for (;;) {
    const Batch r = filled.pop();
    if (r.n <= 0) { if (r.n < 0) err = readErr; break; }
    const int oslot = freeOut.pop();
    if (!ring.outBuf(oslot)) { err = "pinned host allocation failed"; break; }
    size_t total = 0;
    auto t0 = std::chrono::steady_clock::now();
    const int rc = pcs_encode_frames(ctx, ring.inBuf(r.slot), r.n, r.f0, ring.outBuf(oslot), maxWords, ring.lens[oslot].data(), &total);
    gpuSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    freeIn.push(r.slot);
    if (rc != 0) { err = std::string("encode failed: ") + pcs_last_error(ctx) + " (" + std::to_string(rc) + ")"; break; }
    coded.push({oslot, r.f0, r.n, total});
  }
  coded.push({0, 0, 0, 0});
  writer.join();
  if (!err.empty()) { reader.detach(); die(err); }
  reader.join();
  out.close();
  if (video || planes == 3) {  // grey stills write no sidecar (IOManager.ipp:615-620); RGB stills do (CodingEngine.cu:566-577)
    std::ofstream s(_outputFile + "_SIZE", std::ios::binary | std::ios::trunc);
    s << sizes.str();
    std::cout << "The time spent with the app without considering allocation periods is: " << gpuSeconds << std::endl;
  }
  std::cout << "BPC acum time is: " << gpuSeconds << std::endl;  // whole GPU pipeline incl. PCIe (stages are fused)
  std::cout << "Start-up (CUDA context, LUT parsing, device buffers): " << tContext << " s; " << (ring.pin ? "page-locked" : "pageable (short job)") << " host buffers: "
            << ring.allocMicros.load() * 1e-6 << " s (overlapped with coding after the first batch)" << std::endl;
  std::cout << "Start-up detail: LUT parsing " << _tLut << " s, CUDA context creation " << _tCuda << " s, tables + device buffers " << _tCreate << " s" << std::endl;
  pcs_destroy(ctx);
}

void DecodingEngine::engineManager(int cType) {
  const bool video = cType == 1;
  std::ifstream in(_img->name, std::ios::binary);
  if (!in) die("Error opening input file " + _img->name);
  uint16_t hdr[9];
  in.read(reinterpret_cast<char*>(hdr), sizeof hdr);
  pcs_params hp{};
  if ((size_t)in.gcount() != sizeof hdr || pcs_header_unpack(hdr, &hp) != 0) die("Malformed codestream header");
  _img->width = hp.width; _img->height = hp.height; _img->bitDepth = hp.bit_depth; _img->components = hp.components;
  _img->bitsPerSample = hp.bits_per_sample; _img->endianess = hp.endianess; _img->isRGB = hp.is_rgb; _img->signedOrUnsigned = hp.is_signed;
  setCodingPasses(hp.coding_passes); setCBHeight(hp.cb_height); setCBWidth(hp.cb_width); setWaveletLevels(hp.wavelet_levels);
  setWType(hp.lossy != 0); setQSizes(hp.quant_size); setNumberOfFrames(hp.frames); setKFactor(hp.k_factor);
  const int frames = video ? _numberOfFrames : 1;
  const int planes = _img->isRGB ? 3 : 1;
  const size_t fpx = (size_t)_img->width * _img->height * planes;
  const int batch = (int)batchFor(_img, frames);
  // PCS_F_REF_FLAT_CROP keeps the reference's output byte order when the width is not a multiple of 64
  auto tStart = std::chrono::steady_clock::now();
  pcs_ctx* ctx = makeContext(batch, (std::getenv("PICSONG_PROPER_CROP") ? 0u : PCS_F_REF_FLAT_CROP) | (video ? 0u : PCS_F_STILLS));
  const double tContext = secondsSince(tStart);
  std::vector<int32_t> lens((size_t)frames * planes);
  in.seekg(0, std::ios::end);
  const size_t fileWords = (size_t)in.tellg() / 2;
  if (video || planes == 3) {  // lengths from the sidecar (IOManager.ipp:196-208; RGB stills: DecodingEngine.cu:371-383)
    std::ifstream s(_img->name + "_SIZE");
    std::string tok;
    for (int f = 0; f < frames * planes; ++f) {
      if (!std::getline(s, tok, ',')) die("_SIZE sidecar is missing or too short");
      char* end = nullptr;
      const long v = std::strtol(tok.c_str(), &end, 10);
      if (end == tok.c_str() || v <= 0 || v > 0x7fffffffL) die("_SIZE sidecar malformed");
      lens[f] = (int32_t)v;
    }
  } else {
    lens[0] = (int32_t)fileWords;
  }
  in.seekg(0);
  std::remove(_outputFile.c_str());
  std::ofstream out(_outputFile, std::ios::binary | std::ios::trunc);
  if (!out) die("Error opening output file " + _outputFile);
  if (!video && planes == 1) out << "P5\n" << _img->width << " " << _img->height << "\n" << ((2 << (_img->bitDepth - 1)) - 1) << "\n";
  const size_t maxWords = (size_t)batch * pcs_frame_max_stream_words(ctx);
  // here the ring's "in" side holds pixels (output of the GPU) and its "out" side codestream words (input)
  PinnedRing ring((size_t)batch * fpx, maxWords, 1, worthPinning((size_t)frames * fpx));
  Channel<int> freeStream, freePx;
  Channel<Batch> filled, decoded;
  for (int i = 0; i < kRing; ++i) { freeStream.push(i); freePx.push(i); }
  std::string readErr;
  std::thread reader([&] {
    for (int f0 = 0; f0 < frames; f0 += batch) {
      const int b = std::min(batch, frames - f0), slot = freeStream.pop();
      size_t words = 0;
      for (int i = 0; i < b * planes; ++i) words += (size_t)std::max(lens[(size_t)f0 * planes + i], 0);
      if (words > maxWords) { readErr = "frame lengths exceed the maximum stream size"; filled.push({slot, f0, -1, 0}); return; }
      if (!ring.outBuf(slot)) { readErr = "pinned host allocation failed"; filled.push({slot, f0, -1, 0}); return; }
      in.read(reinterpret_cast<char*>(ring.outBuf(slot)), (std::streamsize)(words * 2));
      if ((size_t)in.gcount() != words * 2) { readErr = "Codestream is shorter than its frame lengths"; filled.push({slot, f0, -1, 0}); return; }
      filled.push({slot, f0, b, words});
    }
    filled.push({0, 0, 0, 0});
  });
  std::thread writer([&] {
    for (;;) {
      const Batch w = decoded.pop();
      if (w.n <= 0) return;
      out.write(reinterpret_cast<const char*>(ring.inBuf(w.slot)), (std::streamsize)((size_t)w.n * fpx));
      freePx.push(w.slot);
    }
  });
  double gpuSeconds = 0;
  std::string err;
  for (;;) {
    const Batch r = filled.pop();
    if (r.n <= 0) { if (r.n < 0) err = readErr; break; }
    const int pslot = freePx.pop();
    if (!ring.inBuf(pslot)) { err = "pinned host allocation failed"; break; }
    auto t0 = std::chrono::steady_clock::now();
    const int rc = pcs_decode_frames(ctx, ring.outBuf(r.slot), lens.data() + (size_t)r.f0 * planes, r.n, ring.inBuf(pslot));
    gpuSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    freeStream.push(r.slot);
    if (rc != 0) { err = std::string("decode failed: ") + pcs_last_error(ctx) + " (" + std::to_string(rc) + ")"; break; }
    decoded.push({pslot, r.f0, r.n, 0});
  }
  decoded.push({0, 0, 0, 0});
  writer.join();
  if (!err.empty()) { reader.detach(); die(err); }
  reader.join();
  if (video || planes == 3)
    std::cout << "The time spent with the app without considering allocation periods is: " << gpuSeconds << std::endl;
  std::cout << "BPC acum time is: " << gpuSeconds << std::endl;
  std::cout << "Start-up (CUDA context, LUT parsing, device buffers): " << tContext << " s; " << (ring.pin ? "page-locked" : "pageable (short job)") << " host buffers: "
            << ring.allocMicros.load() * 1e-6 << " s (overlapped with decoding after the first batch)" << std::endl;
  std::cout << "Start-up detail: LUT parsing " << _tLut << " s, CUDA context creation " << _tCuda << " s, tables + device buffers " << _tCreate << " s" << std::endl;
  pcs_destroy(ctx);
}

}  // namespace picsong
