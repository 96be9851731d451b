// engines.cpp -- CodingEngine / DecodingEngine over the C ABI.  Host work only: RAW frame reads
// (IO/IOManager.ipp:72-112; the mirror padding itself happens on the GPU inside pcs_dwt_forward),
// codestream + `_SIZE` sidecar (IOManager.ipp:176-208,615-620), PGM / RAW output (:214-231,267-344).
#include "engines.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

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
  // greyscale codes with the "R" table set; the RGB path codes component i with set i (Engines/Engine.cu:124-139)
  std::vector<int32_t> table(3 << 16);
  int n = 0;
  for (int comp = 0; comp < (_img->isRGB ? 3 : 1); ++comp) {
    const int m = pcs_lut_load_component(_lutPath.c_str(), _waveletLevels, comp, p.lut_header, table.data() + n, (int)table.size() - n);
    if (m <= 0) die("Error opening LUT files in " + _lutPath);
    n += m;
  }
  p.lut = table.data(); p.lut_len = n;
  pcs_ctx* ctx = nullptr;
  const int rc = pcs_create(&p, &ctx);
  if (rc == PCS_E_ARG) die("Incorrect parameters. Please choose valid values.", -1);
  if (rc == PCS_E_UNSUPPORTED) die("This build covers the 8-bit unsigned 2-coding-pass k=0 paths, grey and RGB (cp=3, k>0 and signed input are out of scope).", -1);
  if (rc != 0) die("pcs_create failed with code " + std::to_string(rc) + (rc > 0 ? " (CUDA error: is a GPU present? there is no CPU fallback)" : ""));
  return ctx;
}

static size_t batchFor(const Image* img, int frames) {
  // frames per launch: enough codeblocks to fill a B200 (>= ~9500 resident warps), bounded by memory
  const size_t px = (size_t)img->width * img->height * (img->isRGB ? 3 : 1);
  size_t b = std::max<size_t>(1, (size_t)(40u << 20) / std::max<size_t>(px, 1));
  return std::min<size_t>(std::max<size_t>(b, 1), (size_t)std::max(frames, 1));
}

void CodingEngine::engineManager(int cType) {
  const bool video = cType == 1;
  const int frames = video ? _numberOfFrames : 1;
  if (video && frames <= 0) die("Incorrect parameters. Please choose valid values.", -1);
  const int planes = _img->isRGB ? 3 : 1;  // RGB input: planar R, G, B planes per frame (IOManager.ipp:72-112 with iter = 3f+c)
  const size_t fpx = (size_t)_img->width * _img->height * planes;
  const int batch = (int)batchFor(_img, frames);
  pcs_ctx* ctx = makeContext(batch, 0);
  std::ifstream in(_img->name, std::ios::binary);
  if (!in) die("Error opening input file " + _img->name);
  std::remove(_outputFile.c_str());                       // replaceExistingFile, IOManager.ipp:158-170
  std::remove((_outputFile + "_SIZE").c_str());
  std::ofstream out(_outputFile, std::ios::binary | std::ios::trunc);
  std::vector<uint8_t> px((size_t)batch * fpx);
  std::vector<uint16_t> stream((size_t)batch * pcs_frame_max_stream_words(ctx));
  std::vector<int32_t> lens((size_t)batch * planes);
  std::ostringstream sizes;
  double gpuSeconds = 0;
  for (int f0 = 0; f0 < frames; f0 += batch) {
    const int b = std::min(batch, frames - f0);
    in.read(reinterpret_cast<char*>(px.data()), (std::streamsize)((size_t)b * fpx));
    if ((size_t)in.gcount() != (size_t)b * fpx) die("Input file is shorter than xSize*ySize*frames");
    size_t total = 0;
    auto t0 = std::chrono::steady_clock::now();
    const int rc = pcs_encode_frames(ctx, px.data(), b, f0, stream.data(), stream.size(), lens.data(), &total);
    gpuSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc != 0) die(std::string("encode failed: ") + pcs_last_error(ctx) + " (" + std::to_string(rc) + ")");
    out.write(reinterpret_cast<const char*>(stream.data()), (std::streamsize)(total * 2));
    for (int i = 0; i < b * planes; ++i) sizes << (f0 + i ? "," : "") << lens[i];
  }
  out.close();
  if (video || planes == 3) {  // grey stills write no sidecar (IOManager.ipp:615-620); RGB stills do (CodingEngine.cu:566-577)
    std::ofstream s(_outputFile + "_SIZE", std::ios::binary | std::ios::trunc);
    s << sizes.str();
  }
  if (video || planes == 3) {
    std::cout << "The time spent with the app without considering allocation periods is: " << gpuSeconds << std::endl;
  }
  std::cout << "BPC acum time is: " << gpuSeconds << std::endl;  // whole GPU pipeline incl. PCIe (stages are fused)
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
  pcs_ctx* ctx = makeContext(batch, std::getenv("PICSONG_PROPER_CROP") ? 0u : PCS_F_REF_FLAT_CROP);
  std::vector<int32_t> lens((size_t)frames * planes);
  in.seekg(0, std::ios::end);
  const size_t fileWords = (size_t)in.tellg() / 2;
  if (video || planes == 3) {  // lengths from the sidecar (IOManager.ipp:196-208; RGB stills: DecodingEngine.cu:371-383)
    std::ifstream s(_img->name + "_SIZE");
    std::string tok;
    for (int f = 0; f < frames * planes; ++f) {
      if (!std::getline(s, tok, ',')) die("_SIZE sidecar is missing or too short");
      lens[f] = std::stoi(tok);
    }
  } else {
    lens[0] = (int32_t)fileWords;
  }
  in.seekg(0);
  std::remove(_outputFile.c_str());
  std::ofstream out(_outputFile, std::ios::binary | std::ios::trunc);
  if (!video && planes == 1) out << "P5\n" << _img->width << " " << _img->height << "\n" << ((2 << (_img->bitDepth - 1)) - 1) << "\n";
  std::vector<uint16_t> stream;
  std::vector<uint8_t> px((size_t)batch * fpx);
  double gpuSeconds = 0;
  for (int f0 = 0; f0 < frames; f0 += batch) {
    const int b = std::min(batch, frames - f0);
    size_t words = 0;
    for (int i = 0; i < b * planes; ++i) words += (size_t)lens[(size_t)f0 * planes + i];
    stream.resize(words);
    in.read(reinterpret_cast<char*>(stream.data()), (std::streamsize)(words * 2));
    if ((size_t)in.gcount() != words * 2) die("Codestream is shorter than its frame lengths");
    auto t0 = std::chrono::steady_clock::now();
    const int rc = pcs_decode_frames(ctx, stream.data(), lens.data() + (size_t)f0 * planes, b, px.data());
    gpuSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc != 0) die(std::string("decode failed: ") + pcs_last_error(ctx) + " (" + std::to_string(rc) + ")");
    out.write(reinterpret_cast<const char*>(px.data()), (std::streamsize)((size_t)b * fpx));
  }
  std::cout << "BPC acum time is: " << gpuSeconds << std::endl;
  pcs_destroy(ctx);
}

}  // namespace picsong
