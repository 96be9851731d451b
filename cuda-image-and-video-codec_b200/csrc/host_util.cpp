// host_util.cpp -- host-only parts of the C ABI: geometry, LUT folder loader, codestream header.
// (No CUDA here; these are what the reference does in IO/IOManager.ipp, Engines/Engine.cu,
// BitStreamBuilder/BitStreamBuilder.cpp and Engines/DecodingEngine.cu on the host.)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/picsong_b200.h"

extern "C" {

// SupportFunctions/AuxiliarFunctions.cpp:22-26: both dimensions are padded up to a multiple of 64.
void pcs_adapted_dims(int32_t width, int32_t height, int32_t* adapted_w, int32_t* adapted_h) {
  *adapted_w = (width + PCS_CB - 1) / PCS_CB * PCS_CB;
  *adapted_h = (height + PCS_CB - 1) / PCS_CB * PCS_CB;
}

// The same with the codeblock length the reference was built with (CBLOCK_LENGTH, BPC/BPCEngine.cuh:33-36): the height
// is padded to a multiple of the block length, the width always to 64.
void pcs_adapted_dims_cb(int32_t width, int32_t height, int32_t cb_rows, int32_t* adapted_w, int32_t* adapted_h) {
  const int r = cb_rows == 32 ? 32 : PCS_CB;
  *adapted_w = (width + PCS_CB - 1) / PCS_CB * PCS_CB;
  *adapted_h = (height + r - 1) / r * r;
}

int32_t pcs_lut_size(const int32_t h[8], int32_t wl) {
  const int nBp = h[0], nSub = h[1], cR = h[2], cG = h[3], cS = h[4];
  return (nSub * wl + 1) * nBp * (cR + cS + cG);
}

}  // extern "C"

namespace {

// header.txt: eight `NAME;value` lines (Engines/Engine.cu:190-210, IO/IOManager.ipp:363-386)
bool read_lut_header(const std::string& folder, int32_t h[8]) {
  std::ifstream f(folder + "/header.txt");
  if (!f) return false;
  std::string line;
  int i = 0;
  while (i < 8 && std::getline(f, line)) {
    const size_t semi = line.find(';');
    if (semi == std::string::npos) continue;
    h[i++] = std::atoi(line.c_str() + semi + 1);
  }
  return i == 8;
}

// One section of a table file set (IO/IOManager.ipp:446-536).  Rows are "lvl sb bp : v0 v1 ...".
// The file lists (level, subband) groups in order; a group is recognised by its bit-plane number
// not increasing.  When a new group starts, the planes the finished group did not list are filled
// with 64 (= 1/2 at 7-bit precision) and the write base moves on by nBp*ctx.  Reading stops at the
// first row of a group with  lvl+1 > wl && sb > 0  -- after the fill-and-advance, so a group that is
// cut off by the stop rule is still filled, one that ends with the file is not.
bool read_lut_section(const std::string& path, std::vector<int32_t>& t, int base, int nBp, int ctx, int wl) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line;
  int group_base = base, last_bp = -1;
  while (std::getline(f, line)) {
    const size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    std::istringstream lhs(line.substr(0, colon)), rhs(line.substr(colon + 1));
    int lvl, sb, bp;
    if (!(lhs >> lvl >> sb >> bp)) continue;
    std::vector<int> v(ctx);
    bool ok = true;
    for (int c = 0; c < ctx; ++c) ok = ok && static_cast<bool>(rhs >> v[c]);
    if (!ok) break;
    if (bp <= last_bp) {
      for (int b = last_bp + 1; b < nBp; ++b)
        for (int c = 0; c < ctx; ++c) {
          const int idx = group_base + b * ctx + c;
          if (idx >= 0 && idx < (int)t.size()) t[idx] = 64;
        }
      group_base += nBp * ctx;
    }
    if (lvl + 1 > wl && sb > 0) break;
    last_bp = bp;
    for (int c = 0; c < ctx; ++c) {
      const int idx = group_base + bp * ctx + c;
      if (idx >= 0 && idx < (int)t.size()) t[idx] = v[c];
    }
  }
  return true;
}

}  // namespace

extern "C" {

// Greyscale / k == 0 / 2 coding passes: with three component files the suffixes are R,G,B and grey
// uses "R" (Engines/Engine.cu:124-139, Engines/CodingEngine.cu:662); suffix _0 is the k == 0 set.
int32_t pcs_lut_load(const char* folder, int32_t wl, int32_t header[8], int32_t* table, int32_t capacity) {
  return pcs_lut_load_component(folder, wl, 0, header, table, capacity);
}

// RGB path: component i is coded with the i-th table set, files "R", "G", "B" (Engines/Engine.cu:124-139,
// Engines/CodingEngine.cu:616,696,790).
static int32_t lut_load_file(const char* folder, int32_t wl, int32_t component, int file, int32_t header[8], int32_t* table, int32_t capacity);
int32_t pcs_lut_load_component(const char* folder, int32_t wl, int32_t component, int32_t header[8], int32_t* table, int32_t capacity) {
  return lut_load_file(folder, wl, component, 0, header, table, capacity);
}

// k > 0 (Engines/Engine.cu:12-100): the tables of the bit-plane files `_0 .. _(AMOUNT_OF_BITPLANE_FILES - 1)` of one
// component back to back, pcs_lut_size() ints each; BPCEngine::initializeLUTPointers selects set s by adding s * size
// (BPC/BPCEngine.cu:333).  The file index only enters the file name (IO/IOManager.ipp:445-447).
int32_t pcs_lut_load_k(const char* folder, int32_t wl, int32_t component, int32_t header[8], int32_t* table, int32_t capacity) {
  if (!folder) return PCS_E_ARG;
  if (!read_lut_header(std::string(folder), header)) return PCS_E_LUT;
  const int files = header[7], n = pcs_lut_size(header, wl);
  if (files < 1 || files > 32 || n <= 0) return PCS_E_LUT;
  if ((long long)n * files > capacity) return PCS_E_NOMEM;
  for (int j = 0; j < files; ++j) {
    const int32_t r = lut_load_file(folder, wl, component, j, header, table + (size_t)j * n, n);
    if (r < 0) return r;
  }
  return n * files;
}

static int32_t lut_load_file(const char* folder, int32_t wl, int32_t component, int file, int32_t header[8], int32_t* table, int32_t capacity) {
  if (!folder || wl < 1 || wl > 10 || component < 0 || component > 2) return PCS_E_ARG;
  const std::string dir(folder);
  if (!read_lut_header(dir, header)) return PCS_E_LUT;
  const int nBp = header[0], nSub = header[1], cR = header[2], cG = header[3], cS = header[4];
  if (nBp < 1 || nSub < 1 || cR < 1 || cG < 1 || cS < 1) return PCS_E_LUT;
  const int n = pcs_lut_size(header, wl);
  if (n > capacity) return PCS_E_NOMEM;
  std::vector<int32_t> t(n, 0);  // undefined entries are 0 (the reference relies on fresh zero pages)
  const int R = (nSub * wl + 1) * nBp * cR, S = (nSub * wl + 1) * nBp * cS;
  static const char* const kComp[3] = {"R", "G", "B"};
  const std::string comp = header[6] == 1 ? "" : kComp[component];
  const std::string sfx = ".txt_" + std::to_string(file);
  if (!read_lut_section(dir + "ref" + comp + sfx, t, 0, nBp, cR, wl)) return PCS_E_LUT;
  if (!read_lut_section(dir + "sig" + comp + sfx, t, R, nBp, cS, wl)) return PCS_E_LUT;
  if (!read_lut_section(dir + "sign" + comp + sfx, t, R + S, nBp, cG, wl)) return PCS_E_LUT;
  std::memcpy(table, t.data(), (size_t)n * sizeof(int32_t));
  return n;
}

// BitStreamBuilder::setExtraInformation (BitStreamBuilder/BitStreamBuilder.cpp:35-94)
void pcs_header_pack(const pcs_params* p, uint16_t out[9]) {
  const int sz = p->width * p->height * p->components;
  const int qi = (int)(p->quant_size * 10000);
  const int wl = p->wavelet_levels, H = p->height, comp = p->components, fr = p->frames;
  out[0] = (uint16_t)(sz & 0xFFFF);
  out[1] = (uint16_t)(((uint32_t)sz >> 16) & 0xFFFF);
  out[2] = (uint16_t)((p->coding_passes == 2 ? 0 : 1) | (p->cb_height << 1) | (p->cb_width << 8) | ((wl & 1) << 15));
  out[3] = (uint16_t)(((wl & 7) >> 1) | (p->bit_depth << 3) | ((p->lossy ? 1 : 0) << 10) | ((qi & 31) << 11));
  out[4] = (uint16_t)((qi >> 5) | ((comp & 127) << 9));
  out[5] = (uint16_t)((comp >> 7) | ((p->is_rgb ? 1 : 0) << 7) | (H << 8));
  out[6] = (uint16_t)((H >> 8) | (p->endianess << 8) | (p->bits_per_sample << 9) | ((p->is_signed ? 1 : 0) << 14) | ((fr & 1) << 15));
  out[7] = (uint16_t)((fr >> 1) & 0xFFFF);
  out[8] = (uint16_t)((int)(p->k_factor * 1000));
}

// DecodingEngine::getExtraInformation (Engines/DecodingEngine.cu:567-585) and the way
// DecodingEngine::initMemory consumes it (:146-162): width = size / height, qs = int / 10000.
int32_t pcs_header_unpack(const uint16_t e[9], pcs_params* p) {
  const int size = e[0] | (e[1] << 16);
  const int H = ((e[5] >> 8) & 255) | ((e[6] & 255) << 8);
  if (H <= 0 || size <= 0) return PCS_E_STREAM;
  p->coding_passes = (e[2] & 1) ? 3 : 2;
  p->cb_height = (e[2] >> 1) & 127;
  p->cb_width = (e[2] >> 8) & 127;
  p->wavelet_levels = ((e[2] >> 15) & 1) | ((e[3] & 7) << 1);
  p->bit_depth = (e[3] >> 3) & 127;
  p->lossy = (e[3] >> 10) & 1;
  p->quant_size = (float)((((e[3] >> 11) & 31) | ((e[4] & 511) << 5)) / 10000.0);
  p->components = ((e[4] >> 9) & 127) | ((e[5] & 127) << 9);
  p->is_rgb = (e[5] >> 7) & 1;
  p->height = H;
  p->width = size / H;
  p->endianess = (e[6] >> 8) & 1;
  p->bits_per_sample = (e[6] >> 9) & 31;
  p->is_signed = (e[6] >> 14) & 1;
  p->frames = ((e[6] >> 15) & 1) | (e[7] << 1);
  p->k_factor = (float)(e[8] / 1000.0);
  return PCS_OK;
}

}  // extern "C"
