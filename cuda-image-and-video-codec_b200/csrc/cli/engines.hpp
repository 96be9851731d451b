// engines.hpp -- host-side mirror of the reference's Engine surface (Engines/Engine.cuh:28-44 of
// 13Karl/CUDA-Image-and-Video-codec): the same setters, then engineManager(0 = image, 1 = video).
// Everything below the setters goes through the C ABI (include/picsong_b200.h); this layer only does
// what the reference does on the host: file I/O in the reference's formats and the stdout lines.
#pragma once
#include <string>
#include <vector>

#include "../../../include/picsong_b200.h"

namespace picsong {

// Image/Image.hpp:8-65
struct Image {
  std::string name;
  int width = 0, height = 0, components = 1, bitDepth = 8, endianess = 0, bitsPerSample = 8;
  bool isRGB = false, signedOrUnsigned = false;
};

class Engine {
 public:
  virtual ~Engine() {}
  void setCodingPasses(int cp) { _codingPasses = cp; }
  void setWaveletLevels(int wl) { _waveletLevels = wl; }
  void setNumberOfFrames(int f) { _numberOfFrames = f; }
  void setOutputFile(const std::string& o) { _outputFile = o; }
  void setWType(bool lossy) { _lossy = lossy; }
  void setCBWidth(int w) { _cbWidth = w; }
  void setCBHeight(int h) { _cbHeight = h; }
  void setQSizes(float qs) { _qs = qs; }
  void setFrameStructure(Image* img) { _img = img; }
  void setLUTPath(const std::string& p) { _lutPath = p; }
  void setNumberOfStreams(int n) { _streams = n; }
  void setKFactor(float k) { _k = k; }
  virtual void engineManager(int cType) = 0;  // 0 = IMAGE, 1 = VIDEO; errors print and exit like the reference

 protected:
  pcs_ctx* makeContext(int maxBatch, unsigned flags);
  double _tLut = 0, _tCuda = 0, _tCreate = 0;  // start-up breakdown of makeContext (seconds)
  [[noreturn]] void die(const std::string& msg, int code = EXIT_FAILURE_CODE);
  static constexpr int EXIT_FAILURE_CODE = 1;
  int _codingPasses = 2, _waveletLevels = 5, _numberOfFrames = 0, _cbWidth = 64, _cbHeight = 18, _streams = 2;
  bool _lossy = false;
  float _qs = 1.0f, _k = 0.0f;
  std::string _outputFile, _lutPath;
  Image* _img = nullptr;
};

// Engines/CodingEngine.cu:593-1069 (grey paths)
class CodingEngine : public Engine {
 public:
  void engineManager(int cType) override;
};

// Engines/DecodingEngine.cu:140-235,765-1142 (grey paths)
class DecodingEngine : public Engine {
 public:
  void engineManager(int cType) override;
};

}  // namespace picsong
