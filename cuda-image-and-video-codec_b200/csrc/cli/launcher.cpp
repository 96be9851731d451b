// launcher.cpp -- PICSONG-compatible command line (CUDA_ImCod/Launcher.cu:36-233 and
// IO/CommandLineParser.cpp:4-49 of the reference): same flags, same defaults, same validation and
// echo lines, same file formats.  `-cbLength` (the name the reference's help text uses) is accepted as
// an alias of `-cbHeight` (the name its parser reads).
#include <algorithm>
#include <chrono>
#include <exception>
#include <iostream>
#include <string>
#include <vector>

#include "engines.hpp"

namespace {

class InputParser {  // IO/CommandLineParser.cpp:4-26
 public:
  InputParser(int argc, char** argv) { for (int i = 1; i < argc; ++i) tokens.emplace_back(argv[i]); }
  std::string get(const std::string& opt) const {
    auto it = std::find(tokens.begin(), tokens.end(), opt);
    if (it != tokens.end() && ++it != tokens.end()) return *it;
    return "";
  }
  bool has(const std::string& opt) const { return std::find(tokens.begin(), tokens.end(), opt) != tokens.end(); }
 private:
  std::vector<std::string> tokens;
};

void printHelp() {  // IO/CommandLineParser.cpp:28-49
  std::cout << "PICSONG (B200-native hot path) parameters:\n"
               " -cd [0|1]            0 = code, 1 = decode\n"
               " -i / -o <file>       input / output file (encode input: header-less 8-bit RAW, W*H bytes per frame)\n"
               " -xSize / -ySize      frame size\n"
               " -wl [1..10]          wavelet levels (default 5)\n"
               " -cp [2|3]            coding passes (2 only in this build)\n"
               " -type [0|1]          0 = lossless 5/3, 1 = lossy 9/7\n"
               " -qs [0..1]           quantisation size (lossy)\n"
               " -cbWidth / -cbLength DWT block size written to the header (64 / 18..20); -cbHeight is the same flag\n"
               " -video [0|1] -frames N   video mode (writes <out>_SIZE)\n"
               " -LUTFolder <dir/>    probability tables\n"
               " -k [0..65.535]       complexity scalability (0 = off; > 0 needs the LUT folder's bit-plane files *_0 .. *_15)\n"
               " -numberOfStreams N   accepted for compatibility (frames are batched instead)\n"
               " -bps -endianess -signedOrUnsigned -components -isRGB -avoidSizeCheck   as in the reference\n";
}

}  // namespace

static int run(int argc, char** argv);

int main(int argc, char** argv) {
  try {
    return run(argc, argv);
  } catch (const std::exception&) {  // std::stoi / std::stof on a malformed flag value
    std::cout << "Incorrect parameters. Please choose valid values." << std::endl;  // Launcher.cu:132-136
    return 255;
  }
}

static int run(int argc, char** argv) {
  auto start = std::chrono::steady_clock::now();
  InputParser ip(argc, argv);
  if (ip.has("-h")) { printHelp(); return 0; }
  std::string inputFile, outputFile, lutFolder;
  int cd = 2, xSize = 0, ySize = 0, cbWidth = 64, cbHeight = 18, wLevels = 5, cp = 2, endianess = 0, bps = 8, signedOrUnsigned = 0,
      isVideo = 0, frames = 0, avoidSizeCheck = 0, components = 1, numOfStreams = 2;
  bool isRGB = false, type = false;
  float qs = 1, k = 0;
  auto geti = [&](const char* flag, int& v, const char* echo = nullptr) {
    if (ip.has(flag)) { v = std::stoi(ip.get(flag)); std::cout << "User entered " << (echo ? echo : flag) << " command " << v << std::endl; }
  };
  auto gets = [&](const char* flag, std::string& v) {
    if (ip.has(flag)) { v = ip.get(flag); std::cout << "User entered " << flag << " command " << v << std::endl; }
  };
  geti("-cd", cd);
  if (cd == 0) {
    geti("-wl", wLevels); geti("-cp", cp);
    if (ip.has("-type")) { type = std::stoi(ip.get("-type")) != 0; std::cout << "User entered -t command " << type << std::endl; }
    if (ip.has("-qs")) { qs = std::stof(ip.get("-qs")); std::cout << "User entered -qs command " << qs << std::endl; }
    gets("-i", inputFile); gets("-o", outputFile);
    geti("-cbWidth", cbWidth); geti("-cbHeight", cbHeight); geti("-cbLength", cbHeight, "-cbHeight");
    geti("-xSize", xSize); geti("-ySize", ySize); geti("-endianess", endianess); geti("-bps", bps);
    geti("-signedOrUnsigned", signedOrUnsigned); gets("-LUTFolder", lutFolder); geti("-video", isVideo); geti("-frames", frames);
    geti("-avoidSizeCheck", avoidSizeCheck); geti("-components", components);
    if (ip.has("-isRGB")) { isRGB = std::stoi(ip.get("-isRGB")) != 0; std::cout << "User entered -isRGB command " << isRGB << std::endl; }
    geti("-numberOfStreams", numOfStreams);
    if (ip.has("-k")) { k = std::stof(ip.get("-k")); std::cout << "User entered -k command " << k << std::endl; }
    if (qs < 0 || qs > 1 || wLevels < 1 || xSize <= 0 || ySize <= 0 || wLevels > 10 || inputFile.empty() || outputFile.empty() ||
        cbWidth % 64 != 0 || cbHeight > 20 || cbHeight < 18 || cp < 2 || cp > 3 || k < 0 || k > 65.535f) {
      std::cout << "Incorrect parameters. Please choose valid values." << std::endl;  // Launcher.cu:132-136
      return 255;  // exit(-1)
    }
  } else {
    gets("-i", inputFile); gets("-o", outputFile); gets("-LUTFolder", lutFolder); geti("-numberOfStreams", numOfStreams); geti("-video", isVideo);
  }
  picsong::Image img;
  if (cd == 0) {
    img.bitDepth = bps; img.bitsPerSample = bps; img.components = components; img.endianess = endianess; img.height = ySize;
    img.isRGB = isRGB; img.name = inputFile; img.signedOrUnsigned = signedOrUnsigned != 0; img.width = xSize;
    picsong::CodingEngine e;
    e.setNumberOfStreams(numOfStreams); e.setCBHeight(cbHeight); e.setCBWidth(cbWidth); e.setCodingPasses(cp); e.setFrameStructure(&img);
    e.setLUTPath(lutFolder); e.setNumberOfFrames(frames); e.setOutputFile(outputFile); e.setQSizes(qs); e.setWaveletLevels(wLevels);
    e.setWType(type); e.setKFactor(k);
    e.engineManager(isVideo);
  } else if (cd == 1) {
    img.name = inputFile;
    picsong::DecodingEngine e;
    e.setNumberOfStreams(numOfStreams); e.setFrameStructure(&img); e.setOutputFile(outputFile); e.setLUTPath(lutFolder);
    e.engineManager(isVideo);
  } else {
    printHelp();
    return 0;
  }
  std::cout << "The time spent with the app is: " << std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count() << std::endl;
  return 0;
}
