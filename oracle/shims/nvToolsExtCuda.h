/* Build shim for oracle/_ref: legacy name -> NVTX3 header. */
#pragma once
#include <nvtx3/nvToolsExtCuda.h>
