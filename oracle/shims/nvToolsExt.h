/* Build shim for oracle/_ref (test infrastructure, not product code).
 * CUDA 12.9 dropped the legacy NVTX2 header the reference includes as
 * "nvToolsExt.h"; NVTX3 is header-only and API compatible. */
#pragma once
#include <nvtx3/nvToolsExt.h>
