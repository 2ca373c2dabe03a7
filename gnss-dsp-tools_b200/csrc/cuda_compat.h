// Build-mode glue. The product is compiled by nvcc for sm_100a only. When GNSSACQ_EMU_BUILD
// is defined (tests/cuda_emu/build_emu.sh, never by __graft_entry__.build()), the same
// sources are compiled by g++ against a host shim so tests can check kernel index logic
// without a GPU; that build is not shipped and is not reachable from the product loader.
#pragma once

#ifdef GNSSACQ_EMU_BUILD
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#define GNSSACQ_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define GNSSACQ_DYN_SMEM(type, name)                       \
  extern __shared__ __align__(128) unsigned char gnssacq_dyn_smem_[]; \
  type* name = reinterpret_cast<type*>(gnssacq_dyn_smem_)
#endif
