// cuda_compat.h -- one include for the CUDA runtime.  The product build (nvcc,
// sm_100a) takes the first branch; -DBSK_EMU selects the development-only host
// emulator (emu/cuda_emu.h) used to debug kernel logic without a GPU.
#pragma once
#ifdef BSK_EMU
#include "emu/cuda_emu.h"
#else
#include <cuda_runtime.h>
#define BSK_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BSK_LAUNCH_FLAT(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BSK_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char name##_raw_smem[]; \
  type *name = reinterpret_cast<type *>(name##_raw_smem)
#endif
