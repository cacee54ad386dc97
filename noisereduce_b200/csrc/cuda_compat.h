// cuda_compat.h -- one include for every translation unit of libb200gate.
//
// Product build: nvcc, sm_100a, the real CUDA runtime.
// tests/cusim build (B200_CUSIM defined by tests/cusim/build_cusim.py, g++ only): the same sources
// run on the CPU fiber simulator so kernel logic can be debugged in the GPU-less build container.
// The simulator build is test infrastructure and is never loaded by the Python package.
#pragma once

#ifdef B200_CUSIM_BUILD
#include "cusim.h"
#else
#include <cuda_runtime.h>
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define B200_DYN_SMEM(type, name) \
    extern __shared__ __align__(16) unsigned char b200_dyn_smem_raw[]; \
    type* name = reinterpret_cast<type*>(b200_dyn_smem_raw)
#endif

#include <stdint.h>
