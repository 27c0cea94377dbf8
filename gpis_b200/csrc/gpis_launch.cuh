// Kernel-launch and dynamic-shared-memory macros shared by the headers that tests/emu can also compile
// for the CPU emulation of the execution model (-DGPIS_EMULATE: test infrastructure only, the shipped
// library is compiled by nvcc without it).
#pragma once
#ifdef GPIS_EMULATE
#include "cuda_emu.h"
#define GPIS_LAUNCH(kern, grid, block, stream, ...) emu::launch_kernel(grid, block, 0, kern, __VA_ARGS__)
#define GPIS_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) emu::launch_kernel(grid, block, smem, kern, __VA_ARGS__)
#define GPIS_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem)
#define GPIS_DYN_SMEM_F64(name) double* name = reinterpret_cast<double*>(emu::dyn_smem)
#define GPIS_DYN_SMEM_F64A(name) double* name = reinterpret_cast<double*>(emu::dyn_smem)
#else
#define GPIS_LAUNCH(kern, grid, block, stream, ...) kern<<<grid, block, 0, stream>>>(__VA_ARGS__)
#define GPIS_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define GPIS_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; \
  type* name = reinterpret_cast<type*>(name##_raw)
#define GPIS_DYN_SMEM_F64(name) extern __shared__ double name[]
#define GPIS_DYN_SMEM_F64A(name) extern __shared__ __align__(16) double name[]
#endif
