// runtime-shape instances of the fused kernel (any constraint set, horizons up to Np 32 / Nc 8)
#include "mek_kernels.cuh"
#include "mek_entry.h"
const void* mek_fused_generic_kernel(int inst) {
  if (inst == 0) return (const void*)k_run_population<5, 2, -1, 128, 2>;
  if (inst == 1) return (const void*)k_run_population<20, 5, -1, 128, 2>;
  if (inst == 2) return (const void*)k_run_population<32, 8, -1, 128, 2>;
  return nullptr;
}
