// long-horizon static-shape instance of the fused kernel (config C4: Np = 20, Nc = 5, 80 constraint rows)
#define MEK_LUT_OFF   // tabulated cell functions are served by the runtime-shape instances (mek_cellfn.cuh)
#include "mek_kernels.cuh"
#include "mek_entry.h"
const void* mek_fused_wide_kernel(int threads) {
  if (threads == 256) return (const void*)k_run_population<20, 5, 15, 256, 1>;
  if (threads == 64) return (const void*)k_run_population<20, 5, 15, 64, 1>;
  return nullptr;
}
