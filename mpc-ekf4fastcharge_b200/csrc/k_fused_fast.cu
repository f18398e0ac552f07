// static-shape short-horizon instances of the fused kernel (runMPC's own configuration and the [1 1 0] variant)
#define MEK_LUT_OFF   // tabulated cell functions are served by the runtime-shape instances (mek_cellfn.cuh)
#include "mek_kernels.cuh"
#include "mek_entry.h"
const void* mek_fused_fast_kernel(int cons, int threads) {
  if (cons == 15 && threads == 256) return (const void*)k_run_population<5, 2, 15, 256, 1>;
  if (cons == 15 && threads == 384) return (const void*)k_run_population<5, 2, 15, 384, 1>;
  if (cons == 11 && threads == 256) return (const void*)k_run_population<5, 2, 11, 256, 1>;
  if (cons == 11 && threads == 384) return (const void*)k_run_population<5, 2, 11, 384, 1>;
  return nullptr;
}
