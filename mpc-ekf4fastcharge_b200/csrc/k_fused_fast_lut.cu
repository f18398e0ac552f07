// static-shape short-horizon instances of the fused kernel for a bank with TABULATED cell functions (mek_bank_set_lut):
// the same kernels as k_fused_fast.cu with the table look-up compiled in
#include "mek_kernels.cuh"
#include "mek_entry.h"
const void* mek_fused_fast_lut_kernel(int cons, int threads) {
  if (cons == 15 && threads == 256) return (const void*)k_run_population<5, 2, 15, 256, 1>;
  if (cons == 11 && threads == 256) return (const void*)k_run_population<5, 2, 11, 256, 1>;
  return nullptr;
}
