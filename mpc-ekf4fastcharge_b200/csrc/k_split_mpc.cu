// phase-split form, MPC launch (mek_split.cuh)
#include "mek_split.cuh"
#include "mek_entry.h"
const void* mek_split_mpc_kernel(int cons, int threads) {
  if (cons == 15 && threads == 256) return (const void*)k_split_mpc<5, 2, 15, 256, 1>;
  if (cons == 15 && threads == 384) return (const void*)k_split_mpc<5, 2, 15, 384, 1>;
  if (cons == 15 && threads == 512) return (const void*)k_split_mpc<5, 2, 15, 512, 1>;
  if (cons == 11 && threads == 256) return (const void*)k_split_mpc<5, 2, 11, 256, 1>;
  if (cons == 11 && threads == 384) return (const void*)k_split_mpc<5, 2, 11, 384, 1>;
  if (cons == 11 && threads == 512) return (const void*)k_split_mpc<5, 2, 11, 512, 1>;
  return nullptr;
}
