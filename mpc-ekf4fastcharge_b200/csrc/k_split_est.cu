// phase-split form, estimator launch and block-boundary launch (mek_split.cuh)
#include "mek_split.cuh"
#include "mek_entry.h"
const void* mek_split_est_kernel(int threads) {
  if (threads == 256) return (const void*)k_split_est<256, 1>;
  if (threads == 384) return (const void*)k_split_est<384, 1>;
  if (threads == 512) return (const void*)k_split_est<512, 1>;
  return nullptr;
}
const void* mek_split_block_kernel() { return (const void*)k_split_block<SPLIT_BLK_THREADS>; }
