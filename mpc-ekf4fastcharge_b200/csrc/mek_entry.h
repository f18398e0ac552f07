// Entry points of the fused kernels, one translation unit per family so that they build in parallel.
// Each function returns the kernel's host-side function pointer (what cudaLaunchKernel, cudaFuncSetAttribute and the
// occupancy calculator take) or nullptr when that shape is not built.
#pragma once

// k_run_population<Np, Nc, CONS, THREADS, MINB> (mek_kernels.cuh)
const void* mek_fused_fast_kernel(int cons, int threads);      // <5,2,cons,threads,1>: cons 15 | 11, threads 256 | 384
const void* mek_fused_fast_lut_kernel(int cons, int threads);  // the same with tabulated cell functions: threads 256
const void* mek_fused_wide_kernel(int threads);                // <20,5,15,threads,1>: threads 256 | 64
const void* mek_fused_generic_kernel(int inst);                // <5,2,-1,128,2>, <20,5,-1,128,2>, <32,8,-1,128,2>
