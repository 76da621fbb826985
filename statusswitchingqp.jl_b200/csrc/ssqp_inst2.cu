// ssqp_inst2.cu — one instantiation of the v2 solve kernel (shared-memory resident inverse) per translation
// unit (-DSSQP_NT=256|512|1024 = threads per CTA), so that the variants compile in parallel.
#include "ssqp_kernel2.cuh"
#ifndef SSQP_NT
#error "compile with -DSSQP_NT=<256|512|1024>"
#endif
#define SSQP_CAT2(a, b) a##b
#define SSQP_CAT(a, b) SSQP_CAT2(a, b)
typedef void (*ssqp2_kernel_fn)(const ssqp2::KParams);
ssqp2_kernel_fn SSQP_CAT(ssqp2_kernel_ptr_, SSQP_NT)() { return ssqp2::ssqp_solve_kernel<SSQP_NT>; }
