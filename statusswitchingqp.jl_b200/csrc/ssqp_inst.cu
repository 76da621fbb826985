// ssqp_inst.cu — one instantiation of the solve kernel per translation
// unit (-DSSQP_NT=256|512 = threads per CTA), so that the variants compile in parallel.
#include "ssqp_kernel.cuh"
#ifndef SSQP_NT
#error "compile with -DSSQP_NT=<256|512>"
#endif
#define SSQP_CAT2(a, b) a##b
#define SSQP_CAT(a, b) SSQP_CAT2(a, b)
typedef void (*ssqp_kernel_fn)(const ssqp::KParams);
ssqp_kernel_fn SSQP_CAT(ssqp_kernel_ptr_, SSQP_NT)() { return ssqp::ssqp_solve_kernel<SSQP_NT>; }
