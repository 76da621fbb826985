// ssqp_inst.cu — one instantiation of the solve kernel per translation unit (-DSSQP_CMAX=4|8|12|20|40),
// so that the five variants compile in parallel.  CMAX = ceil((N+M+J)/32): per-lane column accumulators
// of the packed symmetric GEMV / rank-1 update.
#include "ssqp_kernel.cuh"
#ifndef SSQP_CMAX
#error "compile with -DSSQP_CMAX=<4|8|12|20|40>"
#endif
#define SSQP_CAT2(a, b) a##b
#define SSQP_CAT(a, b) SSQP_CAT2(a, b)
typedef void (*ssqp_kernel_fn)(const ssqp::KParams);
ssqp_kernel_fn SSQP_CAT(ssqp_kernel_ptr_, SSQP_CMAX)() { return ssqp::ssqp_solve_kernel<SSQP_CMAX>; }
