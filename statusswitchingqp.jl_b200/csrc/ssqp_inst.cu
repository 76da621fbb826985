// ssqp_inst.cu — one instantiation of the solve kernel per translation unit, so that the variants compile in
// parallel: -DSSQP_NT=128|256|512 (threads per CTA), optionally -DSSQP_ONLY_VW4 (problem sizes with N % 4 == 0 and
// (M+J) % 4 == 0: only the 256-bit streaming loads are compiled in, which keeps the kernel's code much smaller).
#include "ssqp_kernel.cuh"
#ifndef SSQP_NT
#error "compile with -DSSQP_NT=<128|256|512>"
#endif
#define SSQP_CAT3(a, b, c) a##b##c
#define SSQP_CAT(a, b, c) SSQP_CAT3(a, b, c)
#ifdef SSQP_ONLY_VW4
#define SSQP_FLAVOUR _vw4
#else
#define SSQP_FLAVOUR _any
#endif
typedef void (*ssqp_kernel_fn)(const ssqp::KParams);
ssqp_kernel_fn SSQP_CAT(ssqp_kernel_ptr_, SSQP_NT, SSQP_FLAVOUR)() { return ssqp::ssqp_solve_kernel<SSQP_NT>; }
