// k1_cut_small.cu — the K1 kernels of the small size class (contracted graphs of <= 31 nodes: bit-set searches) as a
// translation unit of their own: k1_cut.cu once more, with the compiler's own loop unrolling.  Measured on a B200
// (profiles/r02_k1_warm.md): the small class is faster unrolled (C2 1.11 against 1.18 ms, one candidate at a time 2.00
// against 2.56 ms), the larger graphs' kernels are faster with no loop unrolled (C5 86 against 103 ms: instruction cache).
#define SGUFP_K1_SMALL_TU
#ifndef SGUFP_K1_SMALL_AB
#define SGUFP_K1_UNROLL
#endif
#include "k1_cut.cu"
