// nis_sizes.h -- the 1-D transform lengths the kernels are instantiated for.
// Column pass (transform along image rows, the halved dimension: H, rotation_divisor):
//   X(N, F0,F1,F2, I0,I1,I2, T)   forward radices (the paired, separating stage is F2), inverse radices (paired I0), threads
// Row pass (transform along image columns: W, rotation_channel), first radix fixed at 16:
//   X(N, R1, R2, L, T, LR)        N = 16*R1*R2, L lines per CTA, threads; LR = lines per CTA of the fused fwd->mid->inv kernel
//                                 (two line buffers, so fewer lines keep more CTAs resident)
// Image widths must also be multiples of 32 (one column-pass CTA owns 32 real columns).
#pragma once
#include "nis_fft.cuh"
// T is given for 16 complex lanes per CTA and scaled with NIS_COL_LANES (nis_fft.cuh) so the butterflies per thread stay put
#ifndef NIS_COL_TBASE
#define NIS_COL_TBASE 256     // 128-thread column CTAs at 8 lanes: measured +3.5% over 256 (more, smaller CTAs hide barrier stalls)
#endif
#define NIS_CT(t) ((t) * NIS_COL_LANES / 16)
// (A first radix of 8 would let the middle-stage twiddles hoist for every plan -- measured 7 % slower overall, so only the
// inverse 720 plan, whose first radix is 8 anyway, hoists.)
#define NIS_COL_PLANS(X)                      \
  X(480, 10, 8, 6, 6, 8, 10, NIS_CT(NIS_COL_TBASE))     \
  X(720, 10, 9, 8, 8, 9, 10, NIS_CT(NIS_COL_TBASE))     \
  X(960, 10, 12, 8, 8, 12, 10, NIS_CT(NIS_COL_TBASE))   \
  X(1200, 10, 12, 10, 10, 12, 10, NIS_CT(NIS_COL_TBASE)) \
  X(96, 4, 6, 4, 4, 6, 4, NIS_CT(128))        \
  X(80, 5, 4, 4, 4, 4, 5, NIS_CT(128))        \
  X(64, 4, 4, 4, 4, 4, 4, NIS_CT(128))

#ifndef NIS_ROW_L
#define NIS_ROW_L 4             // 4 lines x 128 threads per row CTA (2 lines in the fused fwd->mid->inv kernel): measured +3.6% over 8 x 256
#endif
#ifndef NIS_ROW_T
#define NIS_ROW_T 128
#endif
#ifndef NIS_ROW_LR
#define NIS_ROW_LR 2
#endif
// the 480-point row pass (polar grid) may use its own geometry: with T a multiple of NS2 = 96 its last-stage twiddles hoist too
#ifndef NIS_ROW480_L
#define NIS_ROW480_L NIS_ROW_L
#define NIS_ROW480_T NIS_ROW_T
#define NIS_ROW480_LR NIS_ROW_LR
#endif
#define NIS_ROW_PLANS(X)      \
  X(640, 8, 5, NIS_ROW_L, NIS_ROW_T, NIS_ROW_LR)     \
  X(480, 6, 5, NIS_ROW480_L, NIS_ROW480_T, NIS_ROW480_LR)     \
  X(1280, 8, 10, 4, 256, 2)   \
  X(1600, 10, 10, 4, 256, 2)  \
  X(128, 8, 1, 8, 128, 8)     \
  X(96, 6, 1, 8, 128, 8)      \
  X(64, 4, 1, 8, 128, 8)
