// nis_sizes.h -- the 1-D transform lengths the kernels are instantiated for.
// Column pass (transform along image rows, the halved dimension: H, rotation_divisor), in place on a digit-addressed tile:
//   X(N, A, B, C, T)   N = A*B*C in stage order of the decimation-in-frequency passes (the fused kernel's forward half runs the
//                      same digits in time order C, B, A); T threads per CTA.  B should be odd (the paired last stage of the
//                      forward pass then alternates bank halves), C even and a divisor of T/8 (stage-B twiddles hoist).
// Row pass (transform along image columns: W, rotation_channel):
//   X(N, R0, R1, R2, L, T, LR)    N = R0*R1*R2 in stage order, L lines per CTA, threads; LR = lines per CTA of the fused
//                                 fwd->mid->inv kernel (two line buffers, so fewer lines keep more CTAs resident); R0 + 1 odd (bank
//                                 conflicts); R2 = 1 selects the two-stage code path
// Image widths must be multiples of 32 and rotation_channel of 16 (one column-pass CTA owns 16 real columns; nis_create checks).
#pragma once
#include "nis_fft.cuh"
#ifndef NIS_COL_T
#define NIS_COL_T 256
#endif
// the two production lengths can be re-planned at build time for A/B runs (-DNIS_P720_A=12 -DNIS_P720_B=5 -DNIS_P720_C=12)
#ifndef NIS_P480_A
#define NIS_P480_A 12
#define NIS_P480_B 5
#define NIS_P480_C 8
#endif
// 480 = 12*5*8 has 21 / 96 / 60 butterfly tasks per stage: 24 thread groups (192 threads, 4 CTAs per SM) fit them better than 32
// (measured 73.9 k -> 76.0 k solves/s, col_fwd_rotate 1.71 -> 1.52 ms per 1000 frames, profiles/ab_r02.md)
#ifndef NIS_P480_T
#define NIS_P480_T 192
#endif
#ifndef NIS_P720_A
#define NIS_P720_A 12       // 720 = 12*5*12: 31 / 144 / 60 butterfly tasks per stage on 32 groups (10*9*8 leaves a 5-of-32 tail round in the
#define NIS_P720_B 5        // paired stages; measured 64.7 k vs 59.3 k solves/s, profiles/ab_r02.md)
#define NIS_P720_C 12
#endif
#define NIS_COL_PLANS(X)          \
  X(480, NIS_P480_A, NIS_P480_B, NIS_P480_C, NIS_P480_T)     \
  X(720, NIS_P720_A, NIS_P720_B, NIS_P720_C, NIS_COL_T)     \
  X(960, 8, 15, 8, NIS_COL_T)     \
  X(1200, 10, 15, 8, NIS_COL_T)   \
  X(96, 4, 3, 8, 128)             \
  X(80, 2, 5, 8, 128)             \
  X(64, 8, 1, 8, 128)

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
// Plan A of the two production row lengths: three stages, radix 16 first (re-plannable at build time for A/B runs:
// -DNIS_R640_0=8 -DNIS_R640_1=10 -DNIS_R640_2=8).  Radix 16 FIRST is deliberate: its runs of 40 / 30 elements touch 2.0x / 3.4x
// the 128-byte lines the data occupies, but the plans with line-aligned runs (640 = 8*10*8 or 8*8*10, 480 = 10*8*6 or 6*8*10)
// measured 59.3 k - 67.0 k solves/s against 68.6 k (profiles/ab_r02.md): sixteen independent loads per thread in flight matter
// more than L1 wavefronts -- the pass is latency bound.
#ifndef NIS_R640_0
#define NIS_R640_0 16
#define NIS_R640_1 8
#define NIS_R640_2 5
#endif
#ifndef NIS_R480_0
#define NIS_R480_0 16
#define NIS_R480_1 6
#define NIS_R480_2 5
#endif
#define NIS_ROW_PLANS(X)      \
  X(640, NIS_R640_0, NIS_R640_1, NIS_R640_2, NIS_ROW_L, NIS_ROW_T, NIS_ROW_LR)     \
  X(480, NIS_R480_0, NIS_R480_1, NIS_R480_2, NIS_ROW480_L, NIS_ROW480_T, NIS_ROW480_LR)     \
  X(1280, 16, 8, 10, 4, 256, 2)   \
  X(1600, 16, 10, 10, 4, 256, 2)  \
  X(128, 16, 8, 1, 8, 128, 8)     \
  X(96, 16, 6, 1, 8, 128, 8)      \
  X(64, 16, 4, 1, 8, 128, 8)

// Plan B: TWO stages (radix 32, then 20 / 15): one shared-memory exchange and one barrier per pass, 32 loads per thread in flight,
// 80-96 registers.  Measured per kernel family against plan A (ms per 1000 frames, profiles/ab_r02.md): fused fwd->mid->inv kernels
// 2.17 -> 1.72 (filter), 1.07 -> 0.80 (store-abs), 0.92 -> 0.84 (mul-conj); row_fwd_h 1.21 -> 1.10; row_fwd 0.60 -> 0.59; but the
// inverse pass with the two-operand product prologue 2.02 -> 2.34 (64 loads per thread, 96 registers).  So each launcher picks its
// plan (NIS_ROWB_* below); sizes without a plan B entry use plan A everywhere.  With the big radix second (20 x 32, 15 x 32) every
// family is slower than with three stages: it is the 32 independent loads of the first stage that pay.
#ifndef NIS_ROWB_L
#define NIS_ROWB_L 3            // 3 lines x 96 threads: stage 1 (32 butterflies per line) fills the CTA exactly
#define NIS_ROWB_T 96
#define NIS_ROWB_LR 3
#endif
#ifndef NIS_RB640_0
#define NIS_RB640_0 32
#define NIS_RB640_1 20
#endif
#ifndef NIS_RB480_0
#define NIS_RB480_0 32
#define NIS_RB480_1 15
#endif
#define NIS_ROW_PLANS_B(X)    \
  X(640, NIS_RB640_0, NIS_RB640_1, 1, NIS_ROWB_L, NIS_ROWB_T, NIS_ROWB_LR)   \
  X(480, NIS_RB480_0, NIS_RB480_1, 1, NIS_ROWB_L, NIS_ROWB_T, NIS_ROWB_LR)
#ifndef NIS_ROWB_FWD
#define NIS_ROWB_FWD 1          // row_fwd (P)
#endif
#ifndef NIS_ROWB_FWDH
#define NIS_ROWB_FWDH 1         // row_fwd_h (keyframe factor H)
#endif
#ifndef NIS_ROWB_INVMC
#define NIS_ROWB_INVMC 0        // row_inv_mulconj
#endif
#ifndef NIS_ROWB_INVMC_AUTO
#define NIS_ROWB_INVMC_AUTO 0    // row_inv_mulconj in its auto form |Z|^2 (one operand)
#endif
#ifndef NIS_ROWB_RR
#define NIS_ROWB_RR 1           // the fused fwd->mid->inv kernels (store-abs, store-square, mul-conj, filter)
#endif
