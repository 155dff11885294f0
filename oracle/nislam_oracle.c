/*
 * nislam_oracle.c -- dependency-free CPU restatement of NI-SLAM's tracking / loop-closure hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ni_slam_b200/ (the product) may link, load or call this
 * file.  Users: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY PINNED TO THE REFERENCE'S SOURCE TEXT: the reference (sair-lab/ni-slam @ 819f252) has no tests / golden vectors and its
 * own build needs Eigen, FFTW3 and OpenCV C++ (absent here), but its hot-path sources compile UNMODIFIED against the stand-in
 * headers of oracle/ref_stubs (oracle/Makefile.ref -> oracle/_ref/libnislam_ref.so); tests/test_oracle_ref.py holds this file
 * equal to that library (integers and theta exact, info within 1e-6; observed bit-identical) and to the committed vectors
 * it produced (tests/golden/golden_ref.npz).  What stays unpinned is the third-party arithmetic itself (FFTW's and Eigen's
 * rounding, see ref_stubs/mini_eigen.h for the evaluation orders assumed).  This file follows
 *   src/correlation_flow.cc:37-243, src/utils.cc:110-131,154-175, include/circ_shift.h:238-244,
 *   src/loop_closure.cc:36-73, include/loop_closure.h:15
 * and restates the published algorithms of the third-party calls on the path:
 *   FFTW3f  fftwf_plan_dft_r2c_2d / c2r_2d (correlation_flow.cc:56-61,70-74) -> own mixed-radix f32 FFT
 *   OpenCV 4.2 cv::warpPolar (correlation_flow.cc:234)                       -> orc_polar (1/32-px remap)
 *   OpenCV 4.2 cv::getRotationMatrix2D + cv::warpAffine (utils.cc:158-159)   -> orc_rotate (10+5 bit fixed point)
 *   Eigen maxCoeff(&row,&col) (correlation_flow.cc:175)                      -> first max in column-major order
 * It is cross-checked against oracle/nislam_ref.py (scipy pocketfft + the genuine cv2) by
 * tests/test_oracle_*.py and against the analytic known answers of SURVEY.md Appendix C.
 *
 * Layout at this interface = the reference's: Eigen column-major arrays (rows fastest).  A real R x C
 * array is C lines of R floats; its half spectrum is (R/2+1) x C complex, C lines of R/2+1 (re,im) pairs.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------ */
/* 1-D complex FFT, planar, vectorised across a contiguous batch: data[n][batch]               */
/* ------------------------------------------------------------------------------------------ */
#define ORC_MAX_FACT 24

typedef struct {
  int n, nf;
  int radix[ORC_MAX_FACT];
  float *wr, *wi; /* exp(-2 pi i t / n), t in [0,n) */
} orc_plan;

#define ORC_MAX_PLANS 16
static orc_plan g_plans[ORC_MAX_PLANS];
static int g_nplans = 0;

static const orc_plan *get_plan(int n) {
  const orc_plan *found = NULL;
#pragma omp critical(orc_plan_cache)
  {
    for (int i = 0; i < g_nplans; ++i)
      if (g_plans[i].n == n) found = &g_plans[i];
    if (!found && g_nplans < ORC_MAX_PLANS) {
      orc_plan *p = &g_plans[g_nplans];
      p->n = n;
      p->nf = 0;
      int m = n;
      while (m % 4 == 0) { p->radix[p->nf++] = 4; m /= 4; }
      while (m % 2 == 0) { p->radix[p->nf++] = 2; m /= 2; }
      while (m % 3 == 0) { p->radix[p->nf++] = 3; m /= 3; }
      while (m % 5 == 0) { p->radix[p->nf++] = 5; m /= 5; }
      for (int f = 7; m > 1; f += 2)
        while (m % f == 0) { p->radix[p->nf++] = f; m /= f; }
      p->wr = (float *)malloc(sizeof(float) * n);
      p->wi = (float *)malloc(sizeof(float) * n);
      for (int t = 0; t < n; ++t) {
        double a = -2.0 * M_PI * (double)t / (double)n;
        p->wr[t] = (float)cos(a);
        p->wi[t] = (float)sin(a);
      }
      g_nplans++;
      found = p;
    }
  }
  return found;
}

/* One Stockham stage: radix R, Ns = product of earlier radices.  sign=-1 forward, +1 inverse. */
static inline __attribute__((always_inline)) void stage_generic(const orc_plan *p, int R, int Ns, int batch, int sign, const float *restrict ir, const float *restrict ii,
                          float *restrict or_, float *restrict oi) {
  const int n = p->n, nb = n / R;
  float tr[32], ti[32];
  for (int j = 0; j < nb; ++j) {
    const int k = j % Ns;
    const int j0 = (j / Ns) * Ns * R + k;
    const int tstep = n / (Ns * R);
    for (int b = 0; b < batch; ++b) {
      for (int r = 0; r < R; ++r) {
        const int t = (int)(((long)r * k * tstep) % n);
        const float wr = p->wr[t], wi = sign < 0 ? p->wi[t] : -p->wi[t];
        const float xr = ir[(size_t)(j + r * nb) * batch + b], xi = ii[(size_t)(j + r * nb) * batch + b];
        tr[r] = xr * wr - xi * wi;
        ti[r] = xr * wi + xi * wr;
      }
      for (int q = 0; q < R; ++q) {
        float sr = 0.f, si = 0.f;
        for (int r = 0; r < R; ++r) {
          const int t = (int)(((long)q * r * (n / R)) % n);
          const float wr = p->wr[t], wi = sign < 0 ? p->wi[t] : -p->wi[t];
          sr += tr[r] * wr - ti[r] * wi;
          si += tr[r] * wi + ti[r] * wr;
        }
        or_[(size_t)(j0 + q * Ns) * batch + b] = sr;
        oi[(size_t)(j0 + q * Ns) * batch + b] = si;
      }
    }
  }
}

/* q*k*tstep <= (R-1)(Ns-1) n/(Ns R) < n: no modulo needed */
#define LOAD_TW(q)                                                   \
  const int t##q = (q) * k * tstep;                                   \
  const float w##q##r = p->wr[t##q], w##q##i = sign < 0 ? p->wi[t##q] : -p->wi[t##q];

static inline __attribute__((always_inline)) void stage2(const orc_plan *p, int Ns, int batch, int sign, const float *restrict ir, const float *restrict ii,
                   float *restrict or_, float *restrict oi) {
  const int n = p->n, nb = n / 2, tstep = n / (Ns * 2);
  for (int j = 0; j < nb; ++j) {
    const int k = j % Ns, j0 = (j / Ns) * Ns * 2 + k;
    LOAD_TW(1)
    const float *restrict a_r = ir + (size_t)j * batch, *restrict a_i = ii + (size_t)j * batch;
    const float *restrict b_r = ir + (size_t)(j + nb) * batch, *restrict b_i = ii + (size_t)(j + nb) * batch;
    float *restrict o0r = or_ + (size_t)j0 * batch, *restrict o0i = oi + (size_t)j0 * batch;
    float *restrict o1r = or_ + (size_t)(j0 + Ns) * batch, *restrict o1i = oi + (size_t)(j0 + Ns) * batch;
    for (int b = 0; b < batch; ++b) {
      const float xr = b_r[b] * w1r - b_i[b] * w1i, xi = b_r[b] * w1i + b_i[b] * w1r;
      o0r[b] = a_r[b] + xr; o0i[b] = a_i[b] + xi;
      o1r[b] = a_r[b] - xr; o1i[b] = a_i[b] - xi;
    }
  }
}

static inline __attribute__((always_inline)) void stage3(const orc_plan *p, int Ns, int batch, int sign, const float *restrict ir, const float *restrict ii,
                   float *restrict or_, float *restrict oi) {
  const int n = p->n, nb = n / 3, tstep = n / (Ns * 3);
  const float c = -0.5f, s = (sign < 0 ? -1.f : 1.f) * 0.86602540378443864676f;
  for (int j = 0; j < nb; ++j) {
    const int k = j % Ns, j0 = (j / Ns) * Ns * 3 + k;
    LOAD_TW(1) LOAD_TW(2)
    const float *restrict x0r = ir + (size_t)j * batch, *restrict x0i = ii + (size_t)j * batch;
    const float *restrict x1r = ir + (size_t)(j + nb) * batch, *restrict x1i = ii + (size_t)(j + nb) * batch;
    const float *restrict x2r = ir + (size_t)(j + 2 * nb) * batch, *restrict x2i = ii + (size_t)(j + 2 * nb) * batch;
    float *restrict o0r = or_ + (size_t)j0 * batch, *restrict o0i = oi + (size_t)j0 * batch;
    float *restrict o1r = or_ + (size_t)(j0 + Ns) * batch, *restrict o1i = oi + (size_t)(j0 + Ns) * batch;
    float *restrict o2r = or_ + (size_t)(j0 + 2 * Ns) * batch, *restrict o2i = oi + (size_t)(j0 + 2 * Ns) * batch;
    for (int b = 0; b < batch; ++b) {
      const float ar = x0r[b], ai = x0i[b];
      const float br = x1r[b] * w1r - x1i[b] * w1i, bi = x1r[b] * w1i + x1i[b] * w1r;
      const float cr = x2r[b] * w2r - x2i[b] * w2i, ci = x2r[b] * w2i + x2i[b] * w2r;
      const float sr = br + cr, si = bi + ci, dr = br - cr, di = bi - ci;
      const float mr = ar + c * sr, mi = ai + c * si;
      o0r[b] = ar + sr; o0i[b] = ai + si;
      o1r[b] = mr - s * di; o1i[b] = mi + s * dr;
      o2r[b] = mr + s * di; o2i[b] = mi - s * dr;
    }
  }
}

static inline __attribute__((always_inline)) void stage4(const orc_plan *p, int Ns, int batch, int sign, const float *restrict ir, const float *restrict ii,
                   float *restrict or_, float *restrict oi) {
  const int n = p->n, nb = n / 4, tstep = n / (Ns * 4);
  const float sg = sign < 0 ? 1.f : -1.f; /* forward: multiply by -i */
  for (int j = 0; j < nb; ++j) {
    const int k = j % Ns, j0 = (j / Ns) * Ns * 4 + k;
    LOAD_TW(1) LOAD_TW(2) LOAD_TW(3)
    const float *restrict x0r = ir + (size_t)j * batch, *restrict x0i = ii + (size_t)j * batch;
    const float *restrict x1r = ir + (size_t)(j + nb) * batch, *restrict x1i = ii + (size_t)(j + nb) * batch;
    const float *restrict x2r = ir + (size_t)(j + 2 * nb) * batch, *restrict x2i = ii + (size_t)(j + 2 * nb) * batch;
    const float *restrict x3r = ir + (size_t)(j + 3 * nb) * batch, *restrict x3i = ii + (size_t)(j + 3 * nb) * batch;
    float *restrict o0r = or_ + (size_t)j0 * batch, *restrict o0i = oi + (size_t)j0 * batch;
    float *restrict o1r = or_ + (size_t)(j0 + Ns) * batch, *restrict o1i = oi + (size_t)(j0 + Ns) * batch;
    float *restrict o2r = or_ + (size_t)(j0 + 2 * Ns) * batch, *restrict o2i = oi + (size_t)(j0 + 2 * Ns) * batch;
    float *restrict o3r = or_ + (size_t)(j0 + 3 * Ns) * batch, *restrict o3i = oi + (size_t)(j0 + 3 * Ns) * batch;
    for (int b = 0; b < batch; ++b) {
      const float ar = x0r[b], ai = x0i[b];
      const float br = x1r[b] * w1r - x1i[b] * w1i, bi = x1r[b] * w1i + x1i[b] * w1r;
      const float cr = x2r[b] * w2r - x2i[b] * w2i, ci = x2r[b] * w2i + x2i[b] * w2r;
      const float dr = x3r[b] * w3r - x3i[b] * w3i, di = x3r[b] * w3i + x3i[b] * w3r;
      const float s0r = ar + cr, s0i = ai + ci, s1r = ar - cr, s1i = ai - ci;
      const float s2r = br + dr, s2i = bi + di, s3r = br - dr, s3i = bi - di;
      o0r[b] = s0r + s2r; o0i[b] = s0i + s2i;
      o2r[b] = s0r - s2r; o2i[b] = s0i - s2i;
      /* forward: X1 = s1 - i s3, X3 = s1 + i s3 ; inverse swaps */
      o1r[b] = s1r + sg * s3i; o1i[b] = s1i - sg * s3r;
      o3r[b] = s1r - sg * s3i; o3i[b] = s1i + sg * s3r;
    }
  }
}

static inline __attribute__((always_inline)) void stage5(const orc_plan *p, int Ns, int batch, int sign, const float *restrict ir, const float *restrict ii,
                   float *restrict or_, float *restrict oi) {
  const int n = p->n, nb = n / 5, tstep = n / (Ns * 5);
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
  const float sg = sign < 0 ? -1.f : 1.f;
  const float s1 = sg * 0.95105651629515357212f, s2 = sg * 0.58778525229247312917f;
  for (int j = 0; j < nb; ++j) {
    const int k = j % Ns, j0 = (j / Ns) * Ns * 5 + k;
    LOAD_TW(1) LOAD_TW(2) LOAD_TW(3) LOAD_TW(4)
    const float *restrict x0r = ir + (size_t)j * batch, *restrict x0i = ii + (size_t)j * batch;
    const float *restrict x1r = ir + (size_t)(j + nb) * batch, *restrict x1i = ii + (size_t)(j + nb) * batch;
    const float *restrict x2r = ir + (size_t)(j + 2 * nb) * batch, *restrict x2i = ii + (size_t)(j + 2 * nb) * batch;
    const float *restrict x3r = ir + (size_t)(j + 3 * nb) * batch, *restrict x3i = ii + (size_t)(j + 3 * nb) * batch;
    const float *restrict x4r = ir + (size_t)(j + 4 * nb) * batch, *restrict x4i = ii + (size_t)(j + 4 * nb) * batch;
    float *restrict o0r = or_ + (size_t)j0 * batch, *restrict o0i = oi + (size_t)j0 * batch;
    float *restrict o1r = or_ + (size_t)(j0 + Ns) * batch, *restrict o1i = oi + (size_t)(j0 + Ns) * batch;
    float *restrict o2r = or_ + (size_t)(j0 + 2 * Ns) * batch, *restrict o2i = oi + (size_t)(j0 + 2 * Ns) * batch;
    float *restrict o3r = or_ + (size_t)(j0 + 3 * Ns) * batch, *restrict o3i = oi + (size_t)(j0 + 3 * Ns) * batch;
    float *restrict o4r = or_ + (size_t)(j0 + 4 * Ns) * batch, *restrict o4i = oi + (size_t)(j0 + 4 * Ns) * batch;
    for (int b = 0; b < batch; ++b) {
      const float ar = x0r[b], ai = x0i[b];
      const float br = x1r[b] * w1r - x1i[b] * w1i, bi = x1r[b] * w1i + x1i[b] * w1r;
      const float cr = x2r[b] * w2r - x2i[b] * w2i, ci = x2r[b] * w2i + x2i[b] * w2r;
      const float dr = x3r[b] * w3r - x3i[b] * w3i, di = x3r[b] * w3i + x3i[b] * w3r;
      const float er = x4r[b] * w4r - x4i[b] * w4i, ei = x4r[b] * w4i + x4i[b] * w4r;
      const float p1r = br + er, p1i = bi + ei, m1r = br - er, m1i = bi - ei;
      const float p2r = cr + dr, p2i = ci + di, m2r = cr - dr, m2i = ci - di;
      o0r[b] = ar + p1r + p2r; o0i[b] = ai + p1i + p2i;
      const float t1r = ar + c1 * p1r + c2 * p2r, t1i = ai + c1 * p1i + c2 * p2i;
      const float t2r = ar + c2 * p1r + c1 * p2r, t2i = ai + c2 * p1i + c1 * p2i;
      /* u = i*(s1*m1 + s2*m2), v = i*(s2*m1 - s1*m2) */
      const float u_r = -(s1 * m1i + s2 * m2i), u_i = (s1 * m1r + s2 * m2r);
      const float v_r = -(s2 * m1i - s1 * m2i), v_i = (s2 * m1r - s1 * m2r);
      o1r[b] = t1r + u_r; o1i[b] = t1i + u_i;
      o4r[b] = t1r - u_r; o4i[b] = t1i - u_i;
      o2r[b] = t2r + v_r; o2i[b] = t2i + v_i;
      o3r[b] = t2r - v_r; o3i[b] = t2i - v_i;
    }
  }
}

/* All stages of one transform for `batch` interleaved lanes held in a small contiguous block [n][batch], so the whole
 * multi-stage transform stays in L1/L2.  Instantiated with a compile-time lane count (8 = one AVX2 vector) so that the
 * inner lane loops of the always_inline stages become single vector operations. Result in (re,im). */
#ifndef ORC_CHUNK
#define ORC_CHUNK 8
#endif
#define ORC_DEFINE_CHUNK(NAME, WIDTH)                                                                             \
  static void NAME(const orc_plan *p, int batch_rt, int sign, float *re, float *im, float *sre, float *sim) {     \
    const int n = p->n, batch = (WIDTH) > 0 ? (WIDTH) : batch_rt;                                                 \
    float *ar = re, *ai = im, *br = sre, *bi = sim;                                                               \
    int Ns = 1;                                                                                                   \
    for (int f = 0; f < p->nf; ++f) {                                                                             \
      const int R = p->radix[f];                                                                                  \
      switch (R) {                                                                                                \
        case 2: stage2(p, Ns, batch, sign, ar, ai, br, bi); break;                                                \
        case 3: stage3(p, Ns, batch, sign, ar, ai, br, bi); break;                                                \
        case 4: stage4(p, Ns, batch, sign, ar, ai, br, bi); break;                                                \
        case 5: stage5(p, Ns, batch, sign, ar, ai, br, bi); break;                                                \
        default: stage_generic(p, R, Ns, batch, sign, ar, ai, br, bi); break;                                     \
      }                                                                                                           \
      Ns *= R;                                                                                                    \
      float *t;                                                                                                   \
      t = ar; ar = br; br = t;                                                                                    \
      t = ai; ai = bi; bi = t;                                                                                    \
    }                                                                                                             \
    if (ar != re) {                                                                                               \
      memcpy(re, ar, sizeof(float) * (size_t)n * batch);                                                          \
      memcpy(im, ai, sizeof(float) * (size_t)n * batch);                                                          \
    }                                                                                                             \
  }
ORC_DEFINE_CHUNK(fft1d_chunk8, ORC_CHUNK)
ORC_DEFINE_CHUNK(fft1d_chunk_any, 0)

/* data [n][batch] planar, in place; transforms are done ORC_CHUNK lanes at a time in a cache-resident scratch block */
static void fft1d_batch(int n, int batch, int sign, float *re, float *im, float *sre, float *sim) {
  const orc_plan *p = get_plan(n);
  (void)sre; (void)sim;
  float *blk = (float *)aligned_alloc(64, sizeof(float) * 4 * (size_t)n * ORC_CHUNK);
  float *cr = blk, *ci = blk + (size_t)n * ORC_CHUNK, *dr = ci + (size_t)n * ORC_CHUNK, *di = dr + (size_t)n * ORC_CHUNK;
  for (int b0 = 0; b0 < batch; b0 += ORC_CHUNK) {
    const int w = batch - b0 < ORC_CHUNK ? batch - b0 : ORC_CHUNK;
    if (w == ORC_CHUNK) {
      for (int i = 0; i < n; ++i) {
        memcpy(cr + (size_t)i * ORC_CHUNK, re + (size_t)i * batch + b0, sizeof(float) * ORC_CHUNK);
        memcpy(ci + (size_t)i * ORC_CHUNK, im + (size_t)i * batch + b0, sizeof(float) * ORC_CHUNK);
      }
      fft1d_chunk8(p, ORC_CHUNK, sign, cr, ci, dr, di);
      for (int i = 0; i < n; ++i) {
        memcpy(re + (size_t)i * batch + b0, cr + (size_t)i * ORC_CHUNK, sizeof(float) * ORC_CHUNK);
        memcpy(im + (size_t)i * batch + b0, ci + (size_t)i * ORC_CHUNK, sizeof(float) * ORC_CHUNK);
      }
    } else {
      for (int i = 0; i < n; ++i) {
        memcpy(cr + (size_t)i * w, re + (size_t)i * batch + b0, sizeof(float) * w);
        memcpy(ci + (size_t)i * w, im + (size_t)i * batch + b0, sizeof(float) * w);
      }
      fft1d_chunk_any(p, w, sign, cr, ci, dr, di);
      for (int i = 0; i < n; ++i) {
        memcpy(re + (size_t)i * batch + b0, cr + (size_t)i * w, sizeof(float) * w);
        memcpy(im + (size_t)i * batch + b0, ci + (size_t)i * w, sizeof(float) * w);
      }
    }
  }
  free(blk);
}

/* ------------------------------------------------------------------------------------------ */
/* workspace                                                                                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  size_t cap;
  float *a_re, *a_im, *b_re, *b_im, *c_re, *c_im;
} orc_ws;

static void ws_reserve(orc_ws *w, size_t n) {
  if (w->cap >= n) return;
  free(w->a_re); free(w->a_im); free(w->b_re); free(w->b_im); free(w->c_re); free(w->c_im);
  w->a_re = (float *)malloc(sizeof(float) * n); w->a_im = (float *)malloc(sizeof(float) * n);
  w->b_re = (float *)malloc(sizeof(float) * n); w->b_im = (float *)malloc(sizeof(float) * n);
  w->c_re = (float *)malloc(sizeof(float) * n); w->c_im = (float *)malloc(sizeof(float) * n);
  w->cap = n;
}

static __thread orc_ws t_ws;

/* ------------------------------------------------------------------------------------------ */
/* 2-D r2c / c2r in the reference's convention (correlation_flow.cc:53-77)                     */
/* ------------------------------------------------------------------------------------------ */
/* x: R x C column-major real.  xf: (R/2+1) x C column-major complex (interleaved), unnormalised. */
void orc_fft2(const float *x, int R, int C, float *xf) {
  const int half = R / 2 + 1, hb = (C + 1) / 2;
  ws_reserve(&t_ws, (size_t)(R > C ? R : C) * (size_t)((half > hb ? half : hb) + 1));
  float *zr = t_ws.a_re, *zi = t_ws.a_im;
  /* pack line b into re, line b+hb into im, transposed to [r][b] (16x16 tiles keep both sides in cache lines) */
  for (int b0 = 0; b0 < hb; b0 += 16)
    for (int r0 = 0; r0 < R; r0 += 16) {
      const int b1 = b0 + 16 < hb ? b0 + 16 : hb, r1 = r0 + 16 < R ? r0 + 16 : R;
      for (int b = b0; b < b1; ++b) {
        const float *la = x + (size_t)b * R;
        const float *lb = (b + hb < C) ? x + (size_t)(b + hb) * R : NULL;
        for (int r = r0; r < r1; ++r) {
          zr[(size_t)r * hb + b] = la[r];
          zi[(size_t)r * hb + b] = lb ? lb[r] : 0.f;
        }
      }
    }
  fft1d_batch(R, hb, -1, zr, zi, t_ws.b_re, t_ws.b_im);
  /* separate the two real lines, write transposed: S[c][k] planar */
  float *sr = t_ws.b_re, *si = t_ws.b_im;
  for (int k0 = 0; k0 < half; k0 += 16)
  for (int b0 = 0; b0 < hb; b0 += 16)
  for (int k = k0; k < (k0 + 16 < half ? k0 + 16 : half); ++k) {
    const int km = (R - k) % R;
    for (int b = b0; b < (b0 + 16 < hb ? b0 + 16 : hb); ++b) {
      const float ar = zr[(size_t)k * hb + b], ai = zi[(size_t)k * hb + b];
      const float br = zr[(size_t)km * hb + b], bi = zi[(size_t)km * hb + b];
      sr[(size_t)b * half + k] = 0.5f * (ar + br);
      si[(size_t)b * half + k] = 0.5f * (ai - bi);
      if (b + hb < C) {
        sr[(size_t)(b + hb) * half + k] = 0.5f * (ai + bi);
        si[(size_t)(b + hb) * half + k] = 0.5f * (br - ar);
      }
    }
  }
  fft1d_batch(C, half, -1, sr, si, t_ws.a_re, t_ws.a_im);
  for (size_t i = 0; i < (size_t)C * half; ++i) {
    xf[2 * i] = sr[i];
    xf[2 * i + 1] = si[i];
  }
}

/* xf: (R/2+1) x C complex -> x: R x C real, divided by `den` (a true division like Eigen's x/x.size(), correlation_flow.cc:76) */
static void ifft2_scaled(const float *xf, int R, int C, float *x, float den) {
  const int half = R / 2 + 1, hb = (C + 1) / 2;
  ws_reserve(&t_ws, (size_t)(R > C ? R : C) * (size_t)((half > hb ? half : hb) + 1));
  float *sr = t_ws.a_re, *si = t_ws.a_im;
  for (size_t i = 0; i < (size_t)C * half; ++i) {
    sr[i] = xf[2 * i];
    si[i] = xf[2 * i + 1];
  }
  fft1d_batch(C, half, +1, sr, si, t_ws.b_re, t_ws.b_im);
  /* z[k][b] = Xa[k] + i Xb[k], Hermitian-extended along k (c2r semantics: imag of k=0 and k=R/2 ignored) */
  float *zr = t_ws.b_re, *zi = t_ws.b_im;
  for (int b0 = 0; b0 < hb; b0 += 16)
  for (int k0 = 0; k0 < half; k0 += 16)
  for (int b = b0; b < (b0 + 16 < hb ? b0 + 16 : hb); ++b) {
    const float *ar = sr + (size_t)b * half, *ai = si + (size_t)b * half;
    const int has_b = (b + hb < C);
    const float *br = has_b ? sr + (size_t)(b + hb) * half : NULL, *bi = has_b ? si + (size_t)(b + hb) * half : NULL;
    for (int k = k0; k < (k0 + 16 < half ? k0 + 16 : half); ++k) {
      float xar = ar[k], xai = ai[k], xbr = has_b ? br[k] : 0.f, xbi = has_b ? bi[k] : 0.f;
      if (k == 0 || 2 * k == R) { xai = 0.f; xbi = 0.f; }
      zr[(size_t)k * hb + b] = xar - xbi;
      zi[(size_t)k * hb + b] = xai + xbr;
      if (k > 0 && 2 * k < R) {
        /* X[R-k] = conj X[k] */
        zr[(size_t)(R - k) * hb + b] = xar + xbi;
        zi[(size_t)(R - k) * hb + b] = -xai + xbr;
      }
    }
  }
  fft1d_batch(R, hb, +1, zr, zi, t_ws.a_re, t_ws.a_im);
  for (int b0 = 0; b0 < hb; b0 += 16)
    for (int r0 = 0; r0 < R; r0 += 16)
      for (int b = b0; b < (b0 + 16 < hb ? b0 + 16 : hb); ++b) {
        float *la = x + (size_t)b * R;
        float *lb = (b + hb < C) ? x + (size_t)(b + hb) * R : NULL;
        for (int r = r0; r < (r0 + 16 < R ? r0 + 16 : R); ++r) {
          la[r] = zr[(size_t)r * hb + b] / den;
          if (lb) lb[r] = zi[(size_t)r * hb + b] / den;
        }
      }
}
/* divided by R*C (correlation_flow.cc:76) */
void orc_ifft2(const float *xf, int R, int C, float *x) { ifft2_scaled(xf, R, C, x, (float)((size_t)R * C)); }
/* unnormalised, like fftwf c2r (the fftw3.h stand-in of oracle/_ref; the reference's own `x/x.size()` follows) */
void orc_ifft2_raw(const float *xf, int R, int C, float *x) { ifft2_scaled(xf, R, C, x, 1.0f); }

/* ------------------------------------------------------------------------------------------ */
/* L1 helpers                                                                                  */
/* ------------------------------------------------------------------------------------------ */
/* utils.cc:110-118: cv::Mat u8 row-major H x W -> ArrayXXf column-major, /255.0 */
void orc_normalize_u8(const uint8_t *img_rowmajor, int H, int W, float *out_colmajor) {
  for (int c = 0; c < W; ++c)
    for (int r = 0; r < H; ++r) out_colmajor[(size_t)c * H + r] = (float)((double)(float)img_rowmajor[(size_t)r * W + c] / 255.0);
}

/* utils.cc:173-175 */
double orc_normalize_degree(double a) { return a - 360.0 * floor((a + 180.0) / 360.0); }

/* correlation_flow.cc:79-87 */
void orc_remove_zero_component(const float *x, int R, int C, float *y) {
  memcpy(y, x, sizeof(float) * (size_t)R * C);
  for (int c = 0; c < C; ++c) y[(size_t)c * R] = (float)((double)(x[(size_t)c * R + 1] + x[(size_t)c * R + R - 1]) / 2.0);
  for (int r = 0; r < R; ++r) y[r] = (float)((double)(x[(size_t)1 * R + r] + x[(size_t)(C - 1) * R + r]) / 2.0);
}

/* circ_shift.h:238-244 */
void orc_fftshift(const float *x, int R, int C, float *y) {
  const int rs = R / 2, cs = C / 2;
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < R; ++r) y[(size_t)c * R + r] = x[(size_t)((c - cs + C) % C) * R + ((r - rs + R) % R)];
}

static inline int cv_round_f(float v) { return (int)lrintf(v); }  /* round half to even */
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

/* bilinear tap weights exactly as OpenCV's BilinearTab_f: w = {(1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx} */
static inline void bil_w(int fx, int fy, float w[4]) {
  const float x = (float)fx * (1.f / 32.f), y = (float)fy * (1.f / 32.f);
  const float x0 = 1.f - x, y0 = 1.f - y;
  w[0] = y0 * x0; w[1] = y0 * x; w[2] = y * x0; w[3] = y * x;
}

/* correlation_flow.cc:228-236: cv::warpPolar(img, Size(Cp, D), center, maxRadius, LINEAR|FILL_OUTLIERS).
 * Stride-generic core: element (y, x) of src lives at src[y*s_ys + x*s_xs], element (phi, rho) of dst at dst[phi*d_ps + rho*d_rs];
 * the two wrappers below instantiate it for the reference's column-major arrays and for row-major cv::Mat data. */
static inline __attribute__((always_inline)) void warp_polar_core(const float *src, int H, int W, size_t s_ys, size_t s_xs, int D, int Cp,
                                                                  float cx, float cy, double maxRadius, float *dst, size_t d_ps,
                                                                  size_t d_rs) {
  const double Kangle = 2.0 * M_PI / D;
  const double Kmag = maxRadius / Cp;
  for (int phi = 0; phi < D; ++phi) {
    const double KKy = Kangle * phi, cp = cos(KKy), sp = sin(KKy);
    for (int rho = 0; rho < Cp; ++rho) {
      const float rf = (float)(rho * Kmag);
      const float mx = (float)((double)rf * cp + (double)cx);
      const float my = (float)((double)rf * sp + (double)cy);
      const int sx = cv_round_f(mx * 32.f), sy = cv_round_f(my * 32.f);
      const int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
      float w[4];
      bil_w(sx & 31, sy & 31, w);
      float v;
      if ((unsigned)ix < (unsigned)(W - 1) && (unsigned)iy < (unsigned)(H - 1)) {
        const float *S = src + (size_t)ix * s_xs + (size_t)iy * s_ys;
        v = S[0] * w[0] + S[s_xs] * w[1] + S[s_ys] * w[2] + S[s_xs + s_ys] * w[3];
      } else if (ix >= W || ix + 1 < 0 || iy >= H || iy + 1 < 0) {
        v = 0.f;
      } else {
        const int x0 = ix, x1 = ix + 1, y0 = iy, y1 = iy + 1;
        const float v0 = ((unsigned)x0 < (unsigned)W && (unsigned)y0 < (unsigned)H) ? src[(size_t)x0 * s_xs + (size_t)y0 * s_ys] : 0.f;
        const float v1 = ((unsigned)x1 < (unsigned)W && (unsigned)y0 < (unsigned)H) ? src[(size_t)x1 * s_xs + (size_t)y0 * s_ys] : 0.f;
        const float v2 = ((unsigned)x0 < (unsigned)W && (unsigned)y1 < (unsigned)H) ? src[(size_t)x0 * s_xs + (size_t)y1 * s_ys] : 0.f;
        const float v3 = ((unsigned)x1 < (unsigned)W && (unsigned)y1 < (unsigned)H) ? src[(size_t)x1 * s_xs + (size_t)y1 * s_ys] : 0.f;
        v = v0 * w[0] + v1 * w[1] + v2 * w[2] + v3 * w[3];
      }
      dst[(size_t)phi * d_ps + (size_t)rho * d_rs] = v;
    }
  }
}

/* src: H x W column-major, dst: D x Cp column-major (rows = angle, cols = radius); centre and radius as CorrelationFlow::polar */
void orc_polar(const float *src, int H, int W, int D, int Cp, float *dst) {
  const double maxRadius = (double)((H / 2) < (W / 2) ? (H / 2) : (W / 2));
  warp_polar_core(src, H, W, 1, (size_t)H, D, Cp, (float)W / 2, (float)H / 2, maxRadius, dst, 1, (size_t)D);
}
/* row-major entry point (cv::Mat data) used by oracle/ref_stubs (the cv::warpPolar stand-in of oracle/_ref) */
void orc_warp_polar_rm(const float *src, int H, int W, int D, int Cp, float cx, float cy, double maxRadius, float *dst) {
  warp_polar_core(src, H, W, (size_t)W, 1, D, Cp, cx, cy, maxRadius, dst, (size_t)Cp, 1);
}

static inline int wrap_idx(int p, int len) {
  if ((unsigned)p < (unsigned)len) return p;
  if (p < 0) p -= ((p - len + 1) / len) * len;
  if (p >= len) p %= len;
  return p;
}

/* inverse affine matrix of cv::getRotationMatrix2D((W/2.,H/2.), degree, 1) as cv::warpAffine computes it */
void orc_rotation_inverse(int H, int W, double degree, double iM[6]) {
  const float cxf = (float)(W / 2.), cyf = (float)(H / 2.);
  const double a = degree * (M_PI / 180.0);
  const double alpha = cos(a), beta = sin(a);
  double M[6];
  M[0] = alpha; M[1] = beta; M[2] = (1 - alpha) * cxf - beta * cyf;
  M[3] = -beta; M[4] = alpha; M[5] = beta * cxf + (1 - alpha) * cyf;
  double Dt = M[0] * M[4] - M[1] * M[3];
  Dt = Dt != 0 ? 1. / Dt : 0;
  const double A11 = M[4] * Dt, A22 = M[0] * Dt;
  M[0] = A11; M[1] *= -Dt; M[3] *= -Dt; M[4] = A22;
  const double b1 = -M[0] * M[2] - M[1] * M[5];
  const double b2 = -M[3] * M[2] - M[4] * M[5];
  M[2] = b1; M[5] = b2;
  memcpy(iM, M, sizeof(M));
}

/* cv::warpAffine(INTER_LINEAR, BORDER_WRAP) given the already inverted matrix iM: AB_BITS=10, INTER_BITS=5.
 * Stride-generic like warp_polar_core: element (y, x) at [y*s_ys + x*s_xs]. */
static inline __attribute__((always_inline)) void warp_affine_core(const float *src, int H, int W, size_t s_ys, size_t s_xs,
                                                                   const double M[6], float *dst) {
  for (int y = 0; y < H; ++y) {
    const int X0 = cv_round_d((M[1] * y + M[2]) * 1024.0) + 16;
    const int Y0 = cv_round_d((M[4] * y + M[5]) * 1024.0) + 16;
    for (int x = 0; x < W; ++x) {
      const int adelta = cv_round_d(M[0] * x * 1024.0), bdelta = cv_round_d(M[3] * x * 1024.0);
      const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
      const int ix = sat_short(X >> 5), iy = sat_short(Y >> 5);
      float w[4];
      bil_w(X & 31, Y & 31, w);
      float v;
      if ((unsigned)ix < (unsigned)(W - 1) && (unsigned)iy < (unsigned)(H - 1)) {
        const float *S = src + (size_t)ix * s_xs + (size_t)iy * s_ys;
        v = S[0] * w[0] + S[s_xs] * w[1] + S[s_ys] * w[2] + S[s_xs + s_ys] * w[3];
      } else {
        const int x0 = wrap_idx(ix, W), x1 = wrap_idx(ix + 1, W), y0 = wrap_idx(iy, H), y1 = wrap_idx(iy + 1, H);
        v = src[(size_t)x0 * s_xs + (size_t)y0 * s_ys] * w[0] + src[(size_t)x1 * s_xs + (size_t)y0 * s_ys] * w[1] +
            src[(size_t)x0 * s_xs + (size_t)y1 * s_ys] * w[2] + src[(size_t)x1 * s_xs + (size_t)y1 * s_ys] * w[3];
      }
      dst[(size_t)x * s_xs + (size_t)y * s_ys] = v;
    }
  }
}

/* utils.cc:154-161 RotateArray: getRotationMatrix2D + warpAffine(INTER_LINEAR, BORDER_WRAP).  col-major in/out. */
void orc_rotate(const float *src, int H, int W, float degree, float *dst) {
  double M[6];
  orc_rotation_inverse(H, W, (double)degree, M);
  warp_affine_core(src, H, W, 1, (size_t)H, M, dst);
}
/* row-major entry point (cv::Mat data) used by oracle/ref_stubs (the cv::warpAffine stand-in of oracle/_ref) */
void orc_warp_affine_inv_rm(const float *src, int H, int W, const double iM[6], float *dst) {
  warp_affine_core(src, H, W, (size_t)W, 1, iM, dst);
}

/* Camera::UndistortImage (camera.cc:92-93): cv::remap(u8, map1 CV_16SC2, map2 CV_16UC1, INTER_LINEAR), BORDER_CONSTANT(0).
 * OpenCV's u8 bilinear remap is integer: BilinearTab_i weights = round(w * 2^15) (exact multiples of 32 for 1/32 fractions),
 * FixedPtCast: (sum + 2^14) >> 15.  raw/out row-major H x W; map1 = (x,y) int16 pairs; map2 = fy*32 + fx. */
void orc_undistort_u8(const uint8_t *raw, int H, int W, const int16_t *map1, const uint16_t *map2, uint8_t *out) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int sx = map1[2 * ((size_t)y * W + x)], sy = map1[2 * ((size_t)y * W + x) + 1];
      const int a = map2[(size_t)y * W + x] & 1023, fx = a & 31, fy = a >> 5;
      const int w0 = (32 - fy) * (32 - fx) * 32, w1 = (32 - fy) * fx * 32, w2 = fy * (32 - fx) * 32, w3 = fy * fx * 32;
      int t[4];
      for (int k = 0; k < 4; ++k) {
        const int xx = sx + (k & 1), yy = sy + (k >> 1);
        t[k] = ((unsigned)xx < (unsigned)W && (unsigned)yy < (unsigned)H) ? raw[(size_t)yy * W + xx] : 0;
      }
      int v = (t[0] * w0 + t[1] * w1 + t[2] * w2 + t[3] * w3 + (1 << 14)) >> 15;
      out[(size_t)y * W + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

/* ------------------------------------------------------------------------------------------ */
/* CorrelationFlow                                                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int height, width; /* overridden by the ctor (correlation_flow.cc:40-41) */
  float lambda;
  int kernel;
  float sigma, offset;
  int power;
  int rotation_divisor, rotation_channel;
} orc_cf_config;

typedef struct {
  int polar_row, polar_col, trans_row, trans_col, hyp;
  float degree;
} orc_peaks;

static inline size_t spec_len(int R, int C) { return (size_t)(R / 2 + 1) * C; }

/* correlation_flow.cc:89-95.  image: H x W col-major f32.  Outputs in reference layout. */
void orc_compute_intermedium(const orc_cf_config *cfg, const float *image, float *fft_result, float *fft_polar) {
  const int H = cfg->height, W = cfg->width, D = cfg->rotation_divisor, Cp = cfg->rotation_channel;
  const size_t n = spec_len(H, W);
  float *mag = (float *)malloc(sizeof(float) * 2 * n);
  float *power = (float *)malloc(sizeof(float) * (size_t)H * W);
  float *hp = (float *)malloc(sizeof(float) * (size_t)H * W);
  float *pol = (float *)malloc(sizeof(float) * (size_t)D * Cp);
  orc_fft2(image, H, W, fft_result);
  for (size_t i = 0; i < n; ++i) {
    mag[2 * i] = hypotf(fft_result[2 * i], fft_result[2 * i + 1]); /* std::abs(complex<float>) */
    mag[2 * i + 1] = 0.f;
  }
  orc_ifft2(mag, H, W, power);
  orc_remove_zero_component(power, H, W, hp);
  orc_fftshift(hp, H, W, power);
  orc_polar(power, H, W, D, Cp, pol);
  orc_fft2(pol, D, Cp, fft_polar);
  free(mag); free(power); free(hp); free(pol);
}

/* kernels: correlation_flow.cc:181-226.  xf,zf half spectra (R/2+1)xC; out likewise. zf==NULL -> auto form. */
static int kernel_fft(const orc_cf_config *cfg, const float *xf, const float *zf, int R, int C, float *out, float *tmp_spec,
                      float *tmp_real) {
  const size_t n = spec_len(R, C), N = (size_t)R * C;
  const float *z = zf ? zf : xf;
  for (size_t i = 0; i < n; ++i) {
    const float xr = xf[2 * i], xi = xf[2 * i + 1], zr = z[2 * i], zi = -z[2 * i + 1];
    tmp_spec[2 * i] = xr * zr - xi * zi;
    tmp_spec[2 * i + 1] = xr * zi + xi * zr;
  }
  orc_ifft2(tmp_spec, R, C, tmp_real);
  float mx = 0.f;
  if (cfg->kernel == 0) {
    for (size_t i = 0; i < N; ++i) {
      /* Eigen 3.3 pow(float array, int) -> std::pow(float,int) -> double pow */
      const float k = (float)pow((double)(tmp_real[i] + cfg->offset), (double)cfg->power);
      tmp_real[i] = k;
      const float a = fabsf(k);
      if (a > mx) mx = a;
    }
  } else if (cfg->kernel == 1) {
    /* quirk kept: sums run over the stored half spectrum only (correlation_flow.cc:184-185) */
    /* |z^2| of a complex array has no packet form in Eigen: the sum is a plain f32 running sum in storage order (pinned by oracle/_ref) */
    float xx = 0.f, zz = 0.f;
    {
      float sx = 0.f, sz = 0.f;
      for (size_t i = 0; i < n; ++i) {
        const float ar = xf[2 * i] * xf[2 * i] - xf[2 * i + 1] * xf[2 * i + 1], ai = 2.f * xf[2 * i] * xf[2 * i + 1];
        const float hx = hypotf(ar, ai);
        sx = i ? sx + hx : hx;
        const float br = z[2 * i] * z[2 * i] - z[2 * i + 1] * z[2 * i + 1], bi = 2.f * z[2 * i] * z[2 * i + 1];
        const float hz = hypotf(br, bi);
        sz = i ? sz + hz : hz;
      }
      xx = sx / (float)(unsigned)N;
      zz = sz / (float)(unsigned)N;
    }
    const float coef = -1.f / (cfg->sigma * cfg->sigma);
    for (size_t i = 0; i < N; ++i) {
      const float d = (xx + zz - 2.f * tmp_real[i]) / (float)(unsigned)N;
      const float k = expf(coef * d);
      tmp_real[i] = k;
      const float a = fabsf(k);
      if (a > mx) mx = a;
    }
  } else {
    return -1; /* std::invalid_argument("Received invalid kernel type") */
  }
  for (size_t i = 0; i < N; ++i) tmp_real[i] = tmp_real[i] / mx;
  orc_fft2(tmp_real, R, C, out);
  return 0;
}

/* Eigen 3.3 sum() of a real f32 array as the reference builds it (-O3 -march=native, AVX): Redux.h linear vectorised traversal,
 * 8-float packets, two packet accumulators, horizontal add ((a0+a4)+(a1+a5))+((a2+a6)+(a3+a7)), scalar tail.
 * f(i) produces element i (GetInfo's variance reduces the expression (g-m)^2 without a temporary). */
#define ORC_EIGEN_SUM(n, ELEM, result)                                                         \
  do {                                                                                         \
    const size_t n_ = (n), a2_ = (n_ / 16) * 16, a1_ = (n_ / 8) * 8;                           \
    float p0_[8], p1_[8], res_;                                                                \
    if (a1_ == 0) {                                                                            \
      res_ = n_ ? ELEM(0) : 0.f;                                                               \
      for (size_t i_ = 1; i_ < n_; ++i_) res_ = res_ + ELEM(i_);                               \
    } else {                                                                                   \
      for (int k_ = 0; k_ < 8; ++k_) p0_[k_] = ELEM((size_t)k_);                               \
      if (a1_ > 8) {                                                                           \
        for (int k_ = 0; k_ < 8; ++k_) p1_[k_] = ELEM((size_t)(8 + k_));                       \
        for (size_t i_ = 16; i_ < a2_; i_ += 16)                                               \
          for (int k_ = 0; k_ < 8; ++k_) {                                                     \
            p0_[k_] = p0_[k_] + ELEM(i_ + k_);                                                 \
            p1_[k_] = p1_[k_] + ELEM(i_ + 8 + k_);                                             \
          }                                                                                    \
        for (int k_ = 0; k_ < 8; ++k_) p0_[k_] = p0_[k_] + p1_[k_];                            \
        if (a1_ > a2_)                                                                         \
          for (int k_ = 0; k_ < 8; ++k_) p0_[k_] = p0_[k_] + ELEM(a2_ + k_);                   \
      }                                                                                        \
      res_ = ((p0_[0] + p0_[4]) + (p0_[1] + p0_[5])) + ((p0_[2] + p0_[6]) + (p0_[3] + p0_[7])); \
      for (size_t i_ = a1_; i_ < n_; ++i_) res_ = res_ + ELEM(i_);                             \
    }                                                                                          \
    (result) = res_;                                                                           \
  } while (0)

/* correlation_flow.cc:238-243: all sums in f32, in Eigen's evaluation order (pinned by oracle/_ref) */
static float get_info(const float *g, size_t n, float response) {
  float s, q;
#define G_AT(i) (g[(i)])
  ORC_EIGEN_SUM(n, G_AT, s);
#undef G_AT
  const float mean = (s - response) / (float)(n - 1);
#define SQ_AT(i) ((g[(i)] - mean) * (g[(i)] - mean))
  ORC_EIGEN_SUM(n, SQ_AT, q);
#undef SQ_AT
  const float sd = sqrtf(q / (float)n);
  return (float)((double)(response - mean) / ((double)sd + 1e-7));
}

/* correlation_flow.cc:145-179.  Returns info; trans[2] = {-(row-h/2), -(col-w/2)}; peak = {row,col}; g_out optional. */
int orc_estimate_trans(const orc_cf_config *cfg, const float *last_fft, const float *cur_fft, int R, int C, int trans[2],
                       int peak[2], float *info, float *g_out) {
  const size_t n = spec_len(R, C), N = (size_t)R * C;
  float *kzz = (float *)malloc(sizeof(float) * 2 * n), *kxz = (float *)malloc(sizeof(float) * 2 * n);
  float *tmp = (float *)malloc(sizeof(float) * 2 * n), *g = (float *)malloc(sizeof(float) * N);
  int rc = kernel_fft(cfg, last_fft, NULL, R, C, kzz, tmp, g);
  if (rc == 0) rc = kernel_fft(cfg, cur_fft, last_fft, R, C, kxz, tmp, g);
  if (rc != 0) { free(kzz); free(kxz); free(tmp); free(g); return rc; }
  const int half = R / 2 + 1;
  for (int c = 0; c < C; ++c)
    for (int k = 0; k < half; ++k) {
      const size_t i = (size_t)c * half + k;
      /* target = FFT(delta[R/2,C/2]) = (-1)^(k+c) (R,C even) */
      const float t = ((k + c) & 1) ? -1.f : 1.f;
      const float dr = kzz[2 * i] + cfg->lambda, di = kzz[2 * i + 1];
      const float den = dr * dr + di * di;
      const float hr = t * dr / den, hi = -t * di / den; /* t / (dr + i di) */
      tmp[2 * i] = hr * kxz[2 * i] - hi * kxz[2 * i + 1];
      tmp[2 * i + 1] = hr * kxz[2 * i + 1] + hi * kxz[2 * i];
    }
  orc_ifft2(tmp, R, C, g);
  size_t best = 0;
  for (size_t i = 1; i < N; ++i)
    if (g[i] > g[best]) best = i; /* column-major first maximum */
  const int col = (int)(best / R), row = (int)(best % R);
  trans[0] = -(row - R / 2);
  trans[1] = -(col - C / 2);
  peak[0] = row; peak[1] = col;
  *info = get_info(g, N, g[best]);
  if (g_out) memcpy(g_out, g, sizeof(float) * N);
  free(kzz); free(kxz); free(tmp); free(g);
  return 0;
}

/* correlation_flow.cc:97-143 (minus the two std::cout lines and the unused `rectify`, :139-141). */
int orc_compute_pose(const orc_cf_config *cfg, const float *last_fft_result, const float *image, const float *last_fft_polar,
                     const float *fft_polar, int not_large_rotation, double pose[3], double info[3], orc_peaks *peaks) {
  const int H = cfg->height, W = cfg->width, D = cfg->rotation_divisor, Cp = cfg->rotation_channel;
  int rots[2], prot[2], rc;
  float info_rots;
  rc = orc_estimate_trans(cfg, last_fft_polar, fft_polar, D, Cp, rots, prot, &info_rots, NULL);
  if (rc) return rc;
  float degree = (float)((double)rots[0] * (2.0 / D) * 180);
  degree = (float)orc_normalize_degree((double)degree);
  float *rot = (float *)malloc(sizeof(float) * (size_t)H * W);
  float *spec = (float *)malloc(sizeof(float) * 2 * spec_len(H, W));
  int trans[2], pk[2], hyp = 0;
  float info_trans;
  if (not_large_rotation) {
    degree = fabsf(degree) > 90 ? degree - 180 : degree;
    orc_rotate(image, H, W, -degree, rot);
    orc_fft2(rot, H, W, spec);
    rc = orc_estimate_trans(cfg, last_fft_result, spec, H, W, trans, pk, &info_trans, NULL);
  } else {
    int tv[2], pv[2];
    float iv;
    orc_rotate(image, H, W, -degree, rot);
    orc_fft2(rot, H, W, spec);
    rc = orc_estimate_trans(cfg, last_fft_result, spec, H, W, trans, pk, &info_trans, NULL);
    orc_rotate(image, H, W, -degree + 180, rot);
    orc_fft2(rot, H, W, spec);
    if (!rc) rc = orc_estimate_trans(cfg, last_fft_result, spec, H, W, tv, pv, &iv, NULL);
    if (!(info_trans > iv)) {
      info_trans = iv; trans[0] = tv[0]; trans[1] = tv[1]; pk[0] = pv[0]; pk[1] = pv[1];
      degree = degree + 180;
      hyp = 1;
    }
  }
  free(rot); free(spec);
  if (rc) return rc;
  if (degree > 180) degree = degree - 360;
  const float theta = (float)((double)(degree / 180) * M_PI);
  info[0] = info_trans; pose[0] = trans[1];
  info[1] = info_trans; pose[1] = trans[0];
  info[2] = info_rots;  pose[2] = theta;
  if (peaks) {
    peaks->polar_row = prot[0]; peaks->polar_col = prot[1];
    peaks->trans_row = pk[0]; peaks->trans_col = pk[1];
    peaks->hyp = hyp; peaks->degree = degree;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* LoopClosure scan  (loop_closure.cc:36-73)                                                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  double position_response_thr, angle_response_thr;
  int frame_gap_thr;
  double distance_thr;
} orc_loop_config;

typedef struct {
  int found;
  int index;     /* index into the candidate list, -1 if none evaluated */
  int frame_id;
  double relative_pose[3];
  double response[3];
} orc_loop_result;

/* Candidates are given as arrays of pointers (reference layout spectra).  threads<=1: the reference's serial loop;
 * threads>1: candidates split across OpenMP threads, merged in iteration order with the same strict '>' rule. */
int orc_find_loop_closure(const orc_cf_config *cfg, const orc_loop_config *thr, const float *image, const float *cur_fft_polar,
                          int cur_id, double cur_dist, int n, const float *const *fft_results, const float *const *fft_polars,
                          const int *frame_ids, const double *dists, int threads, orc_loop_result *out) {
  double *resp = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
  double *pose = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
  char *eval = (char *)calloc((size_t)(n > 0 ? n : 1), 1);
  int err = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
  for (int i = 0; i < n; ++i) {
    if (thr->frame_gap_thr > 0 && abs(cur_id - frame_ids[i]) < thr->frame_gap_thr) continue;
    if (thr->distance_thr > 0 && fabs(cur_dist - dists[i]) < thr->distance_thr) continue;
    int rc = orc_compute_pose(cfg, fft_results[i], image, fft_polars[i], cur_fft_polar, 0, pose + 3 * i, resp + 3 * i, NULL);
    if (rc) err = rc;
    eval[i] = 1;
  }
  out->found = 0; out->index = -1; out->frame_id = -1;
  out->response[0] = out->response[1] = out->response[2] = -1.0; /* loop_closure.h:15 */
  out->relative_pose[0] = out->relative_pose[1] = out->relative_pose[2] = 0.0;
  for (int i = 0; i < n; ++i) {
    if (!eval[i]) continue;
    const double s = resp[3 * i] + resp[3 * i + 1] + resp[3 * i + 2];
    const double b = out->response[0] + out->response[1] + out->response[2];
    if (s > b) {
      memcpy(out->response, resp + 3 * i, sizeof(double) * 3);
      memcpy(out->relative_pose, pose + 3 * i, sizeof(double) * 3);
      out->index = i;
      out->frame_id = frame_ids[i];
    }
  }
  out->found = (out->response[0] > thr->position_response_thr) && (out->response[2] > thr->angle_response_thr);
  free(resp); free(pose); free(eval);
  return err;
}

/* ------------------------------------------------------------------------------------------ */
/* batched drivers used as the timed CPU baseline                                              */
/* ------------------------------------------------------------------------------------------ */
/* Tracking stream with "every frame is a keyframe" (SURVEY 8d): frame t is solved against frame t-1.
 * frames: n u8 images row-major.  poses/infos: (n-1) x 3.  Work per solve = ComputeIntermedium(frame t) + ComputePose. */
int orc_track_stream(const orc_cf_config *cfg, const uint8_t *frames, int n, int threads, double *poses, double *infos) {
  const int H = cfg->height, W = cfg->width, D = cfg->rotation_divisor, Cp = cfg->rotation_channel;
  const size_t nt = 2 * spec_len(H, W), np = 2 * spec_len(D, Cp), npx = (size_t)H * W;
  float *F = (float *)malloc(sizeof(float) * nt * n), *P = (float *)malloc(sizeof(float) * np * n);
  float *img = (float *)malloc(sizeof(float) * npx * n);
  int err = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
  for (int t = 0; t < n; ++t) {
    orc_normalize_u8(frames + npx * t, H, W, img + npx * t);
    orc_compute_intermedium(cfg, img + npx * t, F + nt * t, P + np * t);
  }
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
  for (int t = 1; t < n; ++t) {
    int rc = orc_compute_pose(cfg, F + nt * (t - 1), img + npx * t, P + np * (t - 1), P + np * t, 1, poses + 3 * (t - 1),
                              infos + 3 * (t - 1), NULL);
    if (rc) err = rc;
  }
  free(F); free(P); free(img);
  return err;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
