// stand-in for <fftw3.h> (FFTW3 single precision is not in this image): the two planner calls CorrelationFlow::FFT / IFFT make
// (correlation_flow.cc:56-61, :70-74) mapped onto the C oracle's mixed-radix f32 FFT (oracle/nislam_oracle.c orc_fft2 /
// orc_ifft2_raw).  Same conventions as FFTW: row-major n0 x n1 real <-> n0 x (n1/2+1) complex, unnormalised both ways.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct fftwf_plan_s {
  int kind, n0, n1;
  void* in;
  void* out;
} * fftwf_plan;
#define FFTW_ESTIMATE (1U << 6)
fftwf_plan fftwf_plan_dft_r2c_2d(int n0, int n1, float* in, fftwf_complex* out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_2d(int n0, int n1, fftwf_complex* in, float* out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
void fftw_cleanup(void);
#ifdef __cplusplus
}
#endif
