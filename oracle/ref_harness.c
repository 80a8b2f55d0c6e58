/* Headless driver around the UNMODIFIED reference estimator sources.
 * TEST INFRASTRUCTURE ONLY: compiled by oracle/Makefile together with
 * /root/reference/{fft.c,fft_radix2.c,mtm.c,g-l_dpss.c,avg.c,lmp.c,util.c} into
 * oracle/_ref/libglfer_ref_{f32,f64}.so.  Nothing under oracle/ is linked into,
 * imported by or executed from the product library.
 *
 * This file replaces the GTK glue of the reference:
 *   - it owns the two globals the estimators read (`opt`, `glfer`;
 *     glfer.c:56-62, read at fft.c:99,186 and mtm.c:111);
 *   - its frame loop is the body of audio_available() (source.c:130-165)
 *     without the drawing: one hop block per call, fft_do + fft_psd or mtm_do;
 *   - glfer.first_buffer is TRUE for the first block after init and FALSE
 *     afterwards (what main_window_draw does with autoscale on,
 *     g_main.c:1111-1120).
 * Two builds exist: f32 = what glfer ships without FFTW (fft_radix2.c), used
 * for CPU timing; f64 = -DHAVE_LIBRFFTW over oracle/rfftw_shim.c, "the
 * reference's double-precision path", used as the golden.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>

#include "glfer.h"
#include "util.h"
#include "fft.h"
#include "mtm.h"
#include "avg.h"
#include "lmp.h"
#include "g-l_dpss.h"

opt_t opt;
glfer_t glfer;

/* 1: glfer.first_buffer stays TRUE for the whole run -- what the GUI does with opt.autoscale == 0,
 * where main_window_draw never clears it (g_main.c:1111-1120): every frame gets a zeroed history */
static int g_sticky_first_buffer = 0;
void refh_set_sticky_first_buffer(int on) { g_sticky_first_buffer = on; }

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int refh_is_double(void)
{
#ifdef HAVE_LIBRFFTW
  return 1;
#else
  return 0;
#endif
}

/* hop exactly as prepare_audio computes it (fft.c:70): double product, truncated */
int refh_hop(int n, float overlap)
{
  fft_params_t p;
  memset(&p, 0, sizeof p);
  p.n = n;
  p.overlap = overlap;
  int n_eff = p.n * (1.0 - p.overlap);
  return n_eff;
}

/* window table as fft_init builds it (fft.c:168-187 -> compute_window fft.c:309) */
int refh_window(int n, int window_type, float *w_out)
{
  fft_params_t p;
  memset(&p, 0, sizeof p);
  p.n = n;
  p.window_type = window_type;
  p.overlap = 0.0f;
  opt.autoscale = 0;
  fft_init(&p);
  memcpy(w_out, p.window, sizeof(float) * n);
  fft_close(&p);
  return 0;
}

/* Periodogram spectrogram: the MODE_FFT branch of source.c:141-146 per hop block.
 * rows: [nframes][n/2+1] float (may be NULL); spec: [nframes][n] half-complex as
 * double (may be NULL); phase: [nframes][n/2+1] (may be NULL).
 * Returns the number of frames produced = min(max_frames, nsamples / hop). */
long refh_periodogram(const float *samples, long nsamples, int n, int window_type,
                      float overlap, int sub_mean, float a, int limiter,
                      long max_frames, float *rows, double *spec, float *phase)
{
  fft_params_t p;
  memset(&p, 0, sizeof p);
  p.n = n;
  p.window_type = window_type;
  p.overlap = overlap;
  p.a = a;
  p.limiter = limiter;
  opt.autoscale = sub_mean;          /* fft_init copies it into sub_mean (fft.c:186) */
  fft_init(&p);
  int hop = refh_hop(n, overlap);
  int bins = n / 2 + 1;
  long nframes = hop > 0 ? nsamples / hop : 0;
  if (nframes > max_frames) nframes = max_frames;
  float *blk = malloc(sizeof(float) * (hop > 0 ? hop : 1));
  float *psd = malloc(sizeof(float) * bins);
  glfer.first_buffer = TRUE;
  for (long f = 0; f < nframes; f++) {
    memcpy(blk, samples + f * hop, sizeof(float) * hop);   /* fft_do mutates its input */
    fft_do(blk, &p);
    fft_psd(rows ? rows + f * bins : psd, phase ? phase + f * bins : NULL, &p);
    if (spec)
      for (int i = 0; i < n; i++) spec[f * n + i] = p.outbuf[i];
    if (!g_sticky_first_buffer) glfer.first_buffer = FALSE;
  }
  free(blk);
  free(psd);
  fft_close(&p);
  return nframes;
}

/* DPSS tapers and eigenvalues as mtm_init builds them (mtm.c:88-122 ->
 * compute_dspwf mtm.c:63 -> gl_dpss g-l_dpss.c:288).  tapers: [kmax+1][n] double,
 * lambda[k] = 1 + sig[k]. */
int refh_dpss(int n, float w, int kmax, double *tapers, double *lambda)
{
  mtm_params_t p;
  memset(&p, 0, sizeof p);
  p.fft.n = n;
  p.fft.window_type = RECTANGULAR_WINDOW;
  p.fft.overlap = 0.0f;
  p.w = w;
  p.kmax = kmax;
  opt.autoscale = 0;
  mtm_init(&p);
  for (int k = 0; k <= kmax; k++) {
    lambda[k] = 1.0 + p.sig[k];
    for (int i = 0; i < n; i++) tapers[(long) k * n + i] = p.window[i + 1][k];
  }
  mtm_close(&p);
  return 0;
}

/* Multitaper spectrogram: the MODE_MTM branch of source.c:147-149, parameters set
 * as change_params does (source.c:343-350). */
long refh_mtm(const float *samples, long nsamples, int n, float overlap, int sub_mean,
              float a, int limiter, float w, int kmax, long max_frames, float *rows)
{
  mtm_params_t p;
  memset(&p, 0, sizeof p);
  p.fft.n = n;
  p.fft.window_type = RECTANGULAR_WINDOW;
  p.fft.overlap = overlap;
  p.fft.a = a;
  p.fft.limiter = limiter;
  p.w = w;
  p.kmax = kmax;
  opt.autoscale = sub_mean;
  mtm_init(&p);
  int hop = refh_hop(n, overlap);
  int bins = n / 2 + 1;
  long nframes = hop > 0 ? nsamples / hop : 0;
  if (nframes > max_frames) nframes = max_frames;
  float *blk = malloc(sizeof(float) * (hop > 0 ? hop : 1));
  float *psd = malloc(sizeof(float) * bins);
  glfer.first_buffer = TRUE;
  for (long f = 0; f < nframes; f++) {
    memcpy(blk, samples + f * hop, sizeof(float) * hop);
    mtm_do(blk, rows ? rows + f * bins : psd, NULL, &p);
    if (!g_sticky_first_buffer) glfer.first_buffer = FALSE;
  }
  free(blk);
  free(psd);
  mtm_close(&p);
  return nframes;
}

/* The same loop, also copying out the harmonic F-test mtm_do leaves in its file-static buffer
 * after every block (mtm.c:222-233; ref_mtm_unit.c hands the static out).  Meaningful in the
 * double (FFTW-layout) build only.  ftest: [nframes][n/2+1]. */
extern const float *refh_mtm_ftest_buffer(void);
long refh_mtm_ftest(const float *samples, long nsamples, int n, float overlap, int sub_mean,
                    float w, int kmax, long max_frames, float *rows, float *ftest_rows)
{
  mtm_params_t p;
  memset(&p, 0, sizeof p);
  p.fft.n = n;
  p.fft.window_type = RECTANGULAR_WINDOW;
  p.fft.overlap = overlap;
  p.w = w;
  p.kmax = kmax;
  opt.autoscale = sub_mean;
  mtm_init(&p);
  int hop = refh_hop(n, overlap);
  int bins = n / 2 + 1;
  long nframes = hop > 0 ? nsamples / hop : 0;
  if (nframes > max_frames) nframes = max_frames;
  float *blk = malloc(sizeof(float) * (hop > 0 ? hop : 1));
  glfer.first_buffer = TRUE;
  for (long f = 0; f < nframes; f++) {
    memcpy(blk, samples + f * hop, sizeof(float) * hop);
    mtm_do(blk, rows + f * bins, NULL, &p);
    memcpy(ftest_rows + f * bins, refh_mtm_ftest_buffer(), sizeof(float) * bins);
    if (!g_sticky_first_buffer) glfer.first_buffer = FALSE;
  }
  free(blk);
  mtm_close(&p);
  return nframes;
}

/* LMP spectrogram: the MODE_LMP branch of source.c:155-157, parameters set as change_params
 * does (source.c:394-400).  lmp.c keeps its ring index in a function-static that no init
 * resets (lmp.c:102); the harness feeds zero blocks after the run until that index is back at
 * 0, so that every call starts like a fresh process. */
long refh_lmp(const float *samples, long nsamples, int n, float overlap, int sub_mean,
              float a, int limiter, int nl, long max_frames, float *rows)
{
  lmp_params_t p;
  memset(&p, 0, sizeof p);
  p.fft.n = n;
  p.fft.window_type = RECTANGULAR_WINDOW;
  p.fft.overlap = overlap;
  p.fft.a = a;
  p.fft.limiter = limiter;
  p.avg = nl;
  opt.autoscale = sub_mean;
  lmp_init(&p);
  int hop = refh_hop(n, overlap);
  int bins = n / 2 + 1;
  long nframes = hop > 0 ? nsamples / hop : 0;
  if (nframes > max_frames) nframes = max_frames;
  float *blk = malloc(sizeof(float) * (hop > 0 ? hop : 1));
  float *psd = malloc(sizeof(float) * bins);
  glfer.first_buffer = TRUE;
  for (long f = 0; f < nframes; f++) {
    memcpy(blk, samples + f * hop, sizeof(float) * hop);
    lmp_do(blk, rows ? rows + f * bins : psd, NULL, &p);
    if (!g_sticky_first_buffer) glfer.first_buffer = FALSE;
  }
  for (long f = nframes; f % nl != 0; f++) {            /* realign the static ring index */
    memset(blk, 0, sizeof(float) * hop);
    lmp_do(blk, psd, NULL, &p);
  }
  free(blk);
  free(psd);
  lmp_close(&p);
  return nframes;
}

/* Frame averaging over a sequence of PSD rows: the avg dispatch of
 * main_window_draw (g_main.c:1153-1183).  mode follows avgmode_t (glfer.h:53-55):
 * 1 = AVG_SUMAVG, 2 = AVG_PLAIN, 3 = AVG_SUMEXTREME.  width is the alloc_avg width
 * (the caller passes N, source.c:312); nbins_out (<= width) columns of avg[] are
 * copied out per frame.  peakbin carries over between frames as the caller's
 * variable does. */
int refh_avg(int mode, int width, int depth, int minbin, int maxbin, int max0,
             const float *psd_rows, long nframes, long row_stride, int nbins_out,
             double *avg_rows, double *ret, int *peakbin_out, double *variance_out,
             int peakbin_init)
{
  avg_data_t ad;
  init_avg(&ad);
  alloc_avg(&ad, width, depth);
  int peakbin = peakbin_init;
  /* the reference passes the psd buffer of n/2+1 floats but loops to N; give it a
     padded copy so reads stay in bounds for any minbin/maxbin <= width */
  float *psd = calloc(width + 1, sizeof(float));
  for (long f = 0; f < nframes; f++) {
    double r = 0.0, var = 0.0;
    memcpy(psd, psd_rows + f * row_stride, sizeof(float) * (row_stride < width ? row_stride : width));
    switch (mode) {
    case AVG_SUMAVG:
      r = update_avg_sumavg(&ad, width, psd, max0, minbin, maxbin, &peakbin, &var);
      break;
    case AVG_PLAIN:
      r = update_avg_plain(&ad, width, psd, minbin, maxbin, &peakbin);
      break;
    case AVG_SUMEXTREME:
      r = update_avg_sumextreme(&ad, width, psd, max0, minbin, maxbin, &peakbin);
      break;
    default:
      free(psd);
      delete_avg(&ad);
      return -1;
    }
    if (ret) ret[f] = r;
    if (peakbin_out) peakbin_out[f] = peakbin;
    if (variance_out) variance_out[f] = var;
    if (avg_rows) memcpy(avg_rows + f * nbins_out, ad.avg, sizeof(double) * nbins_out);
  }
  free(psd);
  delete_avg(&ad);
  return 0;
}

/* compute_floor (fft.c:240) on one row */
void refh_floor(float *psd, int n, float *sig, float *floor_pwr, float *peak, unsigned int *peak_bin)
{
  compute_floor(psd, n, sig, floor_pwr, peak, peak_bin);
}

/* Timed loops for the CPU baseline (results discarded except a checksum so the
 * compiler cannot drop the work).  Returns seconds; *frames_out = frames done. */
double refh_time_periodogram(const float *samples, long nsamples, int n, int window_type,
                             float overlap, int sub_mean, long *frames_out, double *checksum)
{
  fft_params_t p;
  memset(&p, 0, sizeof p);
  p.n = n;
  p.window_type = window_type;
  p.overlap = overlap;
  opt.autoscale = sub_mean;
  fft_init(&p);
  int hop = refh_hop(n, overlap);
  int bins = n / 2 + 1;
  long nframes = nsamples / hop;
  float *blk = malloc(sizeof(float) * hop);
  float *psd = malloc(sizeof(float) * bins);
  double acc = 0.0;
  glfer.first_buffer = TRUE;
  double t0 = now_s();
  for (long f = 0; f < nframes; f++) {
    memcpy(blk, samples + f * hop, sizeof(float) * hop);
    fft_do(blk, &p);
    fft_psd(psd, NULL, &p);
    acc += psd[(f * 7) % bins];
    if (!g_sticky_first_buffer) glfer.first_buffer = FALSE;
  }
  double t1 = now_s();
  free(blk);
  free(psd);
  fft_close(&p);
  if (frames_out) *frames_out = nframes;
  if (checksum) *checksum = acc;
  return t1 - t0;
}

double refh_time_mtm(const float *samples, long nsamples, int n, float overlap, int sub_mean,
                     float w, int kmax, long *frames_out, double *checksum)
{
  mtm_params_t p;
  memset(&p, 0, sizeof p);
  p.fft.n = n;
  p.fft.window_type = RECTANGULAR_WINDOW;
  p.fft.overlap = overlap;
  p.w = w;
  p.kmax = kmax;
  opt.autoscale = sub_mean;
  mtm_init(&p);
  int hop = refh_hop(n, overlap);
  int bins = n / 2 + 1;
  long nframes = nsamples / hop;
  float *blk = malloc(sizeof(float) * hop);
  float *psd = malloc(sizeof(float) * bins);
  double acc = 0.0;
  glfer.first_buffer = TRUE;
  double t0 = now_s();
  for (long f = 0; f < nframes; f++) {
    memcpy(blk, samples + f * hop, sizeof(float) * hop);
    mtm_do(blk, psd, NULL, &p);
    acc += psd[(f * 7) % bins];
    if (!g_sticky_first_buffer) glfer.first_buffer = FALSE;
  }
  double t1 = now_s();
  free(blk);
  free(psd);
  mtm_close(&p);
  if (frames_out) *frames_out = nframes;
  if (checksum) *checksum = acc;
  return t1 - t0;
}
