/* glfer_headless.c -- `glfer -f file.wav -n N` without the GTK front end.
 *
 * The per-block loop of the reference's audio_available() (source.c:112-170) and the
 * estimator set-up of change_params() (source.c:267-350), written ONLY against the
 * estimator interface (fft.h / mtm.h / avg.h) and the two globals the estimators read
 * (`opt`, `glfer`, glfer.c:56-62).  The same source therefore builds two ways:
 *
 *   product : gcc -Iinclude tools/glfer_headless.c -Lglfer_b200 -lglfer_b200        (GPU)
 *   oracle  : gcc -DHEADLESS_REFERENCE -I/root/reference -Ioracle/shim ... fft.c mtm.c ...  (CPU,
 *             see oracle/Makefile target _ref/glfer_headless_ref)
 *
 * which is the drop-in claim in executable form: nothing but the library behind the headers
 * changes.  Output: one binary file of float32 rows [frames][N/2+1] (the PSD, or avg[] when
 * averaging is on), plus a one-line summary on stdout.
 *
 *   glfer_headless -f in.wav [-n 1024] [-w 0..7] [-o overlap] [-m fft|mtm|lmp] [-k kmax] [-W nw] [-L lmp_av]
 *                  [-A 0..3] [-d depth] [-b minbin:maxbin] [-s 0|1] [-B] -O rows.f32
 *   -B uses the library's batched entry point (glfer_gram_run_wav) instead of the per-block
 *      calls (product build only).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#ifdef HEADLESS_REFERENCE
#include "glfer.h"
#include "fft.h"
#include "mtm.h"
#include "avg.h"
#include "lmp.h"
opt_t opt;
glfer_t glfer;
#else
#include "fft.h"
#include "mtm.h"
#include "avg.h"
#include "lmp.h"
#include "glfer_b200.h"
/* layout-compatible stand-ins for the reference's globals (glfer.h:62-139): the library
   resolves them weakly by name, exactly as fft.c's `extern opt_t opt; extern glfer_t glfer;` */
typedef struct {
  char *program_name; int mode; int scale_type;
  int data_block_size; float data_blocks_overlap; float display_update_time; float limiter_a; int enable_limiter;
  float mtm_w; int mtm_k; int hparma_t; int hparma_p_e; int lmp_av; int window_type;
  char *audio_device; int sample_rate;
  float dot_time; float dfcw_gap_time; int tx_mode; float dash_dot_ratio; float ptt_delay; float sidetone_freq;
  int sidetone; float dfcw_dot_freq; float dfcw_dash_freq; int beacon_mode; float beacon_pause; int beacon_tx_pause;
  char *ctrl_device; int device_type;
  float offset_freq; float thr_level; int autoscale; float max_level_db; float min_level_db;
  int averaging; int avgsamples; float min_avgband; float max_avgband; int palette;
} opt_t;
typedef struct {
  void *tt; void *qso_menu_item; void *test_menu_item;
  int init_done; int first_buffer; int input_source; float cpu_usage; int current_mode;
  void *scope_window; float avgmax; double avgvar; int avgfill; float peakfreq; float peakval; float avgtime;
} glfer_t;
opt_t opt;
glfer_t glfer;
#define TRUE 1
#define FALSE 0
enum { MODE_FFT = 0, MODE_MTM = 1, MODE_LMP = 3 };      /* glfer.h:45 */
enum { NO_AVG = 0, AVG_SUMAVG, AVG_PLAIN, AVG_SUMEXTREME };
#endif

/* minimal LP64-safe reader with the reference's block semantics (wav_fmt.c:81-121) */
typedef struct { FILE *fh; int bits; int rate; int hop; float *buff; } wav_src;

static int wav_open(wav_src *w, const char *path, int hop)
{
  unsigned char hd[44];
  w->fh = fopen(path, "rb");
  if (!w->fh || fread(hd, 1, 44, w->fh) != 44) return -1;
  w->rate = hd[24] | (hd[25] << 8) | (hd[26] << 16) | (hd[27] << 24);
  w->bits = hd[34] | (hd[35] << 8);
  w->hop = hop;
  w->buff = calloc(hop, sizeof(float));
  return 0;
}

static int wav_next(wav_src *w)
{
  if (w->bits == 8) {
    unsigned char *b = malloc(w->hop);
    size_t n = fread(b, 1, w->hop, w->fh);
    for (size_t i = 0; i < n; i++) w->buff[i] = ((float) b[i] - 128) / 128;
    free(b);
    return n != 0;
  }
  short *b = malloc(sizeof(short) * w->hop);
  size_t n = fread(b, 1, sizeof(short) * w->hop, w->fh);
  for (size_t i = 0; i < n / 2; i++) w->buff[i] = (float) b[i] / 32768;
  free(b);
  return n != 0;
}

int main(int argc, char **argv)
{
  const char *in = NULL, *out = NULL;
  int n = 1024, window = 7, mode = MODE_FFT, kmax = 7, lmp_av = 4, avgmode = NO_AVG, depth = 4, minbin = -1, maxbin = -1;
  int autoscale = 1, batch = 0;
  float overlap = 0.0f, nw = 4.0f;
  int c;
  while ((c = getopt(argc, argv, "f:n:w:o:m:k:W:A:d:b:s:O:BL:")) != -1) {
    switch (c) {
    case 'f': in = optarg; break;
    case 'n': n = atoi(optarg); break;
    case 'w': window = atoi(optarg); break;
    case 'o': overlap = (float) atof(optarg); break;
    case 'm': mode = strcmp(optarg, "mtm") == 0 ? MODE_MTM : (strcmp(optarg, "lmp") == 0 ? MODE_LMP : MODE_FFT); break;
    case 'L': lmp_av = atoi(optarg); break;
    case 'k': kmax = atoi(optarg); break;
    case 'W': nw = (float) atof(optarg); break;
    case 'A': avgmode = atoi(optarg); break;
    case 'd': depth = atoi(optarg); break;
    case 'b': sscanf(optarg, "%d:%d", &minbin, &maxbin); break;
    case 's': autoscale = atoi(optarg); break;
    case 'O': out = optarg; break;
    case 'B': batch = 1; break;
    default: return 2;
    }
  }
  if (!in || !out) {
    fprintf(stderr, "usage: glfer_headless -f in.wav -O rows.f32 [-n N] [-w win] [-o ovl] [-m fft|mtm|lmp] ...\n");
    return 2;
  }
  const int bins = n / 2 + 1;
  const int hop = (int) (n * (1.0 - overlap));
  memset(&opt, 0, sizeof opt);
  memset(&glfer, 0, sizeof glfer);
  opt.data_block_size = n;
  opt.data_blocks_overlap = overlap;
  opt.window_type = window;
  opt.autoscale = autoscale;
  opt.mtm_w = nw;
  opt.mtm_k = kmax;
  opt.lmp_av = lmp_av;
  opt.averaging = avgmode;
  opt.avgsamples = depth;
  FILE *fo = fopen(out, "wb");
  if (!fo) { perror(out); return 1; }
  long frames = 0;
  double checksum = 0.0;

#ifndef HEADLESS_REFERENCE
  if (batch) {
    glfer_gram_config cfg;
    glfer_gram_config_default(&cfg);
    cfg.mode = mode; cfg.n = n; cfg.window_type = window; cfg.overlap = overlap; cfg.sub_mean = autoscale;
    cfg.mtm_w = nw; cfg.mtm_kmax = kmax; cfg.avg_mode = avgmode; cfg.avg_depth = depth; cfg.lmp_av = lmp_av;
    glfer_wav wav;
    if (glfer_wav_load(in, &wav) != 0) { fprintf(stderr, "%s\n", glfer_b200_last_error()); return 1; }
    const float binsize = (float) wav.sample_rate / (float) n;
    cfg.avg_minbin = minbin >= 0 ? minbin : (int) (400.0f / binsize);
    cfg.avg_maxbin = maxbin >= 0 ? maxbin : (int) (1200.0f / binsize);
    glfer_gram_plan *plan = NULL;
    if (glfer_gram_plan_create(&cfg, &plan) != 0) { fprintf(stderr, "%s\n", glfer_b200_last_error()); return 1; }
    frames = (long) glfer_wav_num_frames(plan, &wav);
    float *rows = malloc(sizeof(float) * (size_t) frames * bins);
    float *avg = avgmode != NO_AVG ? malloc(sizeof(float) * (size_t) frames * bins) : NULL;
    if (glfer_gram_run_wav(plan, &wav, rows, avg, NULL, NULL, NULL) != 0) { fprintf(stderr, "%s\n", glfer_b200_last_error()); return 1; }
    const float *res = avg ? avg : rows;
    fwrite(res, sizeof(float), (size_t) frames * bins, fo);
    for (long i = 0; i < frames * bins; i++) checksum += res[i];
    glfer_gram_plan_destroy(plan);
    glfer_wav_free(&wav);
    fclose(fo);
    printf("frames %ld bins %d hop %d mode %s path batch checksum %.9e\n", frames, bins, hop, mode == MODE_MTM ? "mtm" : (mode == MODE_LMP ? "lmp" : "fft"), checksum);
    return 0;
  }
#else
  (void) batch;
#endif

  /* change_params (source.c:320-350) */
  wav_src src;
  if (wav_open(&src, in, hop) != 0) { fprintf(stderr, "cannot read %s\n", in); return 1; }
  opt.sample_rate = src.rate;
  fft_params_t fft_par;
  mtm_params_t mtm_par;
  lmp_params_t lmp_par;
  memset(&fft_par, 0, sizeof fft_par);
  memset(&mtm_par, 0, sizeof mtm_par);
  memset(&lmp_par, 0, sizeof lmp_par);
  avg_data_t avgdata;
  init_avg(&avgdata);
  alloc_avg(&avgdata, n, depth);                        /* source.c:311-312: width N */
  float *psdbuf = calloc(bins, sizeof(float));
  if (mode == MODE_FFT) {
    fft_par.n = n; fft_par.window_type = window; fft_par.overlap = overlap; fft_par.a = 0.0f; fft_par.limiter = 0;
    fft_init(&fft_par);
  } else if (mode == MODE_LMP) {                        /* source.c:394-400 */
    lmp_par.fft.n = n; lmp_par.fft.window_type = RECTANGULAR_WINDOW; lmp_par.fft.overlap = overlap;
    lmp_par.fft.a = 0.0f; lmp_par.fft.limiter = 0; lmp_par.avg = lmp_av;
    lmp_init(&lmp_par);
  } else {
    mtm_par.fft.n = n; mtm_par.fft.window_type = RECTANGULAR_WINDOW; mtm_par.fft.overlap = overlap;
    mtm_par.w = nw; mtm_par.kmax = kmax;
    mtm_init(&mtm_par);
  }
  const float binsize = (float) opt.sample_rate / (float) n;   /* g_main.c:1144-1146 */
  if (minbin < 0) minbin = (int) (400.0f / binsize);
  if (maxbin < 0) maxbin = (int) (1200.0f / binsize);
  int peakbin = 0;
  double variance = 0.0;
  float *row = malloc(sizeof(float) * bins);
  glfer.first_buffer = TRUE;                             /* g_main.c:990 */
  while (wav_next(&src)) {                               /* audio_available, source.c:112-170 */
    if (mode == MODE_FFT) {
      fft_do(src.buff, &fft_par);
      fft_psd(psdbuf, NULL, &fft_par);
    } else if (mode == MODE_LMP) {
      lmp_do(src.buff, psdbuf, NULL, &lmp_par);        /* source.c:155-156 */
    } else {
      mtm_do(src.buff, psdbuf, NULL, &mtm_par);
    }
    glfer.first_buffer = FALSE;                          /* g_main.c:1120 */
    const float *res = psdbuf;
    switch (avgmode) {                                   /* g_main.c:1153-1183 */
    case AVG_SUMAVG: update_avg_sumavg(&avgdata, n, psdbuf, 0, minbin, maxbin, &peakbin, &variance); break;
    case AVG_PLAIN: update_avg_plain(&avgdata, n, psdbuf, minbin, maxbin, &peakbin); break;
    case AVG_SUMEXTREME: update_avg_sumextreme(&avgdata, n, psdbuf, 0, minbin, maxbin, &peakbin); break;
    default: break;
    }
    if (avgmode != NO_AVG) {
      for (int i = 0; i < bins; i++) row[i] = (float) avgdata.avg[i];
      res = row;
    }
    fwrite(res, sizeof(float), bins, fo);
    for (int i = 0; i < bins; i++) checksum += res[i];
    frames++;
  }
  if (mode == MODE_FFT) fft_close(&fft_par);
  else if (mode == MODE_LMP) lmp_close(&lmp_par);
  else mtm_close(&mtm_par);
  delete_avg(&avgdata);
  fclose(fo);
  printf("frames %ld bins %d hop %d mode %s path per-call checksum %.9e\n", frames, bins, hop, mode == MODE_MTM ? "mtm" : (mode == MODE_LMP ? "lmp" : "fft"), checksum);
  return 0;
}
