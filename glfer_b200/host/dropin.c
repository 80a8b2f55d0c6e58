/* dropin.c -- the reference's per-call estimator interface (fft.h:77-83, mtm.h:47-49,
 * avg.h:38-43) on top of the GPU engine: one hop block per call, exactly the sequence
 * audio_available() drives (source.c:130-165).
 *
 * Host side of a call = what the caller can observe in the public structs: the hop-block
 * mean removal done in place in the caller's buffer (fft.c:86-96), the overlap history in
 * params->inbuf_audio (fft.c:98-113), the windowed frame in inbuf_fft (read by
 * g_scope.c:186-197) and the half-complex spectrum in outbuf.  The spectrum estimate
 * itself (window/taper multiply, FFT, |X|^2, taper sum, averaging) runs on the GPU as a
 * batch of one frame through the same kernels as the batched path.
 *
 * Engines are kept in a registry keyed by the params pointer because the reference's
 * structs have no spare field for a handle.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "glb_host.h"

/* ------------------------------------------------------------------ hidden globals
 * Layout mirrors of the two globals fft.c reads (glfer.h:62-139); GTK pointers are plain
 * pointers here.  They are weak: when the host program defines `opt` and `glfer`
 * (glfer.c:56-62) the library reads them, otherwise the setters below provide the values. */
typedef struct {
  char *program_name; int mode; int scale_type;
  int data_block_size; float data_blocks_overlap; float display_update_time; float limiter_a; int enable_limiter;
  float mtm_w; int mtm_k;
  int hparma_t; int hparma_p_e;
  int lmp_av;
  int window_type;
  char *audio_device; int sample_rate;
  float dot_time; float dfcw_gap_time; int tx_mode; float dash_dot_ratio; float ptt_delay; float sidetone_freq;
  int sidetone; float dfcw_dot_freq; float dfcw_dash_freq; int beacon_mode; float beacon_pause; int beacon_tx_pause;
  char *ctrl_device; int device_type;
  float offset_freq; float thr_level; int autoscale; float max_level_db; float min_level_db;
  int averaging; int avgsamples; float min_avgband; float max_avgband; int palette;
} glb_opt_mirror;

typedef struct {
  void *tt; void *qso_menu_item; void *test_menu_item;
  int init_done; int first_buffer; int input_source; float cpu_usage; int current_mode;
  void *scope_window;
  float avgmax; double avgvar; int avgfill; float peakfreq; float peakval; float avgtime;
} glb_glfer_mirror;

extern glb_opt_mirror opt __attribute__((weak));
extern glb_glfer_mirror glfer __attribute__((weak));

static int g_autoscale = 1;        /* glfer.c default: opt.autoscale = 1 */
static int g_first_buffer = 1;

void glfer_b200_set_autoscale(int v) { g_autoscale = v; }
void glfer_b200_set_first_buffer(int v) { g_first_buffer = v; }
int glb_autoscale(void) { return (&opt != NULL) ? opt.autoscale : g_autoscale; }
int glb_first_buffer(void) { return (&glfer != NULL) ? glfer.first_buffer : g_first_buffer; }

void glb_fatal(const char *where)
{
  fprintf(stderr, "libglfer_b200: %s: %s\n", where, glfer_b200_last_error());
  exit(-1);
}

/* fft.c:48-60 */
fft_window_t fft_windows[] = {
  {"/Hanning", HANNING_WINDOW}, {"/Blackman", BLACKMAN_WINDOW}, {"/Gaussian", GAUSSIAN_WINDOW},
  {"/Welch", WELCH_WINDOW}, {"/Bartlett", BARTLETT_WINDOW}, {"/Rectangular", RECTANGULAR_WINDOW},
  {"/Hamming", HAMMING_WINDOW}, {"/Kaiser", KAISER_WINDOW}
};
int num_fft_windows = sizeof(fft_windows) / sizeof(fft_windows[0]);

/* ------------------------------------------------------------------ engine registry */
typedef struct engine {
  const void *key;
  struct engine *next;
  int n, ntapers;
  float taper_scale;
  void *tables, *stream;
  float *d_frame, *d_tapers, *d_psd, *d_spec, *d_hc, *d_phase;
  float *h_psd, *h_spec;     /* pinned */
  float *h_frame;            /* FFTW layout: the frame converted to float for the upload (pinned) */
  int fresh;                 /* h_psd holds the PSD of the last *_do on this key */
  /* averaging engines */
  int width, depth;
  long long frames;          /* frames since alloc_avg */
  float *d_ring;             /* [depth][width] last PSD rows */
  double *d_avg, *d_ret, *d_var;
  int *d_cand;
  float *d_lmp;              /* LMP engines: the statistic row (ring = d_ring [depth][width]) */
  /* k-block streaming entries (fft_do_batch / mtm_do_batch): grow-only stream + rows buffers */
  float *d_bstream, *d_brows, *h_bstream, *h_brows;
  size_t bstream_cap, brows_cap;
  /* fft_psd with phase / compute_floor: buffers kept for the life of the engine (no cudaMalloc per call) */
  float *d_scratch;
  size_t scratch_cap;
} engine;

static engine *g_engines;

static engine *find_engine(const void *key)
{
  for (engine *e = g_engines; e; e = e->next)
    if (e->key == key) return e;
  return NULL;
}

static engine *new_engine(const void *key)
{
  engine *e = calloc(1, sizeof *e);
  if (!e) { glb_set_error("out of memory"); glb_fatal("engine"); }
  e->key = key;
  e->next = g_engines;
  g_engines = e;
  return e;
}

static void drop_engine(const void *key)
{
  for (engine **pp = &g_engines; *pp; pp = &(*pp)->next) {
    engine *e = *pp;
    if (e->key != key) continue;
    *pp = e->next;
    if (e->stream) glb_stream_sync(e->stream);
    glb_free(e->d_frame); glb_free(e->d_tapers); glb_free(e->d_psd); glb_free(e->d_spec);
    glb_free(e->d_hc); glb_free(e->d_phase); glb_free(e->d_ring); glb_free(e->d_avg);
    glb_free(e->d_ret); glb_free(e->d_var); glb_free(e->d_cand); glb_free(e->d_lmp);
    glb_host_free(e->h_psd); glb_host_free(e->h_spec); glb_host_free(e->h_frame);
    glb_free(e->d_bstream); glb_free(e->d_brows); glb_host_free(e->h_bstream); glb_host_free(e->h_brows);
    glb_free(e->d_scratch);
    glb_tables_destroy(e->tables);
    glb_stream_destroy(e->stream);
    free(e);
    return;
  }
}

#define MUST(x, where) do { if ((x) != GLB_OK) glb_fatal(where); } while (0)

static void need_device(const char *where)
{
  int c = 0;
  if (glb_device_count(&c) != GLB_OK || c < 1) {
    glb_set_error("no CUDA device (libglfer_b200 has no CPU fallback)");
    glb_fatal(where);
  }
}

/* grow-only device scratch of an engine (floats) */
static float *engine_scratch(engine *e, size_t floats)
{
  if (floats > e->scratch_cap) {
    if (e->d_scratch) MUST(glb_free(e->d_scratch), "free");
    e->d_scratch = NULL;
    MUST(glb_malloc((void **) &e->d_scratch, sizeof(float) * floats), "malloc");
    e->scratch_cap = floats;
  }
  return e->d_scratch;
}

/* the engine behind calls that have no params of their own (compute_floor, fft_psd on foreign params) */
static engine *scratch_engine(void)
{
  static const char key;
  engine *e = find_engine(&key);
  if (!e) {
    e = new_engine(&key);
    MUST(glb_stream_create(&e->stream), "stream");
  }
  return e;
}

/* device side of an estimator: tables, scaled tapers, one-frame buffers */
static engine *make_estimator(const void *key, int n, int ntapers, const float *scaled_tapers, float scale)
{
  need_device("init");
  if (!glb_fft_supported(n)) {
    fprintf(stderr, "libglfer_b200: n = %d is not a power of 2 in 32..32768\n", n);   /* cf. fft_radix2.c:89-93 */
    exit(-1);
  }
  drop_engine(key);
  engine *e = new_engine(key);
  e->n = n;
  e->ntapers = ntapers;
  e->taper_scale = scale;
  MUST(glb_stream_create(&e->stream), "stream");
  MUST(glb_tables_create(n, &e->tables), "tables");
  MUST(glb_malloc((void **) &e->d_frame, sizeof(float) * n), "malloc");
  MUST(glb_malloc((void **) &e->d_tapers, sizeof(float) * (size_t) ntapers * n), "malloc");
  MUST(glb_malloc((void **) &e->d_psd, sizeof(float) * (n / 2 + 1)), "malloc");
  MUST(glb_malloc((void **) &e->d_spec, sizeof(float) * 2 * (n / 2 + 1)), "malloc");
  MUST(glb_host_alloc((void **) &e->h_psd, sizeof(float) * (n / 2 + 1)), "host_alloc");
  MUST(glb_host_alloc((void **) &e->h_spec, sizeof(float) * 2 * (n / 2 + 1)), "host_alloc");
  MUST(glb_host_alloc((void **) &e->h_frame, sizeof(float) * n), "host_alloc");
  MUST(glb_memcpy_h2d(e->d_tapers, scaled_tapers, sizeof(float) * (size_t) ntapers * n, NULL), "h2d");
  return e;
}

/* one frame through gram_kernel: inbuf_audio -> PSD row (+ spectrum for one taper).
 * Latency path: no copy engine involved.  The frame sits in pinned host memory, which the device reads
 * directly (unified addressing: the host pointer IS the device pointer), and the kernel stores the PSD row
 * and the spectrum straight into pinned host buffers; a call is one host memcpy, one kernel launch and
 * one stream synchronisation instead of three cudaMemcpyAsync round trips (37 -> ~15 us at N = 1024). */
static int g_zero_copy = 1;
void glfer_b200_set_zero_copy(int on) { g_zero_copy = on; }

static void run_frame(engine *e, const float *frame, float a, int limiter, int want_spec)
{
  glb_gram_args g;
  memset(&g, 0, sizeof g);
  const int zc = g_zero_copy && !e->d_lmp;       /* LMP engines keep the PSD on the device for their ring */
  if (zc) {
    if (frame != e->h_frame) memcpy(e->h_frame, frame, sizeof(float) * e->n);
  } else {
    MUST(glb_memcpy_h2d(e->d_frame, frame, sizeof(float) * e->n, e->stream), "h2d");
  }
  g.n = e->n;
  g.hop = e->n;                      /* the frame is complete: no gather, no history */
  g.samples = zc ? e->h_frame : e->d_frame;
  g.general_only = zc;
  g.origin = 0;
  g.count = e->n;
  g.tapers = e->d_tapers;
  g.ntapers = e->ntapers;
  g.ra9mb_a = a;
  g.limiter = limiter;
  g.taper_scale = e->taper_scale;
  g.first_frame = 0;
  g.nframes = 1;
  g.rows = zc ? e->h_psd : e->d_psd;
  g.row_stride = e->n / 2 + 1;
  g.spectrum = want_spec ? (zc ? e->h_spec : e->d_spec) : NULL;
  g.tables = e->tables;
  MUST(glb_launch_gram(&g, e->stream), "launch");
  if (!zc) {
    MUST(glb_memcpy_d2h(e->h_psd, e->d_psd, sizeof(float) * (e->n / 2 + 1), e->stream), "d2h");
    if (want_spec) MUST(glb_memcpy_d2h(e->h_spec, e->d_spec, sizeof(float) * 2 * (e->n / 2 + 1), e->stream), "d2h");
  }
  MUST(glb_stream_sync(e->stream), "sync");
  e->fresh = 1;
}

/* ------------------------------------------------------------------ fft.h */
void compute_window(fft_params_t *params)
{
  glb_window_table(params->n, params->window_type, params->window);
}

/* fft.c:66-165, host-visible part: mean removal in the caller's block, history shift,
 * append, and the windowed copy in inbuf_fft */
void prepare_audio(float *audio_buf, fft_params_t *params)
{
  const int N = params->n;
  const int n_eff = glb_hop(N, params->overlap);
  const int n_overlap = N - n_eff;
  if (params->sub_mean) {
    float sig_mean = 0.0f;
    for (int i = 0; i < n_eff; i++) sig_mean += audio_buf[i];
    sig_mean /= n_eff;
    for (int i = 0; i < n_eff; i++) audio_buf[i] -= sig_mean;
  }
  if (!glb_first_buffer())
    memmove(params->inbuf_audio, params->inbuf_audio + (N - n_overlap), sizeof(glfer_real) * n_overlap);
  else
    memset(params->inbuf_audio, 0, sizeof(glfer_real) * n_overlap);
  for (int i = 0; i < n_eff; i++) params->inbuf_audio[i + n_overlap] = audio_buf[i];
  /* fft.c:127-156 with the reference's own types: inp_val and ftmp are floats, the buffers glfer_real */
  const int windowed = params->window_type != RECTANGULAR_WINDOW;
  for (int i = 0; i < N; i++) {
    if (params->a > 0.0) {
      const float inp_val = params->inbuf_audio[i];
      params->inbuf_fft[i] = inp_val / (params->a + inp_val * inp_val);
      if (windowed) params->inbuf_fft[i] *= params->window[i];
    } else if (windowed) {
      params->inbuf_fft[i] = params->window[i] * params->inbuf_audio[i];
    } else {
      params->inbuf_fft[i] = params->inbuf_audio[i];
    }
    if (params->limiter == 1) {
      const float ftmp = log(fabs(params->inbuf_fft[i]));
      params->inbuf_fft[i] = (params->inbuf_fft[i] > 0 ? exp(ftmp * 0.1) : -exp(ftmp * 0.1));
    }
  }
}

/* the three buffers of an estimator (fft.c:171-180, mtm.c:96-106, lmp.c:66-75) */
static void alloc_buffers(fft_params_t *fp, const char *where)
{
  const int n = fp->n;
  fp->inbuf_audio = calloc(n, sizeof(glfer_real));
  fp->inbuf_fft = calloc(n, sizeof(glfer_real));
#ifdef GLFER_FFTW_LAYOUT
  fp->plan = NULL;
  fp->outbuf = calloc(n, sizeof(glfer_real));
#else
  fp->outbuf = fp->inbuf_fft;
#endif
  if (!fp->inbuf_audio || !fp->inbuf_fft || !fp->outbuf) { glb_set_error("out of memory"); glb_fatal(where); }
}

static void free_buffers(fft_params_t *fp)
{
  free(fp->inbuf_audio); fp->inbuf_audio = NULL;
#ifdef GLFER_FFTW_LAYOUT
  free(fp->outbuf);
#endif
  free(fp->inbuf_fft); fp->inbuf_fft = NULL;
  fp->outbuf = NULL;
}

/* the frame as the device wants it: floats (the samples ARE floats, whatever the buffer type) */
static const float *frame_as_float(engine *e, const fft_params_t *fp)
{
#ifdef GLFER_FFTW_LAYOUT
  for (int i = 0; i < fp->n; i++) e->h_frame[i] = (float) fp->inbuf_audio[i];
  return e->h_frame;
#else
  (void) e;
  return fp->inbuf_audio;
#endif
}

void fft_init(fft_params_t *params)
{
  const int n = params->n;
  alloc_buffers(params, "fft_init");
  params->window = malloc(n * sizeof(float));
  if (!params->window) { glb_set_error("out of memory"); glb_fatal("fft_init"); }
  compute_window(params);
  params->sub_mean = glb_autoscale();            /* fft.c:186 */
  const float scale = (float) (1.0 / (2.0 * sqrt((double) n)));
  float *scaled = malloc(sizeof(float) * n);
  for (int i = 0; i < n; i++) {
    const double w = (params->window_type == RECTANGULAR_WINDOW) ? 1.0 : (double) params->window[i];
    scaled[i] = (float) (w * (double) scale);
  }
  make_estimator(params, n, 1, scaled, scale);
  free(scaled);
}

/* half-complex layout of fft_real_radix2_transform / rfftw_one: out[k] = Re, out[n-k] = Im */
static void spectrum_to_halfcomplex(const float *spec, int n, glfer_real *hc)
{
  hc[0] = spec[0];
  for (int k = 1; k < (n + 1) / 2; k++) {
    hc[k] = spec[2 * k];
    hc[n - k] = spec[2 * k + 1];
  }
  if (n % 2 == 0) hc[n / 2] = spec[2 * (n / 2)];
}

void fft_do(float *audio_buf, fft_params_t *params)
{
  engine *e = find_engine(params);
  if (!e) { glb_set_error("fft_do on parameters that did not go through fft_init"); glb_fatal("fft_do"); }
  prepare_audio(audio_buf, params);
  run_frame(e, frame_as_float(e, params), params->a, params->limiter, 1);
  spectrum_to_halfcomplex(e->h_spec, params->n, params->outbuf);   /* overwrites inbuf_fft, as in-place FFT does */
}

void fft_psd(float *psd_buf, float *phase_buf, fft_params_t *params)
{
  const int n = params->n, bins = n / 2 + 1;
  engine *e = find_engine(params);
  if (e && e->fresh && !phase_buf) {
    if (psd_buf) memcpy(psd_buf, e->h_psd, sizeof(float) * bins);
    return;
  }
  /* spectrum that was not produced by fft_do on these params (callers that fill outbuf
     themselves), or phase wanted: run fft_psd's formula on the device from outbuf */
  need_device("fft_psd");
  /* device scratch lives with the engine of these params (a shared one for foreign params): no
     cudaMalloc / cudaFree per call */
  engine *se = e ? e : scratch_engine();
  float *d_hc = engine_scratch(se, (size_t) n + 2 * (size_t) bins);
  float *d_psd = d_hc + n, *d_ph = d_psd + bins;
#ifdef GLFER_FFTW_LAYOUT
  float *hc32 = malloc(sizeof(float) * n);
  if (!hc32) { glb_set_error("out of memory"); glb_fatal("fft_psd"); }
  for (int i = 0; i < n; i++) hc32[i] = (float) params->outbuf[i];
  MUST(glb_memcpy_h2d(d_hc, hc32, sizeof(float) * n, se->stream), "h2d");
  MUST(glb_stream_sync(se->stream), "sync");
  free(hc32);
#else
  MUST(glb_memcpy_h2d(d_hc, params->outbuf, sizeof(float) * n, se->stream), "h2d");
#endif
  MUST(glb_launch_halfcomplex_psd(d_hc, n, psd_buf ? d_psd : NULL, phase_buf ? d_ph : NULL, se->stream), "launch");
  if (psd_buf) {
    if (e && e->fresh) memcpy(psd_buf, e->h_psd, sizeof(float) * bins);
    else MUST(glb_memcpy_d2h(psd_buf, d_psd, sizeof(float) * bins, se->stream), "d2h");
  }
  if (phase_buf) MUST(glb_memcpy_d2h(phase_buf, d_ph, sizeof(float) * bins, se->stream), "d2h");
  MUST(glb_stream_sync(se->stream), "sync");
}

/* k hop blocks per call (include/fft.h): exactly the sequence
 *     for b in 0..nblocks-1: fft_do(blocks + b * hop, params); fft_psd(rows + b * bins, NULL, params); first_buffer = FALSE
 * with ONE upload, ONE kernel launch and ONE download for the first nblocks - 1 blocks: the launch + copy
 * latency of the per-block calls (~40 us, against 7 us of CPU time for an N = 1024 frame) is paid once per
 * call instead of once per block.  The last block goes through fft_do itself, so that everything a caller
 * can observe afterwards (inbuf_audio history, inbuf_fft, outbuf, the block means removed in place) is what
 * the sequence of single calls leaves behind. */
static void run_batch(engine *e, fft_params_t *fp, float *blocks, int nblocks, float *rows, float a, int limiter)
{
  const int N = fp->n, hop = glb_hop(N, fp->overlap), n_ov = N - hop, bins = N / 2 + 1;
  if (nblocks <= 0) return;
  /* block means, in place, as prepare_audio does (fft.c:86-96: float accumulator, in order) */
  if (fp->sub_mean) {
    for (int b = 0; b < nblocks; b++) {
      float *blk = blocks + (size_t) b * hop, m = 0.0f;
      for (int i = 0; i < hop; i++) m += blk[i];
      m /= hop;
      for (int i = 0; i < hop; i++) blk[i] -= m;
    }
  }
  const size_t ns = (size_t) n_ov + (size_t) nblocks * hop;
  if (ns > e->bstream_cap) {
    glb_free(e->d_bstream); glb_host_free(e->h_bstream);
    e->d_bstream = e->h_bstream = NULL;
    MUST(glb_malloc((void **) &e->d_bstream, sizeof(float) * (ns + 4)), "malloc");
    MUST(glb_host_alloc((void **) &e->h_bstream, sizeof(float) * ns), "host_alloc");
    e->bstream_cap = ns;
  }
  if ((size_t) nblocks * bins > e->brows_cap) {
    glb_free(e->d_brows); glb_host_free(e->h_brows);
    e->d_brows = e->h_brows = NULL;
    MUST(glb_malloc((void **) &e->d_brows, sizeof(float) * (size_t) nblocks * bins), "malloc");
    MUST(glb_host_alloc((void **) &e->h_brows, sizeof(float) * (size_t) nblocks * bins), "host_alloc");
    e->brows_cap = (size_t) nblocks * bins;
  }
  /* the stream the frames see: the overlap history (zeros when glfer.first_buffer, fft.c:99-108), then the blocks */
  if (glb_first_buffer()) memset(e->h_bstream, 0, sizeof(float) * n_ov);
  else for (int i = 0; i < n_ov; i++) e->h_bstream[i] = (float) fp->inbuf_audio[N - n_ov + i];
  memcpy(e->h_bstream + n_ov, blocks, sizeof(float) * (size_t) nblocks * hop);
  /* frame F0 + b covers stream samples [(F0 + b) hop - n_ov, (F0 + b + 1) hop) = buffer [b hop, b hop + N) */
  const long long F0 = (n_ov + hop - 1) / hop;
  glb_gram_args g;
  memset(&g, 0, sizeof g);
  MUST(glb_memcpy_h2d(e->d_bstream, e->h_bstream, sizeof(float) * ns, e->stream), "h2d");
  g.n = N;
  g.hop = hop;
  g.samples = e->d_bstream;
  g.origin = F0 * hop - n_ov;
  g.count = (long long) ns;
  g.tapers = e->d_tapers;
  g.ntapers = e->ntapers;
  g.ra9mb_a = a;
  g.limiter = limiter;
  g.taper_scale = e->taper_scale;
  g.first_frame = F0;
  g.nframes = nblocks;
  g.rows = e->d_brows;
  g.row_stride = bins;
  g.tables = e->tables;
  MUST(glb_launch_gram(&g, e->stream), "launch");
  MUST(glb_memcpy_d2h(e->h_brows, e->d_brows, sizeof(float) * (size_t) nblocks * bins, e->stream), "d2h");
  MUST(glb_stream_sync(e->stream), "sync");
  memcpy(rows, e->h_brows, sizeof(float) * (size_t) nblocks * bins);
  /* history as nblocks calls leave it: the last N samples of the stream */
  if (ns >= (size_t) N) for (int i = 0; i < N; i++) fp->inbuf_audio[i] = e->h_bstream[ns - N + i];
}

void fft_do_batch(float *audio_blocks, int nblocks, float *psd_rows, fft_params_t *params)
{
  engine *e = find_engine(params);
  if (!e) { glb_set_error("fft_do_batch on parameters that did not go through fft_init"); glb_fatal("fft_do_batch"); }
  if (nblocks <= 0) return;
  const int hop = glb_hop(params->n, params->overlap), bins = params->n / 2 + 1;
  if (nblocks > 1) {
    run_batch(e, params, audio_blocks, nblocks - 1, psd_rows, params->a, params->limiter);
    glfer_b200_set_first_buffer(0);
    if (&glfer != NULL) glfer.first_buffer = 0;
  }
  /* the last block through the single-block path: inbuf_fft / outbuf / the cached PSD are those of fft_do */
  fft_do(audio_blocks + (size_t) (nblocks - 1) * hop, params);
  fft_psd(psd_rows + (size_t) (nblocks - 1) * bins, NULL, params);
}

void mtm_do_batch(float *audio_blocks, int nblocks, float *psd_rows, mtm_params_t *params)
{
  engine *e = find_engine(params);
  if (!e) { glb_set_error("mtm_do_batch on parameters that did not go through mtm_init"); glb_fatal("mtm_do_batch"); }
  if (nblocks <= 0) return;
  const int hop = glb_hop(params->fft.n, params->fft.overlap), bins = params->fft.n / 2 + 1;
  if (nblocks > 1) {
    run_batch(e, &params->fft, audio_blocks, nblocks - 1, psd_rows, 0.0f, 0);
    glfer_b200_set_first_buffer(0);
    if (&glfer != NULL) glfer.first_buffer = 0;
  }
  mtm_do(audio_blocks + (size_t) (nblocks - 1) * hop, psd_rows + (size_t) (nblocks - 1) * bins, NULL, params);
}

void fft_close(fft_params_t *params)
{
  drop_engine(params);
  free_buffers(params);
  free(params->window); params->window = NULL;
}

void compute_floor(float *psd_buf, int n, float *sig_pwr_p, float *floor_pwr_p, float *peak_pwr_p,
                   unsigned int *peak_bin_p)
{
  need_device("compute_floor");
  /* called once per frame by main_window_draw (g_main.c:1109): scratch kept for the life of the process */
  engine *se = scratch_engine();
  float st[4];
  float *d_row = engine_scratch(se, (size_t) n + 4), *d_st = d_row + n;
  MUST(glb_memcpy_h2d(d_row, psd_buf, sizeof(float) * n, se->stream), "h2d");
  MUST(glb_launch_floor_stats(d_row, n, n, 1, d_st, se->stream), "launch");
  MUST(glb_memcpy_d2h(st, d_st, sizeof st, se->stream), "d2h");
  MUST(glb_stream_sync(se->stream), "sync");
  *sig_pwr_p = st[0];
  *floor_pwr_p = st[1];
  *peak_pwr_p = st[2];
  *peak_bin_p = (unsigned int) st[3];
}

/* ------------------------------------------------------------------ mtm.h */
static double **nr_dmatrix(long nrl, long nrh, long ncl, long nch)
{
  /* 1-offset row pointers over one contiguous block, as the caller of mtm.c:118 expects */
  const long nrow = nrh - nrl + 1, ncol = nch - ncl + 1;
  double **m = malloc(sizeof(double *) * (nrow + 1));
  double *blk = malloc(sizeof(double) * (nrow * ncol + 1));
  if (!m || !blk) { glb_set_error("out of memory"); glb_fatal("mtm_init"); }
  m += 1;
  m -= nrl;
  m[nrl] = blk + 1 - ncl;
  for (long i = nrl + 1; i <= nrh; i++) m[i] = m[i - 1] + ncol;
  return m;
}

static void nr_free_dmatrix(double **m, long nrl, long ncl)
{
  free(m[nrl] + ncl - 1);
  free(m + nrl - 1);
}

void mtm_init(mtm_params_t *params)
{
  const int n = params->fft.n, kmax = params->kmax;
  if (kmax < 0 || kmax > 31) { fprintf(stderr, "libglfer_b200: mtm kmax %d outside 0..31\n", kmax); exit(-1); }
  alloc_buffers(&params->fft, "mtm_init");       /* mtm.c:96-106 */
  params->fft.sub_mean = glb_autoscale();        /* mtm.c:111 */
  params->window = nr_dmatrix(1, n, 0, kmax);    /* mtm.c:118 */
  params->sig = malloc(sizeof(double) * (kmax + 1));
  double *tap = malloc(sizeof(double) * (size_t) (kmax + 1) * n);
  double *lam = malloc(sizeof(double) * (kmax + 1));
  if (!params->sig || !tap || !lam) {
    glb_set_error("out of memory");
    glb_fatal("mtm_init");
  }
  if (glb_dpss(n, (double) params->w, kmax, tap, lam) != 0) fprintf(stderr, "Error: DPSS computation failed\n");
  const float scale = (float) (1.0 / (2.0 * sqrt((double) n)));
  float *scaled = malloc(sizeof(float) * (size_t) (kmax + 1) * n);
  for (int k = 0; k <= kmax; k++) {
    params->sig[k] = lam[k] - 1.0;
    const double g = (double) scale / sqrt(fabs(lam[k]));
    for (int i = 0; i < n; i++) {
      params->window[i + 1][k] = tap[(size_t) k * n + i];
      scaled[(size_t) k * n + i] = (float) (tap[(size_t) k * n + i] * g);
    }
  }
  make_estimator(params, n, kmax + 1, scaled, scale);
  free(scaled); free(tap); free(lam);
}

void mtm_do(float *audio_buf, float *psd_buf, float *phase_buf, mtm_params_t *params)
{
  (void) phase_buf;                              /* never written by the reference either */
  engine *e = find_engine(params);
  if (!e) { glb_set_error("mtm_do on parameters that did not go through mtm_init"); glb_fatal("mtm_do"); }
  prepare_audio(audio_buf, &params->fft);
  run_frame(e, frame_as_float(e, &params->fft), 0.0f, 0, 0);
  memcpy(psd_buf, e->h_psd, sizeof(float) * (params->fft.n / 2 + 1));
}

void mtm_close(mtm_params_t *params)
{
  drop_engine(params);
  free_buffers(&params->fft);
  if (params->window) nr_free_dmatrix(params->window, 1, 0);
  params->window = NULL;
  free(params->sig); params->sig = NULL;
}

/* ------------------------------------------------------------------ lmp.h */
void lmp_init(lmp_params_t *params)
{
  const int n = params->fft.n, nl = params->avg;
  if (nl < 2) { fprintf(stderr, "libglfer_b200: lmp avg %d < 2 (lmp.c:145 divides by avg - 1)\n", nl); exit(-1); }
  alloc_buffers(&params->fft, "lmp_init");       /* lmp.c:66-75 */
  params->fft.sub_mean = glb_autoscale();        /* lmp.c:80 */
  const float scale = (float) (1.0 / (2.0 * sqrt((double) n)));
  float *scaled = malloc(sizeof(float) * n);
  for (int i = 0; i < n; i++) scaled[i] = scale;          /* the frame goes into the FFT as it is (lmp.c:112-114) */
  engine *e = make_estimator(params, n, 1, scaled, scale);
  free(scaled);
  e->width = n / 2 + 1;
  e->depth = nl;
  e->frames = 0;
  MUST(glb_malloc((void **) &e->d_ring, sizeof(float) * (size_t) nl * e->width), "malloc");
  MUST(glb_malloc((void **) &e->d_lmp, sizeof(float) * e->width), "malloc");
  MUST(glb_memset(e->d_ring, 0, sizeof(float) * (size_t) nl * e->width, e->stream), "memset");   /* lmp.c:88-93 */
}

void lmp_do(float *audio_buf, float *psd_buf, float *phase_buf, lmp_params_t *params)
{
  (void) phase_buf;                              /* never written by the reference either */
  engine *e = find_engine(params);
  if (!e || !e->d_lmp) { glb_set_error("lmp_do on parameters that did not go through lmp_init"); glb_fatal("lmp_do"); }
  const int n = params->fft.n, bins = n / 2 + 1;
  prepare_audio(audio_buf, &params->fft);
  run_frame(e, frame_as_float(e, &params->fft), 0.0f, 0, 1);
  spectrum_to_halfcomplex(e->h_spec, n, params->fft.outbuf);       /* the in-place FFT of lmp.c:119 */
  const int slot = (int) (e->frames % e->depth);                   /* j_l, lmp.c:124 */
  MUST(glb_memcpy_d2d(e->d_ring + (size_t) slot * bins, e->d_psd, sizeof(float) * bins, e->stream), "d2d");
  MUST(glb_launch_lmp(e->d_ring, 0, bins, e->depth, bins, e->frames, 1, e->depth, 0, e->d_lmp, bins, e->stream), "launch");
  MUST(glb_memcpy_d2h(e->h_psd, e->d_lmp, sizeof(float) * bins, e->stream), "d2h");
  MUST(glb_stream_sync(e->stream), "sync");
  memcpy(psd_buf, e->h_psd, sizeof(float) * bins);
  e->fresh = 0;                                  /* h_psd no longer holds the PSD of the frame */
  e->frames++;
}

void lmp_close(lmp_params_t *params)
{
  drop_engine(params);
  free_buffers(&params->fft);
}

/* ------------------------------------------------------------------ avg.h */
void init_avg(avg_data_t *avgdata)
{
  avgdata->avgwidth = 0;
  avgdata->avgdepth = 0;
  avgdata->effdepth = 0;
}

void alloc_avg(avg_data_t *avgdata, int width, int depth)
{
  avgdata->avgwidth = width;
  avgdata->avgdepth = depth;
  avgdata->avgarray = malloc(width * sizeof(double *));
  for (int i = 0; i < width; i++) avgdata->avgarray[i] = calloc(depth > 0 ? depth : 1, sizeof(double));
  avgdata->avg = calloc(width > 0 ? width : 1, sizeof(double));
  avgdata->cum = calloc(width > 0 ? width : 1, sizeof(double));
  avgdata->effdepth = 0;
  need_device("alloc_avg");
  drop_engine(avgdata);
  engine *e = new_engine(avgdata);
  e->width = width;
  e->depth = depth;
  e->frames = 0;
  MUST(glb_stream_create(&e->stream), "stream");
  MUST(glb_malloc((void **) &e->d_ring, sizeof(float) * (size_t) (depth > 0 ? depth : 1) * width), "malloc");
  MUST(glb_malloc((void **) &e->d_avg, sizeof(double) * width), "malloc");
  MUST(glb_malloc((void **) &e->d_ret, sizeof(double)), "malloc");
  MUST(glb_malloc((void **) &e->d_var, sizeof(double)), "malloc");
  MUST(glb_malloc((void **) &e->d_cand, sizeof(int)), "malloc");
}

void delete_avg(avg_data_t *avgdata)
{
  if (avgdata->avgwidth != 0) {
    for (int i = 0; i < avgdata->avgwidth; i++) free(avgdata->avgarray[i]);
    free(avgdata->avgarray);
    free(avgdata->avg);
    free(avgdata->cum);
    drop_engine(avgdata);
  }
  avgdata->avgwidth = 0;
  avgdata->avgdepth = 0;
  avgdata->effdepth = 0;
}

static double update_avg_common(int mode, avg_data_t *ad, int N, float *psd, int max0, int minbin, int maxbin,
                                int *peakbin, double *variance)
{
  engine *e = find_engine(ad);
  if (!e) { glb_set_error("update_avg on data that did not go through alloc_avg"); glb_fatal("update_avg"); }
  const int width = e->width;
  if (minbin < 0 || maxbin > width || maxbin < minbin || N > width) {
    glb_set_error("update_avg: band or N outside the allocated width");
    glb_fatal("update_avg");
  }
  /* the reference reads psd[minbin .. maxbin) of the caller's row: upload that span into
     the ring row of this frame */
  float *ring_row = e->d_ring + (size_t) (e->frames % e->depth) * width;
  if (maxbin > minbin)
    MUST(glb_memcpy_h2d(ring_row + minbin, psd + minbin, sizeof(float) * (maxbin - minbin), e->stream), "h2d");
  else
    MUST(glb_memcpy_h2d(ring_row + minbin, psd + minbin, sizeof(float), e->stream), "h2d");
  glb_avg_args a;
  memset(&a, 0, sizeof a);
  a.mode = mode;
  a.depth = e->depth;
  a.minbin = minbin;
  a.maxbin = maxbin;
  a.max0 = max0;
  a.nbins = N;
  a.psd = e->d_ring;
  a.psd_first_frame = 0;
  a.psd_stride = width;
  a.psd_ring_rows = e->depth;
  a.first_frame = e->frames;
  a.nframes = 1;
  a.out_double = 1;
  a.avg_rows = e->d_avg;
  a.out_stride = width;
  a.ret = e->d_ret;
  a.peak_cand = e->d_cand;
  a.variance = e->d_var;
  a.peakbin_init = *peakbin;
  a.sequential = 1;
  MUST(glb_launch_avg(&a, e->stream), "launch");
  double ret = 0.0, var = 0.0;
  int cand = -1;
  MUST(glb_memcpy_d2h(ad->avg, e->d_avg, sizeof(double) * N, e->stream), "d2h");
  MUST(glb_memcpy_d2h(&ret, e->d_ret, sizeof(double), e->stream), "d2h");
  MUST(glb_memcpy_d2h(&var, e->d_var, sizeof(double), e->stream), "d2h");
  MUST(glb_memcpy_d2h(&cand, e->d_cand, sizeof(int), e->stream), "d2h");
  MUST(glb_stream_sync(e->stream), "sync");
  if (cand >= 0) *peakbin = cand;
  if (variance) *variance = var;
  e->frames++;
  if (ad->effdepth < ad->avgdepth) ad->effdepth++;
  return ret;
}

double update_avg_plain(avg_data_t *avgdata, int N, float *psd, int minbin, int maxbin, int *peakbin)
{
  return update_avg_common(GLFER_AVG_PLAIN, avgdata, N, psd, 0, minbin, maxbin, peakbin, NULL);
}

double update_avg_sumextreme(avg_data_t *avgdata, int N, float *psd, int max0, int minbin, int maxbin, int *peakbin)
{
  return update_avg_common(GLFER_AVG_SUMEXTREME, avgdata, N, psd, max0, minbin, maxbin, peakbin, NULL);
}

double update_avg_sumavg(avg_data_t *avgdata, int N, float *psd, int max0, int minbin, int maxbin, int *peakbin,
                         double *variance)
{
  return update_avg_common(GLFER_AVG_SUMAVG, avgdata, N, psd, max0, minbin, maxbin, peakbin, variance);
}
