/* glfer_b200.h -- batched spectrogram extension of the estimator interface.
 *
 * The reference computes one frame per call (fft_do / mtm_do / update_avg_*, called from
 * audio_available(), source.c:130-165).  A GPU launch per hop block is latency bound, so
 * the library adds a batched entry point that a file source or a headless harness calls
 * once per recording (or per time shard).  Its results equal the sequence of per-call
 * results of the reference loop:
 *     for each hop block f:  fft_do; fft_psd   (or mtm_do);   [update_avg_*]
 * with glfer.first_buffer TRUE on block 0 only: frame f covers stream samples
 * [f*hop - (N-hop), f*hop + hop) with zeros before the stream start (fft.c:98-113), block
 * means removed per hop block when sub_mean (fft.c:86-96), averaging warm-up and the
 * effdepth+1 divisor counted from frame 0 (avg.c:116-156).
 *
 * All functions return 0 on success or a negative GLFER_E* code; glfer_b200_last_error()
 * returns the message.  There is no CPU fallback: without a CUDA device plan creation
 * fails with GLFER_ENODEV.
 */
#ifndef GLFER_B200_H
#define GLFER_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLFER_OK 0
#define GLFER_ECUDA (-1)
#define GLFER_EINVAL (-2)
#define GLFER_ENOMEM (-3)
#define GLFER_ENODEV (-4)

/* estimator mode: values of the reference's mode enum (glfer.h:47) */
#define GLFER_MODE_FFT 0
#define GLFER_MODE_MTM 1
#define GLFER_MODE_LMP 3   /* MODE_LMP: rectangular periodogram ring -> per-bin detector statistic (lmp.c) */
/* averaging mode: values of avgmode_t (glfer.h:53-55) */
#define GLFER_NO_AVG 0
#define GLFER_AVG_SUMAVG 1
#define GLFER_AVG_PLAIN 2
#define GLFER_AVG_SUMEXTREME 3

typedef struct {
  int mode;            /* GLFER_MODE_FFT | GLFER_MODE_MTM | GLFER_MODE_LMP */
  int n;               /* FFT size (opt.data_block_size): power of two, 32..32768 */
  int window_type;     /* fft.h window enum; ignored for MTM (forced rectangular, source.c:344) */
  float overlap;       /* opt.data_blocks_overlap; hop = (int)(n * (1.0 - overlap)) (fft.c:70) */
  float a;             /* opt.limiter_a, RA9MB parameter; <= 0 disables */
  int limiter;         /* opt.enable_limiter */
  int sub_mean;        /* opt.autoscale as latched by fft_init (fft.c:186) */
  float mtm_w;         /* opt.mtm_w (N*W) */
  int mtm_kmax;        /* opt.mtm_k: kmax + 1 tapers are used */
  int avg_mode;        /* GLFER_NO_AVG ... */
  int avg_depth;       /* opt.avgsamples */
  int avg_minbin;      /* (int)(opt.min_avgband / binsize) (g_main.c:1145) */
  int avg_maxbin;
  int avg_max0;        /* scale_type is *_MAX0 (g_main.c:1148-1151) */
  int avg_peakbin_init;/* the caller's peakbin before frame 0 */
  int scale_db;        /* 1: rows are 10*log10(value) (g_main.c:1191-1199) */
  int device;          /* CUDA device ordinal */
  int lmp_av;          /* opt.lmp_av: rows in the LMP ring (>= 2); the "psd" rows of an LMP plan are the
                          statistic of lmp_do; window forced rectangular (source.c:395), RA9MB / limiter
                          do not reach the spectrum (lmp.c:112-114); no frame averaging in this mode */
  int avg_band_only;   /* 1: averaged rows hold the band only, [nframes][avg_maxbin - avg_minbin] (column j = bin
                          avg_minbin + j).  The reference fills the rest of avg[] with the constant 1e-15
                          (avg.c:152-153); writing and shipping that constant is most of the cost of the averaging
                          pass when the band is a few dozen bins out of thousands */
  int mtm_ftest;       /* 1 (MTM plans): also prepare Thomson's harmonic F-test, which mtm_do computes into a file-static
                          nothing reads (mtm.c:59,165-174,204-210,222-233); rows through glfer_gram_run_mtm_ftest() */
  int zero_history;    /* 1: glfer.first_buffer stays TRUE, i.e. prepare_audio zeroes the N-hop history on EVERY
                          frame (fft.c:99-108).  That is what the reference GUI does with opt.autoscale == 0:
                          first_buffer is cleared only inside `if (opt.autoscale)` (g_main.c:1111-1120).
                          0 (default): TRUE on block 0 only, the sequence a sane caller drives */
} glfer_gram_config;

typedef struct glfer_gram_plan glfer_gram_plan;

const char *glfer_b200_last_error(void);
int glfer_b200_device_count(void);
/* pinned host memory for fast transfers (optional; any host pointer is accepted) */
void *glfer_b200_host_alloc(size_t bytes);
void glfer_b200_host_free(void *p);
/* kernels launched by this process so far */
unsigned long long glfer_b200_kernel_launches(void);

void glfer_gram_config_default(glfer_gram_config *cfg);     /* the defaults of glfer.c:238-279 */
int glfer_gram_plan_create(const glfer_gram_config *cfg, glfer_gram_plan **plan);
void glfer_gram_plan_destroy(glfer_gram_plan *plan);

int glfer_gram_hop(const glfer_gram_plan *plan);
int glfer_gram_bins(const glfer_gram_plan *plan);                       /* n/2 + 1 */
int glfer_gram_avg_cols(const glfer_gram_plan *plan);                   /* columns of an averaged row: bins, or the band (avg_band_only) */
long long glfer_gram_num_frames(const glfer_gram_plan *plan, long long nsamples);  /* nsamples / hop */
/* stream span [lo, hi) the frames [first_frame, first_frame + nframes) read (lo may be
 * negative: indices < 0 are the zero history) */
void glfer_gram_required_span(const glfer_gram_plan *plan, long long first_frame, long long nframes,
                              long long *lo, long long *hi);
/* window / taper tables as the estimator uses them (unit energy, before internal scaling) */
int glfer_gram_window(const glfer_gram_plan *plan, float *window /* [n] */);
int glfer_gram_tapers(const glfer_gram_plan *plan, double *tapers /* [kmax+1][n] */, double *lambda /* [kmax+1] */);

/* ---- one call, host buffers in and out (transfers pipelined with the kernels) ----
 * samples points at stream index `origin`, `count` samples long, and must cover
 * glfer_gram_required_span() clipped to [0, inf).  Any output pointer may be NULL.
 *   psd_rows [nframes][bins]   PSD (or dB when scale_db)
 *   avg_rows [nframes][bins]   averaged rows (avg_mode != GLFER_NO_AVG)
 *   avg_ret / avg_peakbin / avg_variance [nframes]: return value, *peakbin, *variance of
 *   update_avg_* after each frame. */
int glfer_gram_run(glfer_gram_plan *plan, const float *samples, long long origin, long long count,
                   long long first_frame, long long nframes, float *psd_rows, float *avg_rows,
                   double *avg_ret, int *avg_peakbin, double *avg_variance);
/* same, 16-bit PCM input converted on the device as wav_fmt.c:113 does */
int glfer_gram_run_pcm16(glfer_gram_plan *plan, const short *pcm, long long origin, long long count,
                         long long first_frame, long long nframes, float *psd_rows, float *avg_rows,
                         double *avg_ret, int *avg_peakbin, double *avg_variance);

/* Multitaper rows plus the harmonic F-test of every frame (plan created with mtm_ftest = 1):
 *   ftest[f][i] = kmax |mu(i)|^2 sum_U0_sqr / sum_j |y_j(i) - mu(i) U0_j|^2        (mtm.c:204-233)
 * with mu = FFT(frame * hn), hn = sum_j U0_j v_j / sum_U0_sqr, U0_j = sum_i v_j[i] (mtm.c:78-84,125-136).  As in the
 * reference the DC bin uses real parts only and, for even n, ftest[f][n/2] = num / 0 (inf, or nan for a zero
 * numerator).  Either output may be NULL. */
int glfer_gram_run_mtm_ftest(glfer_gram_plan *plan, const float *samples, long long origin, long long count,
                             long long first_frame, long long nframes, float *psd_rows, float *ftest_rows);

/* ---- device-resident path: stage once, execute many times, fetch when wanted ---- */
int glfer_gram_stage(glfer_gram_plan *plan, const float *samples, long long origin, long long count);
int glfer_gram_stage_pcm16(glfer_gram_plan *plan, const short *pcm, long long origin, long long count);
/* frames [first_frame, first_frame + nframes) from the staged samples into device rows;
 * returns after the work is queued.  kernel_ms (may be NULL) receives the device time of
 * the launch sequence measured with CUDA events on the plan's stream (this call then
 * waits for completion). */
int glfer_gram_exec(glfer_gram_plan *plan, long long first_frame, long long nframes, float *kernel_ms);
/* device time of the fused spectrogram kernel alone in the last glfer_gram_exec that was
 * given a non-NULL kernel_ms (the other launches are block means / averaging) */
int glfer_gram_last_gram_ms(glfer_gram_plan *plan, float *ms);
int glfer_gram_sync(glfer_gram_plan *plan);
int glfer_gram_fetch(glfer_gram_plan *plan, float *psd_rows, float *avg_rows, double *avg_ret,
                     int *avg_peakbin, double *avg_variance);

/* ---- display mapping: what main_window_draw does with a row (g_main.c:1109-1229) ----
 * compute_floor statistics -> AGC of the display range (autoscale) or fixed dB levels ->
 * level = 10*log10 (through the reference's `short` level buffer: integer dB) or linear ->
 * 8-bit palette index with threshold and clipping, pixel i = bin bins-1-i -> optional RGB.
 * The rows shown are the averaged rows when the plan averages, else the PSD rows. */
typedef struct {
  int log_scale;          /* opt.scale_type is SCALE_LOG or SCALE_LOG_MAX0 (g_main.c:1132,1191) */
  int autoscale;          /* opt.autoscale (g_main.c:1111) */
  float max_level_db;     /* opt.max_level_db / opt.min_level_db, used when !autoscale (g_main.c:1126-1128) */
  float min_level_db;
  float thr_level;        /* opt.thr_level, percent (g_main.c:1098) */
  const unsigned char *colortab;   /* 256 RGB triplets (set_palette, g_main.c:651-762) or NULL */
} glfer_display_config;
/* levels [nframes][bins] and/or rgb [nframes][bins][3]; display_range [nframes][2] (max, min)
 * when autoscale (optional).  agc_state [2] = (display_max_lvl, display_min_lvl): in when
 * first_frame > 0, out always (lets time shards chain the recurrence); may be NULL for runs
 * that start at frame 0. */
int glfer_gram_run_display(glfer_gram_plan *plan, const float *samples, long long origin, long long count,
                           long long first_frame, long long nframes, const glfer_display_config *dc,
                           float *agc_state, unsigned char *levels, unsigned char *rgb, float *display_range);
int glfer_gram_run_display_pcm16(glfer_gram_plan *plan, const short *pcm, long long origin, long long count,
                                 long long first_frame, long long nframes, const glfer_display_config *dc,
                                 float *agc_state, unsigned char *levels, unsigned char *rgb, float *display_range);

/* The level mapping alone on caller-provided rows (host memory): rows [nrows][nbins] -> levels [nrows][nbins],
 * pixel i of a row = bin nbins-1-i.  range: [nrows][2] = (display_max, display_min) per row as main_window_draw
 * holds them (dB in the log scales), or NULL for the fixed levels of dc (dc->autoscale is ignored). */
int glfer_b200_map_levels(const float *rows, long long nrows, int nbins, const glfer_display_config *dc,
                          const float *range, unsigned char *levels, int device);
/* per-call interface (fft_do / mtm_do / lmp_do): 1 (default) = the frame is read from, and the results are written
 * to, pinned host memory by the kernel itself (no copy-engine round trips); 0 = staged through device buffers */
void glfer_b200_set_zero_copy(int on);
/* 0 (default) = frame averaging as a second pass over the PSD rows; 1 = fused into the spectrogram kernel where it
 * can be (avg_band_only plans, N = 4096 / 8192, regular overlap, band history <= 2 KB): bit-identical rows, but
 * measured slower on B200 (1.31 ms against 0.82 + 0.20 ms on BASELINE config C2), so it is opt-in */
void glfer_b200_set_fused_avg(int on);
/* testing aid: 0 = glfer_gram_run_display always maps the levels in a second pass over float rows; 1 (default) =
 * with a fixed display range and no averaging the spectrogram kernel writes the 8-bit levels itself */
void glfer_b200_set_fused_levels(int on);
/* the palettes of set_palette (g_main.c:649-762; glfer.h:47: 0 HSV, 1 THRESH, 2 COOL, 3 HOT, 4 BW, 5 BONE,
 * 6 COPPER, 7 OTD): 256 RGB triplets into tab[768] */
int glfer_palette(int palette, unsigned char *tab);

/* ---- time-sharded multi-GPU run inside one process (one host thread per device) ----
 * Frames [0, nframes) are split into ndev contiguous ranges; device g gets samples
 * [F_g*hop - halo, F_{g+1}*hop) (halo = N - hop, plus (depth-1) frames when averaging);
 * no inter-GPU communication.  devices == NULL means 0..ndev-1. */
int glfer_gram_run_sharded(const glfer_gram_config *cfg, int ndev, const int *devices, const float *samples,
                           long long nsamples, float *psd_rows, float *avg_rows, double *avg_ret,
                           int *avg_peakbin, double *avg_variance);
/* the shard arithmetic on its own: frame range of shard g of ndev */
void glfer_gram_shard_range(long long nframes, int ndev, int g, long long *first, long long *count);

/* ---- WAV source (wav_fmt.c:45-121 semantics, LP64-safe) ---- */
typedef struct {
  int sample_rate;
  int bits;             /* 8 or 16 */
  int channels;
  long long nsamples;   /* samples in the data chunk (channels not de-interleaved, as the reference) */
  void *data;           /* raw PCM as read */
} glfer_wav;
int glfer_wav_load(const char *path, glfer_wav *wav);
/* extension (the reference treats a stereo file's interleaved samples as one channel): keep one channel, in place */
int glfer_wav_select_channel(glfer_wav *wav, int channel);
void glfer_wav_free(glfer_wav *wav);
/* spectrogram of a WAV exactly as `glfer -f file` would see it block by block, including
 * the stale tail of a short final read (wav_fmt.c:102-119).  Returns frames via *nframes;
 * psd_rows must hold glfer_wav_num_frames() rows. */
long long glfer_wav_num_frames(const glfer_gram_plan *plan, const glfer_wav *wav);
int glfer_gram_run_wav(glfer_gram_plan *plan, const glfer_wav *wav, float *psd_rows, float *avg_rows,
                       double *avg_ret, int *avg_peakbin, double *avg_variance);

/* ---- hidden inputs of the per-call interface (fft.c:99,186: globals `glfer`, `opt`) ----
 * When the host program defines `opt` and `glfer` (as glfer.c:56-62 does) the per-call
 * functions read them.  Otherwise these setters provide the two values. */
void glfer_b200_set_autoscale(int autoscale);
void glfer_b200_set_first_buffer(int first_buffer);

#ifdef __cplusplus
}
#endif
#endif
