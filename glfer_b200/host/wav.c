/* wav.c -- WAV file source for the batched path.
 *
 * Restates the reference's reader (wav_fmt.c:45-121) for LP64: its header struct uses
 * u_long (wav_fmt.h:36-50) and mis-parses the 44-byte canonical header on x86-64, so the
 * fields are read here at their byte offsets.  Semantics kept: the canonical 44-byte
 * header is skipped whatever chunks follow (wav_fmt.c:58-66), everything after it is
 * audio (trailing RIFF chunks included), one hop block per read (:87,102), 8-bit samples
 * map to (x-128)/128 and 16-bit ones to x/32768 (:105-116) with channels left
 * interleaved, and a short final read leaves the previous block's tail in place (:102-119;
 * the conversion buffer is allocated once, zero-filled, :91-100).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "glb_host.h"

static int wfail(const char *msg)
{
  glb_set_error(msg);
  return GLFER_EINVAL;
}

int glfer_wav_load(const char *path, glfer_wav *wav)
{
  memset(wav, 0, sizeof *wav);
  FILE *fh = fopen(path, "rb");
  if (!fh) return wfail("cannot open WAV file");
  unsigned char hd[44];
  if (fread(hd, 1, 44, fh) != 44) { fclose(fh); return wfail("input file shorter than a WAV header"); }
  if (memcmp(hd, "RIFF", 4) != 0) { fclose(fh); return wfail("input file not in WAV format"); }
  const unsigned format = hd[20] | (hd[21] << 8);
  if (format != 1) fprintf(stderr, "input is not a PCM WAV file");          /* wav_fmt.c:67-68: warn, go on */
  wav->channels = hd[22] | (hd[23] << 8);
  wav->sample_rate = (int) (hd[24] | (hd[25] << 8) | (hd[26] << 16) | ((unsigned) hd[27] << 24));
  wav->bits = hd[34] | (hd[35] << 8);
  if (wav->bits != 8 && wav->bits != 16) { fclose(fh); return wfail("only 8 and 16 bit PCM is supported"); }
  fseek(fh, 0, SEEK_END);
  const long end = ftell(fh);
  fseek(fh, 44, SEEK_SET);
  const size_t bytes = end > 44 ? (size_t) (end - 44) : 0;
  wav->data = malloc(bytes ? bytes : 1);
  if (!wav->data) { fclose(fh); glb_set_error("out of memory"); return GLFER_ENOMEM; }
  if (fread(wav->data, 1, bytes, fh) != bytes) { fclose(fh); free(wav->data); wav->data = NULL; return wfail("short read"); }
  fclose(fh);
  wav->nsamples = (long long) (bytes / (wav->bits / 8));
  return 0;
}

/* An extension the reference lacks (it feeds the interleaved samples of a stereo file to the estimator as if they
   were one channel): keep ONE channel of the loaded file, in place.  channel in [0, channels); a trailing partial
   sample frame is dropped.  Afterwards the file looks mono to glfer_gram_run_wav(). */
int glfer_wav_select_channel(glfer_wav *wav, int channel)
{
  if (!wav || !wav->data) return wfail("no WAV data");
  if (wav->channels < 1 || channel < 0 || channel >= wav->channels) return wfail("no such channel in the WAV file");
  if (wav->channels == 1) return 0;
  const long long nfr = wav->nsamples / wav->channels;
  if (wav->bits == 16) {
    short *b = wav->data;
    for (long long i = 0; i < nfr; i++) b[i] = b[i * wav->channels + channel];
  } else {
    unsigned char *b = wav->data;
    for (long long i = 0; i < nfr; i++) b[i] = b[i * wav->channels + channel];
  }
  wav->nsamples = nfr;
  wav->channels = 1;
  return 0;
}

void glfer_wav_free(glfer_wav *wav)
{
  free(wav->data);
  wav->data = NULL;
}

long long glfer_wav_num_frames(const glfer_gram_plan *plan, const glfer_wav *wav)
{
  /* one frame per non-empty read (wav_fmt.c:119): a read that returns only an odd byte of a
     16-bit stream still counts */
  const long long hop = glfer_gram_hop(plan);
  const long long bps = wav->bits / 8;
  const long long bytes = wav->nsamples * bps;
  return (bytes + hop * bps - 1) / (hop * bps);
}

/* the sub_mean flag of a plan (gram.c) */
int glb_plan_sub_mean(const glfer_gram_plan *plan);

int glfer_gram_run_wav(glfer_gram_plan *plan, const glfer_wav *wav, float *psd_rows, float *avg_rows,
                       double *avg_ret, int *avg_peakbin, double *avg_variance)
{
  const long long hop = glfer_gram_hop(plan);
  const long long nframes = glfer_wav_num_frames(plan, wav);
  if (nframes == 0) return 0;
  const long long total = nframes * hop;
  /* The stream as the block-by-block reader presents it, as floats: 8-bit (x-128)/128 and
     16-bit x/32768 exactly (wav_fmt.c:108,113).  A short final read leaves the previous
     block's tail in the reader's buffer (:102-119) -- and with sub_mean that buffer has already
     had the previous block's mean subtracted in place by prepare_audio (fft.c:93-95), so the
     tail is padded with the mean-removed values. */
  const int partial = (wav->nsamples % hop) != 0;
  if (wav->bits == 16 && !(partial && glb_plan_sub_mean(plan) && nframes >= 2)) {
    /* common case: ship the PCM and convert on the device */
    short *pcm = malloc(sizeof(short) * (size_t) total);
    if (!pcm) { glb_set_error("out of memory"); return GLFER_ENOMEM; }
    memcpy(pcm, wav->data, sizeof(short) * (size_t) wav->nsamples);
    for (long long i = wav->nsamples; i < total; i++) pcm[i] = (i >= hop) ? pcm[i - hop] : 0;
    const int rc = glfer_gram_run_pcm16(plan, pcm, 0, total, 0, nframes, psd_rows, avg_rows, avg_ret, avg_peakbin,
                                        avg_variance);
    free(pcm);
    return rc;
  }
  float *x = malloc(sizeof(float) * (size_t) total);
  if (!x) { glb_set_error("out of memory"); return GLFER_ENOMEM; }
  if (wav->bits == 16) {
    const short *b = wav->data;
    for (long long i = 0; i < wav->nsamples; i++) x[i] = (float) b[i] / 32768;
  } else {
    const unsigned char *b = wav->data;
    for (long long i = 0; i < wav->nsamples; i++) x[i] = ((float) b[i] - 128) / 128;
  }
  float prev_mean = 0.0f;
  if (partial && glb_plan_sub_mean(plan) && nframes >= 2) {
    const float *pb = x + (nframes - 2) * hop;
    for (long long i = 0; i < hop; i++) prev_mean += pb[i];
    prev_mean /= hop;
  }
  for (long long i = wav->nsamples; i < total; i++) x[i] = (i >= hop) ? x[i - hop] - prev_mean : 0.0f;
  const int rc = glfer_gram_run(plan, x, 0, total, 0, nframes, psd_rows, avg_rows, avg_ret, avg_peakbin, avg_variance);
  free(x);
  return rc;
}
