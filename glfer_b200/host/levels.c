/* levels.c -- host side of the display mapping (main_window_draw, g_main.c:1186-1229, palettes
 * g_main.c:649-762): the tables that let the device reproduce the reference's arithmetic exactly.
 *
 * In the log scales the level of a bin goes through the GUI's `short` level buffer
 * (`sig_level = levbuf[..] = 10.0 * log10(x)`, g_main.c:68,1193-1195): it is the dB value truncated
 * towards zero, computed by the HOST's libm in double.  A device log10 may differ from glibc's in
 * the last place, which flips the truncation when 10 log10(x) sits on an integer.  So the library
 * never evaluates that expression on the device: it tabulates, with the host libm, the smallest
 * float (and double) whose level is >= j for every integer j a finite positive input can reach
 * (j = -450 .. 390), and the device finds the level by comparing against those thresholds -- a fast
 * estimate from lg2.approx decides every case that is not within 5e-4 dB of an integer, the table
 * decides the rest.  The result is the host expression bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "glb_host.h"

/* (short) (10.0 * log10(x)) as x86-64 evaluates it: cvttsd2si gives 0x80000000 for NaN / +-inf /
   out of range, whose low 16 bits are 0 (x <= 0, NaN and +inf all land there) */
int glb_short_db_d(double x)
{
  const double d = 10.0 * log10(x);
  if (!(fabs(d) < 2147483648.0)) return 0;
  return (short) (int) d;
}

int glb_short_db_f(float x) { return glb_short_db_d((double) x); }

static float f_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static double d_from_bits(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

/* thr[j - GLB_DB_JMIN] = the smallest positive finite float x with short_db(x) >= j (+inf when no
   float reaches j).  short_db is non-decreasing over the positive floats, which are ordered like
   their bit patterns: one bisection per j. */
void glb_db_thresholds_f(float *thr)
{
  for (int j = GLB_DB_JMIN; j <= GLB_DB_JMAX; j++) {
    uint32_t lo = 1u, hi = 0x7f7fffffu;             /* smallest subnormal .. FLT_MAX */
    if (glb_short_db_f(f_from_bits(hi)) < j) { thr[j - GLB_DB_JMIN] = INFINITY; continue; }
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if (glb_short_db_f(f_from_bits(mid)) >= j) hi = mid; else lo = mid + 1;
    }
    thr[j - GLB_DB_JMIN] = f_from_bits(lo);
  }
}

/* the same over the positive doubles whose level lies in the float range (averaged rows are doubles
   in the reference, avg.h:32, g_main.c:1193) */
void glb_db_thresholds_d(double *thr)
{
  for (int j = GLB_DB_JMIN; j <= GLB_DB_JMAX; j++) {
    uint64_t lo = 1ull, hi = 0x7fefffffffffffffull;
    while (lo < hi) {
      const uint64_t mid = lo + (hi - lo) / 2;
      if (glb_short_db_d(d_from_bits(mid)) >= j) hi = mid; else lo = mid + 1;
    }
    thr[j - GLB_DB_JMIN] = d_from_bits(lo);
  }
}

/* g_main.c:1206,1221-1229: f = 255 * ((sig_level - display_min) / (display_max - display_min)) in float;
   below threshold -> 0, above 255 -> 255, else (f - 255 thr) / (1 - thr) in double, truncated */
unsigned char glb_level_u8(float sig_level, float dmin, float dmax, float thr)
{
  const float f = 255 * ((sig_level - dmin) / (dmax - dmin));
  unsigned char v;
  if (f < 255.0 * thr) v = 0;
  else if (f > 255) v = 255;
  else v = (f - 255.0 * thr) / (1.0 - thr);
  return v;
}

/* fixed display range: the level of every integer dB value j */
void glb_level_lut(float dmin, float dmax, float thr, unsigned char *lut)
{
  for (int j = GLB_DB_JMIN; j <= GLB_DB_JMAX; j++) lut[j - GLB_DB_JMIN] = glb_level_u8((float) j, dmin, dmax, thr);
}

/* display range for fixed levels (autoscale off), g_main.c:1126-1135 */
void glb_fixed_display_range(float max_level_db, float min_level_db, int log_scale, float *dmax, float *dmin)
{
  float mx = pow(10.0, max_level_db / 10.0);
  float mn = pow(10.0, min_level_db / 10.0);
  mn = (mx > mn ? mn : mx / 10.0);                 /* "prevent stupid entries" */
  if (log_scale) {
    *dmax = 10.0 * log10(mx);
    *dmin = 10.0 * log10(mn);
  } else {
    *dmax = mx;
    *dmin = mn;
  }
}

/* The eight palettes of set_palette (g_main.c:649-762; enum glfer.h:47: HSV, THRESH, COOL, HOT, BW,
   BONE, COPPER, OTD), 256 RGB triplets.  Piecewise-linear ramps over the index c; the arithmetic
   types (long index, double products truncated to unsigned char) are the reference's. */
static unsigned char u8(double v) { return (unsigned char) v; }

int glfer_palette(int palette, unsigned char *tab)
{
  if (!tab || palette < 0 || palette > 7) return GLFER_EINVAL;
  for (long c = 0; c < 256; c++) {
    unsigned char *p = tab + 3 * c;
    const double x = (double) c;
    unsigned char r, g, b;
    switch (palette) {
    case 0: case 1:                                 /* HSV, thresholded HSV (black below 16) */
      if (palette == 1 && c < 16) { r = g = b = 0; }
      else if (c < 64) { r = 0; g = u8(x * 4.0); b = 255; }
      else if (c < 128) { r = 0; g = 255; b = u8(510.0 - x * 4.0); }
      else if (c < 192) { r = u8(x * 4.0 - 510.0); g = 255; b = 0; }
      else { r = 255; g = u8(1020.0 - x * 4.0); b = 0; }
      break;
    case 2:                                         /* cool */
      r = (unsigned char) c; g = (unsigned char) (255 - c); b = 255;
      break;
    case 3:                                         /* hot */
      if (c < 96) { r = u8(x * 2.66667 + 0.5); g = 0; b = 0; }
      else if (c < 192) { r = 255; g = u8(x * 2.66667 - 254); b = 0; }
      else { r = 255; g = 255; b = u8(x * 4.0 - 766.0); }
      break;
    case 5:                                         /* bone */
      if (c < 96) { r = u8(x * 0.88889); g = u8(x * 0.88889); b = u8(x * 1.20000); }
      else if (c < 192) { r = u8(x * 0.88889); g = u8(x * 1.20000 - 29); b = u8(x * 0.88889 + 29); }
      else { r = u8(x * 1.20000 - 60); g = u8(x * 0.88889 + 29); b = u8(x * 0.88889 + 29); }
      break;
    case 6:                                         /* copper */
      r = c < 208 ? u8(x * 1.23) : 255; g = u8(x * 0.78); b = u8(x * 0.5);
      break;
    case 7:                                         /* OTD */
      if (c < 128) { r = 0; g = u8(2.0 * x - 1.0); b = u8(2.0 * (127.0 - x) + 1.0); }
      else { r = u8(2.0 * (x - 127.0) - 1.0); g = u8(2.0 * (255.0 - x) + 1.0); b = 0; }
      break;
    default:                                        /* black and white */
      r = g = b = (unsigned char) c;
      break;
    }
    p[0] = r; p[1] = g; p[2] = b;
  }
  return GLFER_OK;
}
