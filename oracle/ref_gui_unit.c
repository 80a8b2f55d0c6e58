/* The reference's GUI unit g_main.c, UNMODIFIED, as a translation unit of the oracle, so that
 * main_window_draw() -- compute_floor, the display AGC, the dB / linear level mapping through the
 * `short` level buffer, threshold + clipping to the 8-bit palette index, bin reversal and the
 * palette look-up (g_main.c:1072-1281, palettes :649-762) -- can be RUN here and pins the display
 * path of the product.  TEST INFRASTRUCTURE ONLY.
 *
 * g_main.c is #included (found through -I$(REF)) because everything the function touches is a
 * file-static: n, l, pixmap_height, levbuf, rgbbuf, colortab, drawing_area, spec_da and the AGC state
 * inside the function.  GTK is replaced by oracle/shim_gui/gtk/gtk.h (types only) and by do-nothing
 * stubs generated for every undefined symbol (gen_gui_stubs.py); the one GTK call that matters,
 * gdk_draw_rgb_image(), is defined below and captures the RGB line the reference drew.
 * The AGC state lives in function-statics that nothing resets; glfer.first_buffer = TRUE re-seeds
 * it exactly as the GUI does after a parameter change (g_main.c:313,990,1112-1120).
 */
#define PACKAGE_STRING "glfer (oracle build)"
#include "g_main.c"

avg_data_t avgdata;             /* glfer.c:62 */

static GtkStyle oracle_style;
static GtkWidget oracle_da, oracle_spec;
static GString oracle_gstring;
static char oracle_gstring_buf[8] = "";
static unsigned char *oracle_rgb_capture;
static int oracle_rgb_rows, oracle_last_x;

GString *g_string_new(const gchar *init)
{
  oracle_gstring.str = oracle_gstring_buf;
  return &oracle_gstring;
}

/* gdk_draw_rgb_image(pixmap, gc, x, 0, 1, n_zoom * n, GDK_RGB_DITHER_NONE, rgbbuf, 3), g_main.c:1253 */
int gdk_draw_rgb_image(void *drawable, void *gc, int x, int y, int width, int height, int dither,
                       unsigned char *rgb_buf, int rowstride)
{
  oracle_last_x = x;
  if (oracle_rgb_capture) {
    int i;
    for (i = 0; i < height && i < oracle_rgb_rows; i++) {
      oracle_rgb_capture[3 * i] = rgb_buf[rowstride * i];
      oracle_rgb_capture[3 * i + 1] = rgb_buf[rowstride * i + 1];
      oracle_rgb_capture[3 * i + 2] = rgb_buf[rowstride * i + 2];
    }
  }
  return 0;
}

/* what main_window_init / drawing_area_configure_event leave behind for a spectrum of nbins bins
 * (g_main.c:798,846,424-431), without creating widgets */
void refh_gui_setup(int nbins, int palette)
{
  n = nbins;
  n_zoom = 1;
  l = 600;
  pixmap_height = n;
  pixmap_width = l;
  free(rgbbuf);
  free(levbuf);
  rgbbuf = calloc(3 * n * n_zoom, sizeof(guchar));
  levbuf = calloc((size_t) pixmap_width * pixmap_height, sizeof(short));
  oracle_da.style = &oracle_style;
  oracle_spec.style = &oracle_style;
  drawing_area = &oracle_da;
  spec_da = &oracle_spec;
  set_palette(palette);
  glfer.first_buffer = TRUE;
}

void refh_gui_palette(int palette, unsigned char *tab /* [768] */)
{
  set_palette(palette);
  memcpy(tab, colortab, 768);
}

/* one call of main_window_draw on a PSD row.  rgb: [n][3] the line drawn (pixel i = bin n-1-i);
 * lev: [n] the `short` level column; scal[6]: sig_pwr, floor_pwr, peak_pwr, peak_bin, avgmax, peakfreq */
void refh_gui_draw(float *psd_row, unsigned char *rgb, short *lev, float *scal)
{
  int i;
  oracle_rgb_capture = rgb;
  oracle_rgb_rows = n;
  main_window_draw(psd_row);
  oracle_rgb_capture = NULL;
  if (lev)
    for (i = 0; i < n; i++) lev[i] = levbuf[oracle_last_x * pixmap_height + i];
  if (scal) {
    scal[0] = sig_pwr; scal[1] = floor_pwr; scal[2] = peak_pwr; scal[3] = (float) peak_bin;
    scal[4] = glfer.avgmax; scal[5] = glfer.peakfreq;
  }
}

/* display options the function reads from `opt` (g_main.c:1098,1111,1126-1146) */
void refh_gui_options(int scale_type, int autoscale, float max_level_db, float min_level_db, float thr_level,
                      float overlap, int averaging, int avgsamples, float min_avgband, float max_avgband,
                      int sample_rate, int data_block_size)
{
  opt.scale_type = scale_type;
  opt.autoscale = autoscale;
  opt.max_level_db = max_level_db;
  opt.min_level_db = min_level_db;
  opt.thr_level = thr_level;
  opt.data_blocks_overlap = overlap;
  opt.averaging = averaging;
  opt.avgsamples = avgsamples;
  opt.min_avgband = min_avgband;
  opt.max_avgband = max_avgband;
  opt.sample_rate = sample_rate;
  opt.data_block_size = data_block_size;
}

void refh_gui_first_buffer(int on) { glfer.first_buffer = on; }

/* (re)allocate the averaging state main_window_draw updates (source.c:312: width = block size) */
void refh_gui_alloc_avg(int width, int depth)
{
  static int have;
  if (have) delete_avg(&avgdata);
  else init_avg(&avgdata);
  alloc_avg(&avgdata, width, depth);
  have = 1;
  glfer.avgfill = 0;
}
