/* Minimal stand-in for <gtk/gtk.h>, TEST INFRASTRUCTURE ONLY.
 *
 * The reference's glfer.h (glfer.h:25) includes <gtk/gtk.h> but the spectrum
 * estimator sources (fft.c, mtm.c, g-l_dpss.c, avg.c, util.c) only need the
 * type names that appear in opt_t / glfer_t (glfer.h:62-139).  GTK2 is not
 * installed in this image, so the oracle build puts this directory on the
 * include path.  Nothing here is used by the product library.
 */
#ifndef ORACLE_GTK_STUB_H
#define ORACLE_GTK_STUB_H
typedef char gchar;
typedef int gint;
typedef void *gpointer;
typedef struct oracle_gtk_widget GtkWidget;
typedef struct oracle_gtk_tooltips GtkTooltips;
#endif
