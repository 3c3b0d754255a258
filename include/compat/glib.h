/*
 * include/compat/glib.h -- minimal stand-in for the GLib symbols that appear in
 * the cl_ops public API (GError out-params, gchar, GQuark).
 *
 * The reference's API reports errors through GLib's GError (e.g.
 * src/cl_ops/sort/clo_sort_abstract.in.h:116-120, src/cl_ops/common/_g_err_macros.h:61-96).
 * GLib is not vendored by the reference and is absent from this image, so the
 * library is built against this header.  The struct layout equals GLib's
 * (domain, code, message), so a client compiled against the real <glib.h> can
 * read the errors; it must free them with clo_b200_error_free() (or link a
 * libglib whose allocator is malloc-compatible).
 */
#ifndef CLO_B200_COMPAT_GLIB_H
#define CLO_B200_COMPAT_GLIB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef char gchar;
typedef int gint;
typedef unsigned int guint;
typedef uint32_t guint32;
typedef uint32_t GQuark;
typedef int gboolean;
typedef void* gpointer;

typedef struct _GError {
	GQuark domain;
	gint code;
	gchar* message;
} GError;

/* Subset of the GLib functions a caller of the cl_ops API needs. */
GQuark g_quark_from_static_string(const gchar* string);
const gchar* g_quark_to_string(GQuark quark);
void g_error_free(GError* error);
void g_clear_error(GError** err);
void g_set_error(GError** err, GQuark domain, gint code, const gchar* format, ...);
void g_propagate_error(GError** dest, GError* src);

#ifdef __cplusplus
}
#endif
#endif
