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
/* the real <glib.h> brings these in, and the reference's drivers rely on it */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

/* ---- what the reference's own drivers use on top of that (src/benchmarks/clo_*_bench.c,
 *      src/benchmarks/clo_bench.c, src/cl_ops/common/_g_err_macros.h), so that they compile
 *      and link UNCHANGED against this library (oracle/build_ref_drivers.py).  Semantics
 *      follow GLib's documentation; GRand is GLib's MT19937 (seeding of GLib >= 2.2,
 *      g_rand_double from two draws, g_rand_int_range with rejection), which is what makes
 *      the drivers' input data the same as with the real GLib. */
typedef unsigned char guchar;
typedef unsigned short gushort;
typedef long glong;
typedef unsigned long gulong;
typedef int32_t gint32;
typedef int64_t gint64;
typedef uint64_t guint64;
typedef double gdouble;
typedef float gfloat;
typedef size_t gsize;
typedef const void* gconstpointer;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#define G_MAXUSHORT 0xffffu
#define G_MAXUINT 0xffffffffu
#define G_MAXULONG (~0ul)
#define G_STRINGIFY(x) G_STRINGIFY_ARG(x)
#define G_STRINGIFY_ARG(x) #x
#define G_STRLOC __FILE__ ":" G_STRINGIFY(__LINE__)
#define G_STRFUNC ((const char*) (__func__))
#define G_GNUC_UNUSED __attribute__((unused))

gpointer g_malloc(gsize n);
gpointer g_malloc0(gsize n);
void g_free(gpointer p);
gchar* g_strdup(const gchar* s);
gchar* g_strconcat(const gchar* first, ...);
gint g_strcmp0(const gchar* a, const gchar* b);
gboolean g_str_has_prefix(const gchar* s, const gchar* prefix);
gint g_ascii_strncasecmp(const gchar* a, const gchar* b, gsize n);
gpointer g_slice_alloc(gsize n);
void g_slice_free1(gsize n, gpointer p);
typedef void (*GPrintFunc)(const gchar* string);
GPrintFunc g_set_print_handler(GPrintFunc func);     /* g_print goes through the handler when one is set */
void g_print(const gchar* format, ...);
void clo_b200_g_debug(const gchar* format, ...);      /* printed when G_MESSAGES_DEBUG is set */
void clo_b200_g_assert_fail(const char* expr, const char* loc);
#define g_new(type, n) ((type*) g_malloc(sizeof(type) * (gsize) (n)))
#define g_new0(type, n) ((type*) g_malloc0(sizeof(type) * (gsize) (n)))
#define g_debug(...) clo_b200_g_debug(__VA_ARGS__)
#define g_assert(expr) do { if (!(expr)) clo_b200_g_assert_fail(#expr, G_STRLOC); } while (0)
#define g_assert_not_reached() clo_b200_g_assert_fail("not reached", G_STRLOC)
#define g_assert_no_error(err) g_assert((err) == NULL)
#define g_return_if_fail(expr) do { if (!(expr)) return; } while (0)
#define g_return_val_if_fail(expr, val) do { if (!(expr)) return (val); } while (0)

typedef struct _GRand GRand;
GRand* g_rand_new_with_seed(guint32 seed);
void g_rand_free(GRand* r);
guint32 g_rand_int(GRand* r);
gint32 g_rand_int_range(GRand* r, gint32 begin, gint32 end);
gdouble g_rand_double(GRand* r);
gdouble g_rand_double_range(GRand* r, gdouble begin, gdouble end);
#define g_rand_boolean(r) ((g_rand_int(r) & (1 << 15)) != 0)

/* GLib's test harness, as far as src/tests/test_rng.c:444-462 uses it */
typedef void (*GTestFunc)(void);
void g_test_init(int* argc, char*** argv, ...);
void g_test_add_func(const char* testpath, GTestFunc test_func);
int g_test_run(void);

typedef struct _GTimer GTimer;
GTimer* g_timer_new(void);
void g_timer_stop(GTimer* t);
gdouble g_timer_elapsed(GTimer* t, gulong* microseconds);
void g_timer_destroy(GTimer* t);

typedef enum { G_OPTION_ARG_NONE, G_OPTION_ARG_STRING, G_OPTION_ARG_INT } GOptionArg;
typedef struct _GOptionEntry {
	const gchar* long_name;
	gchar short_name;
	gint flags;
	GOptionArg arg;
	gpointer arg_data;
	const gchar* description;
	const gchar* arg_description;
} GOptionEntry;
typedef struct _GOptionContext GOptionContext;
GOptionContext* g_option_context_new(const gchar* parameter_string);
void g_option_context_add_main_entries(GOptionContext* c, const GOptionEntry* entries, const gchar* domain);
gboolean g_option_context_parse(GOptionContext* c, gint* argc, gchar*** argv, GError** error);
void g_option_context_free(GOptionContext* c);

#ifdef __cplusplus
}
#endif
#endif
