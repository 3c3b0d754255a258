/*
 * compat_glib.cu -- the GLib functions the reference's own benchmark drivers call
 * (/root/reference/src/benchmarks/clo_bench.c:67-142, clo_sort_bench.c:129-145,
 * clo_scan_bench.c), so that those drivers link unchanged against this library.  GLib is an
 * un-vendored dependency of the reference and absent from this image; the semantics follow
 * GLib's published documentation.  Host-side support code only: nothing here is on the
 * device path.
 */
#include "clo_internal.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <strings.h>
#include <time.h>

extern "C" {

gpointer g_malloc(gsize n) { return n ? malloc(n) : NULL; }
gpointer g_malloc0(gsize n) { return n ? calloc(1, n) : NULL; }
void g_free(gpointer p) { free(p); }
gpointer g_slice_alloc(gsize n) { return g_malloc(n); }
void g_slice_free1(gsize n, gpointer p) { (void) n; free(p); }

gchar* g_strdup(const gchar* s) {
	if (!s) return NULL;
	const size_t n = strlen(s) + 1;
	gchar* r = (gchar*) malloc(n);
	memcpy(r, s, n);
	return r;
}

gchar* g_strconcat(const gchar* first, ...) {
	if (!first) return NULL;
	size_t len = strlen(first);
	va_list ap;
	va_start(ap, first);
	for (const gchar* s = va_arg(ap, const gchar*); s; s = va_arg(ap, const gchar*)) len += strlen(s);
	va_end(ap);
	gchar* r = (gchar*) malloc(len + 1);
	strcpy(r, first);
	va_start(ap, first);
	for (const gchar* s = va_arg(ap, const gchar*); s; s = va_arg(ap, const gchar*)) strcat(r, s);
	va_end(ap);
	return r;
}

gint g_strcmp0(const gchar* a, const gchar* b) {
	if (!a) return b ? -1 : 0;
	if (!b) return 1;
	return strcmp(a, b);
}

gboolean g_str_has_prefix(const gchar* s, const gchar* prefix) {
	return s && prefix && strncmp(s, prefix, strlen(prefix)) == 0;
}

gint g_ascii_strncasecmp(const gchar* a, const gchar* b, gsize n) { return strncasecmp(a, b, n); }

static GPrintFunc g_print_handler = NULL;

GPrintFunc g_set_print_handler(GPrintFunc func) {
	GPrintFunc old = g_print_handler;
	g_print_handler = func;
	return old;
}

void g_print(const gchar* format, ...) {
	va_list ap;
	va_start(ap, format);
	if (g_print_handler) {
		char buf[4096];
		vsnprintf(buf, sizeof(buf), format, ap);
		g_print_handler(buf);
	} else {
		vfprintf(stdout, format, ap);
	}
	va_end(ap);
}

/* ---- test harness */
struct clo_gtest { const char* path; GTestFunc fn; };
static clo_gtest g_tests[64];
static int g_ntests = 0;

void g_test_init(int* argc, char*** argv, ...) { (void) argc; (void) argv; }

void g_test_add_func(const char* testpath, GTestFunc test_func) {
	if (g_ntests < 64) { g_tests[g_ntests].path = testpath; g_tests[g_ntests].fn = test_func; ++g_ntests; }
}

int g_test_run(void) {
	for (int i = 0; i < g_ntests; ++i) {
		printf("%s: ", g_tests[i].path);
		fflush(stdout);
		g_tests[i].fn();             /* a failing g_assert aborts, as in GLib */
		printf("OK\n");
	}
	return 0;
}

void clo_b200_g_debug(const gchar* format, ...) {
	if (!getenv("G_MESSAGES_DEBUG")) return;
	va_list ap;
	va_start(ap, format);
	fprintf(stderr, "** (debug) ");
	vfprintf(stderr, format, ap);
	fprintf(stderr, "\n");
	va_end(ap);
}

void clo_b200_g_assert_fail(const char* expr, const char* loc) {
	fprintf(stderr, "**\nERROR:%s: assertion failed: (%s)\n", loc, expr);
	abort();
}

/* ---- GRand: MT19937 as GLib seeds and draws it */
struct _GRand {
	guint32 mt[624];
	guint mti;
};

GRand* g_rand_new_with_seed(guint32 seed) {
	GRand* r = (GRand*) malloc(sizeof(GRand));
	r->mt[0] = seed;
	for (r->mti = 1; r->mti < 624; r->mti++)
		r->mt[r->mti] = 1812433253u * (r->mt[r->mti - 1] ^ (r->mt[r->mti - 1] >> 30)) + r->mti;
	return r;
}

void g_rand_free(GRand* r) { free(r); }

guint32 g_rand_int(GRand* r) {
	static const guint32 mag01[2] = { 0x0u, 0x9908b0dfu };
	guint32 y;
	if (r->mti >= 624) {
		int kk;
		for (kk = 0; kk < 624 - 397; kk++) {
			y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
			r->mt[kk] = r->mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1];
		}
		for (; kk < 623; kk++) {
			y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
			r->mt[kk] = r->mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1];
		}
		y = (r->mt[623] & 0x80000000u) | (r->mt[0] & 0x7fffffffu);
		r->mt[623] = r->mt[396] ^ (y >> 1) ^ mag01[y & 1];
		r->mti = 0;
	}
	y = r->mt[r->mti++];
	y ^= (y >> 11);
	y ^= (y << 7) & 0x9d2c5680u;
	y ^= (y << 15) & 0xefc60000u;
	y ^= (y >> 18);
	return y;
}

gint32 g_rand_int_range(GRand* r, gint32 begin, gint32 end) {
	const guint32 dist = (guint32) end - (guint32) begin;
	guint32 v = 0;
	if (dist != 0) {
		guint32 maxvalue;
		if (dist <= 0x80000000u) {
			guint32 leftover = (0x80000000u % dist) * 2;
			if (leftover >= dist) leftover -= dist;
			maxvalue = 0xffffffffu - leftover;
		} else {
			maxvalue = dist - 1;
		}
		do v = g_rand_int(r); while (v > maxvalue);
		v %= dist;
	}
	return (gint32) ((guint32) begin + v);
}

gdouble g_rand_double(GRand* r) {
	const double T = 2.3283064365386962890625e-10;   /* 2^-32 */
	double v = g_rand_int(r) * T;
	v = (v + g_rand_int(r)) * T;
	if (v >= 1.0) return g_rand_double(r);
	return v;
}

gdouble g_rand_double_range(GRand* r, gdouble begin, gdouble end) {
	const gdouble v = g_rand_double(r);
	return v * end - (v - 1) * begin;
}

/* ---- GTimer */
struct _GTimer {
	struct timespec start, end;
	int active;
};

GTimer* g_timer_new(void) {
	GTimer* t = (GTimer*) malloc(sizeof(GTimer));
	clock_gettime(CLOCK_MONOTONIC, &t->start);
	t->active = 1;
	return t;
}

void g_timer_stop(GTimer* t) { clock_gettime(CLOCK_MONOTONIC, &t->end); t->active = 0; }

gdouble g_timer_elapsed(GTimer* t, gulong* microseconds) {
	struct timespec now = t->end;
	if (t->active) clock_gettime(CLOCK_MONOTONIC, &now);
	const double s = (double) (now.tv_sec - t->start.tv_sec) + 1e-9 * (double) (now.tv_nsec - t->start.tv_nsec);
	if (microseconds) *microseconds = (gulong) ((s - (double) (long) s) * 1e6);
	return s;
}

void g_timer_destroy(GTimer* t) { free(t); }

/* ---- GOption: "--name value", "--name=value", "-n value", flags, --help */
struct _GOptionContext {
	gchar* summary;
	const GOptionEntry* entries;
};

GOptionContext* g_option_context_new(const gchar* parameter_string) {
	GOptionContext* c = (GOptionContext*) calloc(1, sizeof(GOptionContext));
	c->summary = g_strdup(parameter_string ? parameter_string : "");
	return c;
}

void g_option_context_add_main_entries(GOptionContext* c, const GOptionEntry* entries, const gchar* domain) {
	(void) domain;
	c->entries = entries;
}

void g_option_context_free(GOptionContext* c) {
	if (!c) return;
	free(c->summary);
	free(c);
}

static const GOptionEntry* find_entry(const GOptionContext* c, const char* lname, size_t llen, char sname) {
	if (!c->entries) return NULL;
	for (const GOptionEntry* e = c->entries; e->long_name; ++e) {
		if (lname && strlen(e->long_name) == llen && strncmp(e->long_name, lname, llen) == 0) return e;
		if (!lname && sname && e->short_name == sname) return e;
	}
	return NULL;
}

gboolean g_option_context_parse(GOptionContext* c, gint* argc, gchar*** argv, GError** error) {
	const GQuark dom = g_quark_from_static_string("g-option-context-error-quark");
	if (!argc || !argv) return TRUE;
	int out = 1;
	for (int i = 1; i < *argc; ++i) {
		char* a = (*argv)[i];
		const GOptionEntry* e = NULL;
		const char* inline_val = NULL;
		if (strcmp(a, "--help") == 0 || strcmp(a, "-?") == 0 || (strcmp(a, "-h") == 0 && !find_entry(c, NULL, 0, 'h'))) {
			printf("Usage:\n  %s [OPTION...]%s\n\nApplication Options:\n", (*argv)[0], c->summary);
			if (c->entries)
				for (const GOptionEntry* q = c->entries; q->long_name; ++q)
					printf("  -%c, --%s%s%s\t%s\n", q->short_name, q->long_name, q->arg == G_OPTION_ARG_NONE ? "" : "=",
						q->arg == G_OPTION_ARG_NONE ? "" : (q->arg_description ? q->arg_description : ""),
						q->description ? q->description : "");
			exit(0);
		}
		if (a[0] == '-' && a[1] == '-' && a[2]) {
			const char* eq = strchr(a + 2, '=');
			e = find_entry(c, a + 2, eq ? (size_t) (eq - (a + 2)) : strlen(a + 2), 0);
			if (eq) inline_val = eq + 1;
		} else if (a[0] == '-' && a[1] && a[1] != '-') {
			e = find_entry(c, NULL, 0, a[1]);
			if (e && a[2]) inline_val = a + 2;
		} else {
			(*argv)[out++] = a;          /* not an option: left for the caller */
			continue;
		}
		if (!e) {
			g_set_error(error, dom, 1, "Unknown option %s", a);
			return FALSE;
		}
		if (e->arg == G_OPTION_ARG_NONE) {
			*(gboolean*) e->arg_data = TRUE;
			continue;
		}
		const char* val = inline_val;
		if (!val) {
			if (i + 1 >= *argc) {
				g_set_error(error, dom, 0, "Missing argument for %s", a);
				return FALSE;
			}
			val = (*argv)[++i];
		}
		if (e->arg == G_OPTION_ARG_INT) {
			char* end = NULL;
			const long v = strtol(val, &end, 0);
			if (!end || *end || end == val) {
				g_set_error(error, dom, 0, "Cannot parse integer value '%s' for %s", val, a);
				return FALSE;
			}
			*(gint*) e->arg_data = (gint) v;
		} else {
			free(*(gchar**) e->arg_data);
			*(gchar**) e->arg_data = g_strdup(val);
		}
	}
	*argc = out;
	return TRUE;
}

} /* extern "C" */
