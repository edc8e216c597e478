/*
 * ref_runtime.c -- see ref_runtime.h.  TEST INFRASTRUCTURE (oracle/), never linked into the product.
 */
#include "ref_runtime.h"

#include <time.h>

/* ================================================================= arrays */
static long guard_elems_for(int rank, const long *ext)
{
    /* reads up to 4 planes outside a rank>=3 array happen in the reference (see header) */
    if (rank >= 3) return 4 * ext[0] * ext[1] + 64;
    return 64;
}

static char *guarded_alloc(long nelem, size_t elsz, long guard, char **raw)
{
    size_t gb = (size_t)guard * elsz, body = (size_t)(nelem > 0 ? nelem : 0) * elsz;
    char *r = (char *)malloc(gb * 2 + body + 16);
    if (!r) { fprintf(stderr, "ref_runtime: out of memory (%zu bytes)\n", gb * 2 + body); abort(); }
    memset(r, 0xFF, gb);
    memset(r + gb + body, 0xFF, gb + 16);
    /* ALLOCATE leaves the contents undefined; poison them too so a read-before-write shows up */
    memset(r + gb, 0xFF, body);
    *raw = r;
    return r + gb;
}

void f_allocate(f_desc *d, int rank, size_t elsz, const long *lb, const long *ub)
{
    if (d->base) { fprintf(stderr, "ref_runtime: ALLOCATE of an allocated array\n"); abort(); }
    long n = 1;
    d->rank = rank;
    d->elsz = elsz;
    for (int r = 0; r < rank; ++r) {
        d->lb[r] = lb[r];
        d->ext[r] = ub[r] - lb[r] + 1;
        if (d->ext[r] < 0) d->ext[r] = 0;
        n *= d->ext[r];
    }
    d->base = guarded_alloc(n, elsz, guard_elems_for(rank, d->ext), &d->raw);
}

void f_deallocate(f_desc *d)
{
    if (d->raw) free(d->raw);
    d->raw = d->base = NULL;
}

long f_size(const f_desc *d)
{
    long n = 1;
    for (int r = 0; r < d->rank; ++r) n *= d->ext[r];
    return n;
}

void f_assign_alloc(f_desc *dst, const f_desc *src)
{
    int same = dst->base && dst->rank == src->rank;
    for (int r = 0; same && r < src->rank; ++r) same = dst->ext[r] == src->ext[r];
    if (!same) {
        long lb[4], ub[4];
        f_deallocate(dst);
        for (int r = 0; r < src->rank; ++r) { lb[r] = src->lb[r]; ub[r] = src->lb[r] + src->ext[r] - 1; }
        f_allocate(dst, src->rank, src->elsz, lb, ub);
    }
    memcpy(dst->base, src->base, (size_t)f_size(src) * src->elsz);
}

void *f_auto(long nelem, size_t elsz, long guard)
{
    char *raw;
    return guarded_alloc(nelem, elsz, guard, &raw);
}

void f_auto_free(void *p, size_t elsz, long guard)
{
    if (p) free((char *)p - (size_t)guard * elsz);
}

/* ================================================================= strings */
#define ARENA_BYTES (1 << 16)
static char arena[ARENA_BYTES];
static size_t arena_pos;

static char *arena_get(long n)
{
    if (n < 0) n = 0;
    if ((size_t)n + 1 > ARENA_BYTES) { fprintf(stderr, "ref_runtime: string too long\n"); abort(); }
    if (arena_pos + (size_t)n + 1 > ARENA_BYTES) arena_pos = 0;   /* ring: temporaries live for one statement */
    char *p = arena + arena_pos;
    arena_pos += (size_t)n + 1;
    return p;
}

fstr f_lit(const char *s, long n) { fstr r = { s, n }; return r; }
fstr f_var(const char *p, long n) { fstr r = { p, n }; return r; }

fstr f_concat(fstr a, fstr b)
{
    char *p = arena_get(a.n + b.n);
    memcpy(p, a.p, (size_t)a.n);
    memcpy(p + a.n, b.p, (size_t)b.n);
    fstr r = { p, a.n + b.n };
    return r;
}

fstr f_substr(fstr a, long lo, long hi)
{
    fstr r = { a.p + (lo - 1), hi - lo + 1 };
    if (r.n < 0) r.n = 0;
    return r;
}

int f_len_trim(fstr a)
{
    long n = a.n;
    while (n > 0 && a.p[n - 1] == ' ') --n;
    return (int)n;
}

fstr f_trim(fstr a) { fstr r = { a.p, f_len_trim(a) }; return r; }

fstr f_char(int code)
{
    char *p = arena_get(1);
    p[0] = (char)code;
    fstr r = { p, 1 };
    return r;
}

int f_str_eq(fstr a, fstr b)
{
    long n = a.n > b.n ? a.n : b.n;
    for (long i = 0; i < n; ++i) {
        char ca = i < a.n ? a.p[i] : ' ', cb = i < b.n ? b.p[i] : ' ';
        if (ca != cb) return 0;
    }
    return 1;
}

void f_str_assign(char *dst, long n, fstr src)
{
    /* the source may overlap the destination (meshname = filename(1:k)//endname does not, but be safe) */
    char *tmp = arena_get(n);
    for (long i = 0; i < n; ++i) tmp[i] = i < src.n ? src.p[i] : ' ';
    memcpy(dst, tmp, (size_t)n);
}

/* ================================================================= STOP */
jmp_buf ref_stop_jmp;
int ref_stop_armed = 0;

void f_stop(void)
{
    if (ref_stop_armed) longjmp(ref_stop_jmp, 1);
    fprintf(stderr, "ref_runtime: STOP outside a guarded call\n");
    exit(0);
}

/* ================================================================= PRINT capture */
static ref_print_item *pr_items;
static long pr_n, pr_cap, pr_limit = 4000000;
static int pr_line, pr_echo;

static void pr_push(int kind, long i, double r, fstr s)
{
    if (pr_n >= pr_limit) return;
    if (pr_n == pr_cap) {
        pr_cap = pr_cap ? pr_cap * 2 : 1024;
        pr_items = (ref_print_item *)realloc(pr_items, (size_t)pr_cap * sizeof *pr_items);
    }
    ref_print_item *it = &pr_items[pr_n++];
    it->line = pr_line; it->kind = kind; it->i = i; it->r = r;
    long n = s.n < (long)sizeof it->s - 1 ? s.n : (long)sizeof it->s - 1;
    if (n > 0) memcpy(it->s, s.p, (size_t)n);
    it->s[n > 0 ? n : 0] = 0;
}

static const fstr no_str = { "", 0 };
void f_pr_begin(int line) { pr_line = line; if (pr_echo) fputc(' ', stdout); }
void f_pr_s(fstr s) { pr_push(0, 0, 0., s); if (pr_echo) fwrite(s.p, 1, (size_t)s.n, stdout); }
void f_pr_i(long v) { pr_push(1, v, 0., no_str); if (pr_echo) printf(" %11ld", v); }
void f_pr_r(double v) { pr_push(2, 0, v, no_str); if (pr_echo) printf(" %25.17g", v); }
void f_pr_end(void) { pr_push(3, 0, 0., no_str); if (pr_echo) { fputc('\n', stdout); fflush(stdout); } }
long ref_print_count(void) { return pr_n; }
const ref_print_item *ref_print_get(long idx) { return (idx >= 0 && idx < pr_n) ? &pr_items[idx] : NULL; }
void ref_print_clear(void) { pr_n = 0; }
void ref_print_echo(int on) { pr_echo = on; }
void ref_print_limit(long m) { pr_limit = m; }

/* ================================================================= files */
#define MAX_UNIT 64
static FILE *units[MAX_UNIT];
static char outdir[1024];

void f_set_outdir(const char *dir) { snprintf(outdir, sizeof outdir, "%s", dir ? dir : ""); }

static FILE *unit_file(int unit)
{
    if (unit < 0 || unit >= MAX_UNIT || !units[unit]) { fprintf(stderr, "ref_runtime: unit %d not open\n", unit); abort(); }
    return units[unit];
}

void f_open(int unit, fstr file, fstr status, fstr access, fstr form)
{
    (void)access; (void)form;
    char name[2048], path[3100];
    int n = f_len_trim(file);
    if (n >= (int)sizeof name) n = sizeof name - 1;
    memcpy(name, file.p, (size_t)n);
    name[n] = 0;
    int old = f_str_eq(status, f_lit("old", 3)) || f_str_eq(status, f_lit("OLD", 3));
    if (name[0] != '/' && outdir[0] && !old) snprintf(path, sizeof path, "%s/%s", outdir, name);
    else snprintf(path, sizeof path, "%s", name);
    if (unit < 0 || unit >= MAX_UNIT) { fprintf(stderr, "ref_runtime: bad unit %d\n", unit); abort(); }
    if (units[unit]) fclose(units[unit]);
    units[unit] = fopen(path, old ? "rb" : "wb");
    if (!units[unit]) { fprintf(stderr, "ref_runtime: cannot open '%s'\n", path); f_stop(); }
    static char big[MAX_UNIT][1];
    (void)big;
    setvbuf(units[unit], NULL, _IOFBF, 1 << 20);
}

void f_close(int unit)
{
    if (unit >= 0 && unit < MAX_UNIT && units[unit]) { fclose(units[unit]); units[unit] = NULL; }
}

void f_read(int unit, void *dst, size_t nbytes)
{
    if (fread(dst, 1, nbytes, unit_file(unit)) != nbytes) { fprintf(stderr, "ref_runtime: read past end of file\n"); f_stop(); }
}

void f_write(int unit, const void *src, size_t nbytes) { fwrite(src, 1, nbytes, unit_file(unit)); }
void f_write_s(int unit, fstr s) { fwrite(s.p, 1, (size_t)s.n, unit_file(unit)); }

/* ---- explicit formats: A[w], Iw, Fw.d, repeat counts and one level of nested groups ------------ */
typedef struct { char kind; int w, d; } edit;
static edit edits[256];
static int n_edits, cur_edit;
static char fmt_out[4096];
static long fmt_len;

static const char *parse_group(const char *p, const char *end, int depth)
{
    while (p < end) {
        while (p < end && (*p == ' ' || *p == ',')) ++p;
        if (p >= end) break;
        if (*p == ')') return p + 1;
        int rep = 0, has_rep = 0;
        while (p < end && *p >= '0' && *p <= '9') { rep = rep * 10 + (*p - '0'); ++p; has_rep = 1; }
        if (!has_rep) rep = 1;
        if (*p == '(') {
            int start = n_edits;
            p = parse_group(p + 1, end, depth + 1);
            int len = n_edits - start;
            for (int r = 1; r < rep; ++r)
                for (int q = 0; q < len; ++q) edits[n_edits++] = edits[start + q];
            continue;
        }
        char k = *p++;
        if (k >= 'a' && k <= 'z') k = (char)(k - 32);
        edit e = { k, 0, 0 };
        int has_w = 0;
        while (p < end && *p >= '0' && *p <= '9') { e.w = e.w * 10 + (*p - '0'); ++p; has_w = 1; }
        if (!has_w) e.w = -1;
        if (p < end && *p == '.') { ++p; while (p < end && *p >= '0' && *p <= '9') { e.d = e.d * 10 + (*p - '0'); ++p; } }
        if (k != 'A' && k != 'I' && k != 'F') { fprintf(stderr, "ref_runtime: edit descriptor %c not supported\n", k); abort(); }
        for (int r = 0; r < rep; ++r) edits[n_edits++] = e;
    }
    return p;
}

void f_fmt_begin(fstr fmt)
{
    n_edits = cur_edit = 0;
    fmt_len = 0;
    const char *p = fmt.p, *end = fmt.p + fmt.n;
    while (p < end && *p != '(') ++p;
    parse_group(p + 1, end, 0);
}

static edit next_edit(char want)
{
    if (n_edits == 0) { fprintf(stderr, "ref_runtime: empty format\n"); abort(); }
    if (cur_edit >= n_edits) { cur_edit = 0; fmt_out[fmt_len++] = '\n'; }   /* format reversion */
    edit e = edits[cur_edit++];
    if (e.kind != want) { fprintf(stderr, "ref_runtime: format/item mismatch (%c vs %c)\n", e.kind, want); abort(); }
    return e;
}

static void fmt_put(const char *s, long n, int w)
{
    if (w < 0) w = (int)n;
    if (n >= w) { memcpy(fmt_out + fmt_len, s, (size_t)w); }             /* A editing truncates on the right */
    else { memset(fmt_out + fmt_len, ' ', (size_t)(w - n)); memcpy(fmt_out + fmt_len + (w - n), s, (size_t)n); }
    fmt_len += w;
}

void f_fmt_s(fstr s) { edit e = next_edit('A'); fmt_put(s.p, s.n, e.w); }

void f_fmt_i(long v)
{
    edit e = next_edit('I');
    char b[64];
    int n = snprintf(b, sizeof b, "%ld", v);
    if (n > e.w) { memset(fmt_out + fmt_len, '*', (size_t)e.w); fmt_len += e.w; }
    else fmt_put(b, n, e.w);
}

void f_fmt_r(double v)
{
    edit e = next_edit('F');
    char b[512];
    int n = snprintf(b, sizeof b, "%.*f", e.d, v);
    if (n > e.w && b[0] == '0') { memmove(b, b + 1, (size_t)n); --n; }   /* optional leading zero */
    if (n > e.w) { memset(fmt_out + fmt_len, '*', (size_t)e.w); fmt_len += e.w; }
    else fmt_put(b, n, e.w);
}

void f_fmt_end_internal(char *dst, long n)
{
    for (long i = 0; i < n; ++i) dst[i] = i < fmt_len ? fmt_out[i] : ' ';
}

void f_fmt_end_unit(int unit)
{
    FILE *f = unit == -2 ? stdout : unit_file(unit);
    fwrite(fmt_out, 1, (size_t)fmt_len, f);
    fputc('\n', f);
}

/* ---- list-directed output (libgfortran spacing, see header) ------------------------------------- */
static FILE *ld_file;
static int ld_first, ld_prev_char;

void f_ld_begin(int unit, int line) { ld_file = unit_file(unit); ld_first = 1; ld_prev_char = 0; pr_line = line; }

static void ld_sep(int is_char)
{
    if (ld_first) { fputc(' ', ld_file); ld_first = 0; }
    else if (!(is_char && ld_prev_char)) fputc(' ', ld_file);
    ld_prev_char = is_char;
}

void f_ld_s(fstr s) { ld_sep(1); fwrite(s.p, 1, (size_t)s.n, ld_file); }
void f_ld_i(long v) { ld_sep(0); fprintf(ld_file, "%11ld", v); }

void f_ld_r(double v)
{
    ld_sep(0);
    char out[64];
    if (isnan(v)) { fprintf(ld_file, "%25s", "NaN"); return; }
    if (isinf(v)) { fprintf(ld_file, "%25s", v > 0 ? "Infinity" : "-Infinity"); return; }
    if (v == 0.) { fprintf(ld_file, "%20s     ", signbit(v) ? "-0.0000000000000000" : "0.0000000000000000"); return; }
    char e[64];
    snprintf(e, sizeof e, "%.16e", fabs(v));            /* d.dddddddddddddddde+XX : 17 significant digits */
    char dig[18];
    dig[0] = e[0];
    memcpy(dig + 1, e + 2, 16);
    dig[17] = 0;
    int x = atoi(strchr(e, 'e') + 1), e10 = x + 1;      /* 10^(e10-1) <= |v| < 10^e10 after rounding */
    const char *sg = v < 0 ? "-" : "";
    if (e10 >= 0 && e10 <= 17) {
        char body[64];
        if (e10 == 0) snprintf(body, sizeof body, "%s0.%s", sg, dig);
        else {
            char ip[32], fp[32];
            memcpy(ip, dig, (size_t)e10); ip[e10] = 0;
            snprintf(fp, sizeof fp, "%s", dig + e10);
            snprintf(body, sizeof body, "%s%s.%s", sg, ip, fp);
        }
        snprintf(out, sizeof out, "%20s     ", body);
    } else {
        char body[64];
        snprintf(body, sizeof body, "%s%c.%sE%c%03d", sg, dig[0], dig + 1, x < 0 ? '-' : '+', abs(x));
        snprintf(out, sizeof out, "%25s", body);
    }
    fputs(out, ld_file);
}

void f_ld_end(void) { fputc('\n', ld_file); }

/* ================================================================= misc */
void f_cpu_time(double *t) { *t = (double)clock() / CLOCKS_PER_SEC; }

static char args[4][1024];
void ref_set_arg(int n, const char *value) { if (n >= 0 && n < 4) snprintf(args[n], sizeof args[n], "%s", value); }
void f_getarg(int n, char *dst, long len)
{
    const char *s = (n >= 0 && n < 4) ? args[n] : "";
    f_str_assign(dst, len, f_lit(s, (long)strlen(s)));
}
