/*
 * ref_runtime.h -- run-time support for the C code that oracle/f90_to_c.py emits from the
 * reference's Fortran sources.  TEST INFRASTRUCTURE (oracle/), never linked into the product.
 *
 * Nothing in here restates the reference's algorithm: it is the part of a Fortran run time the
 * translated code needs (array descriptors, blank-padded strings, stream / formatted I/O, PRINT
 * capture, STOP).  Semantics it stands in for:
 *   - gfortran MAX/MIN on reals:  mvar = a; if (b .op. mvar || isnan(mvar)) mvar = b
 *   - ALLOCATE / automatic arrays: every allocation gets a guard zone on both sides filled with 0xFF
 *     bytes (NaN for reals, -1 for integers): the reference reads a few planes outside its arrays in
 *     places (set3d.f90:528-536, firstDeriv order 8 on band cells near the boundary); in the gfortran
 *     binary those reads return unrelated heap contents, here they return a quiet NaN, so anything
 *     that depends on them is visibly poisoned instead of silently arbitrary.
 *   - list-directed output (PRINT*, WRITE(u,*)): values are CAPTURED (see ref_print_*), and written
 *     to files with the spacing libgfortran uses for INTEGER(4) (I12 incl. separator) and REAL(8)
 *     (1PG25.17E3 incl. separator) -- that spacing is libgfortran's, not the reference's; it is
 *     restated from the library's documented behaviour and marked as such in DESIGN.md.
 */
#ifndef REF_RUNTIME_H
#define REF_RUNTIME_H
#include <math.h>
#include <setjmp.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- array descriptors (ALLOCATABLE arrays, rank <= 4) --------------------------------------- */
typedef struct {
    char  *base;      /* address of the element with all subscripts at their lower bound; NULL = not allocated */
    int    rank;
    long   lb[4], ext[4];
    size_t elsz;
    char  *raw;       /* malloc'ed block incl. guard zones */
} f_desc;

void  f_allocate(f_desc *d, int rank, size_t elsz, const long *lb, const long *ub);
void  f_deallocate(f_desc *d);
void  f_assign_alloc(f_desc *dst, const f_desc *src);   /* dst = src with (re)allocation on assignment */
long  f_size(const f_desc *d);
void *f_auto(long nelem, size_t elsz, long guard_elems); /* automatic (stack) arrays of a procedure */
void  f_auto_free(void *p, size_t elsz, long guard_elems);

/* ---- reals -------------------------------------------------------------------------------------- */
static inline double f_max(double a, double b) { return (b > a || isnan(a)) ? b : a; }
static inline double f_min(double a, double b) { return (b < a || isnan(a)) ? b : a; }
static inline float  f_maxf(float a, float b) { return (b > a || isnan(a)) ? b : a; }
static inline float  f_minf(float a, float b) { return (b < a || isnan(a)) ? b : a; }
static inline int    f_imax(int a, int b) { return b > a ? b : a; }
static inline int    f_imin(int a, int b) { return b < a ? b : a; }

/* ---- blank-padded character values ------------------------------------------------------------ */
typedef struct { const char *p; long n; } fstr;
fstr f_lit(const char *s, long n);
fstr f_var(const char *p, long n);
fstr f_concat(fstr a, fstr b);
fstr f_substr(fstr a, long lo, long hi);
fstr f_trim(fstr a);
fstr f_char(int code);
int  f_len_trim(fstr a);
int  f_str_eq(fstr a, fstr b);
void f_str_assign(char *dst, long n, fstr src);

/* ---- STOP --------------------------------------------------------------------------------------- */
extern jmp_buf ref_stop_jmp;
extern int     ref_stop_armed;
void f_stop(void);

/* ---- PRINT capture ------------------------------------------------------------------------------ */
typedef struct { int line; int kind; /* 0 str 1 int 2 real 3 end-of-record */ long i; double r; char s[96]; } ref_print_item;
void f_pr_begin(int line);
void f_pr_s(fstr s);
void f_pr_i(long v);
void f_pr_r(double v);
void f_pr_end(void);
long ref_print_count(void);
const ref_print_item *ref_print_get(long idx);
void ref_print_clear(void);
void ref_print_echo(int on);
void ref_print_limit(long max_items);    /* ring behaviour off: items beyond the limit are dropped */

/* ---- external files ----------------------------------------------------------------------------- */
void f_set_outdir(const char *dir);      /* prefix for files opened with a relative name */
void f_open(int unit, fstr file, fstr status, fstr access, fstr form);
void f_close(int unit);
void f_read(int unit, void *dst, size_t nbytes);
void f_write(int unit, const void *src, size_t nbytes);
void f_write_s(int unit, fstr s);
/* formatted output with an explicit format: unit -1 = internal (to dst), unit -2 = stdout */
void f_fmt_begin(fstr fmt);
void f_fmt_s(fstr s);
void f_fmt_i(long v);
void f_fmt_r(double v);
void f_fmt_end_internal(char *dst, long n);
void f_fmt_end_unit(int unit);
/* list-directed output to a unit */
void f_ld_begin(int unit, int line);
void f_ld_s(fstr s);
void f_ld_i(long v);
void f_ld_r(double v);
void f_ld_end(void);

void f_cpu_time(double *t);
void f_getarg(int n, char *dst, long len);
void ref_set_arg(int n, const char *value);

#ifdef __cplusplus
}
#endif
#endif
