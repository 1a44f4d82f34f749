/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Batch driver around the UNMODIFIED reference headers.  The reference sources are
 * compiled from where they lie (-I/root/reference/src); nothing is copied.  The only
 * things changed are done by the preprocessor, from the outside:
 *
 *   exit(c)          -> oracle_trap(c)        longjmp back to the per-item loop, so every
 *                                             reference exit path becomes a per-item status
 *                                             
 *   __assert_fail    -> oracle_assert_fail    same for assert() (asserts stay live: the
 *                                             reference builds without -DNDEBUG, test.sh:12)
 *   fprintf/printf   -> swallowed             the reference prints one line per failing item
 *   malloc/calloc/free -> per-thread bump arena reset per item (ORACLE_ARENA=1, default):
 *                                             plonk_prove leaks 1856 B per call and every
 *                                             trapped exit leaks everything live, so libc
 *                                             malloc would run out of memory on 2^20 items.
 *                                             This makes the CPU baseline FASTER, not slower.
 *   srs_eval_at_s    -> counted wrapper       (only inside plonk.h) so that the four
 *                                             "exceeds SRS size" exits can be told apart; it
 *                                             also records each commitment, and
 *   poly_eval        -> logging wrapper       (only inside plonk.h) records each evaluation:
 *                                             the Fiat-Shamir driver (fs_spec.inc) needs the
 *                                             commitments and openings of items that exit later
 *
 * Output of this file: oracle/_ref/libref_oracle.so (git-ignored, travels to the GPU box).
 * Role: (1) pins the C restatement in oracle/plonk_port.c, (2) generates tests/golden/,
 * (3) is the `--impl reference` CPU arm of bench.py (kind "reference").
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdbool.h>
#include <setjmp.h>
#include <pthread.h>
#include <time.h>

#ifndef ORACLE_ARENA
#define ORACLE_ARENA 1
#endif

/* ------------------------------------------------------------------ trap layer */
typedef struct {
  jmp_buf jb;
  int armed;
  int kind;          /* 1 = exit(), 2 = assert */
  unsigned line;     /* assert line */
  const char *fmt;   /* last fprintf format string seen (classifies the exit) */
  int commits;       /* srs_eval_at_s calls completed inside plonk_prove */
  unsigned char commit_bytes[12][3];   /* ... and what they returned (x, y, infinite), for the Fiat-Shamir driver */
  int n_evals;                         /* poly_eval calls completed inside plonk_prove ... */
  unsigned char eval_val[32], eval_commits[32];   /* ... their values and the commit count at the time */
  unsigned char *arena;
  size_t arena_off, arena_cap;
} oracle_tls_t;

static __thread oracle_tls_t T;

__attribute__((noreturn)) static void oracle_trap(int code) {
  (void)code;
  if (!T.armed) { fputs("ref_driver: reference exit() outside a trapped region\n", stderr); abort(); }
  T.kind = 1;
  longjmp(T.jb, 1);
}

__attribute__((noreturn)) void oracle_assert_fail(const char *expr, const char *file,
                                                  unsigned int line, const char *func) {
  (void)expr; (void)file; (void)func;
  if (!T.armed) { fprintf(stderr, "ref_driver: assert outside trapped region: %s:%u\n", file, line); abort(); }
  T.kind = 2;
  T.line = line;
  longjmp(T.jb, 2);
}

static int oracle_note(const char *fmt, ...) { T.fmt = fmt; return 0; }
static int oracle_mute(const char *fmt, ...) { (void)fmt; return 0; }

#if ORACLE_ARENA
#define ORACLE_ARENA_BYTES (4u << 20)
static void *oracle_malloc(size_t n) {
  if (!T.arena) {
    T.arena = (unsigned char *)(malloc)(ORACLE_ARENA_BYTES);
    T.arena_cap = ORACLE_ARENA_BYTES;
    T.arena_off = 0;
    if (!T.arena) abort();
  }
  size_t need = (n + 15u) & ~(size_t)15u;
  if (need == 0) need = 16;
  if (T.arena_off + need > T.arena_cap) { fputs("ref_driver: arena exhausted\n", stderr); abort(); }
  void *p = T.arena + T.arena_off;
  T.arena_off += need;
  return p;
}
static void *oracle_calloc(size_t a, size_t b) {
  void *p = oracle_malloc(a * b);
  memset(p, 0, a * b);
  return p;
}
static void oracle_free(void *p) { (void)p; }
#endif

/* system headers first, so that the macros below only touch the reference's code */
#include <assert.h>
#undef assert
#define __assert_fail oracle_assert_fail
#define exit(c) oracle_trap(c)
#define fprintf(stream, ...) oracle_note(__VA_ARGS__)
#define printf(...) oracle_mute(__VA_ARGS__)
#if ORACLE_ARENA
#define malloc(n) oracle_malloc(n)
#define calloc(a, b) oracle_calloc(a, b)
#define free(p) oracle_free(p)
#endif

/* ------------------------------------------------------------------ the reference, unmodified */
#include "hf.h"
#include "gf.h"
#include "poly.h"
#include "matrix.h"
#include "g1.h"
#include "g2.h"
#include "gt.h"
#include "srs.h"
static G1 oracle_counted_commit(const SRS *srs, const POLY *p) {
  G1 r = srs_eval_at_s(srs, p);
  if (T.commits < 12) {
    T.commit_bytes[T.commits][0] = r.x.value;
    T.commit_bytes[T.commits][1] = r.y.value;
    T.commit_bytes[T.commits][2] = r.infinite ? 1 : 0;
  }
  T.commits++;
  return r;
}
static HF oracle_logged_eval(const POLY *p, HF x) {
  HF r = poly_eval(p, x);
  if (T.n_evals < 32) {
    T.eval_val[T.n_evals] = r.value;
    T.eval_commits[T.n_evals] = (unsigned char)T.commits;
  }
  T.n_evals++;
  return r;
}
#define srs_eval_at_s(s, p) oracle_counted_commit(s, p)
#define poly_eval(p, x) oracle_logged_eval(p, x)
#include "plonk.h"
#undef srs_eval_at_s
#undef poly_eval
#include "pairing.h"

/* ------------------------------------------------------------------ helpers */
#define TRAP_BEGIN() (T.armed = 1, T.kind = 0, T.line = 0, T.fmt = NULL, setjmp(T.jb))
#define TRAP_END() (T.armed = 0)

static size_t arena_mark(void) {
#if ORACLE_ARENA
  return T.arena_off;
#else
  return 0;
#endif
}
static void arena_reset(size_t mark) {
#if ORACLE_ARENA
  T.arena_off = mark;
#else
  (void)mark;
#endif
}

typedef void (*range_fn)(void *ctx, size_t lo, size_t hi);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; } job_t;
static void *job_main(void *p) {
  job_t *j = (job_t *)p;
  j->fn(j->ctx, j->lo, j->hi);
#if ORACLE_ARENA
  if (T.arena) { (free)(T.arena); T.arena = NULL; }
#endif
  return NULL;
}
static void run_ranges(range_fn fn, void *ctx, size_t n, int nthreads) {
  if (nthreads <= 1 || n < 2) { fn(ctx, 0, n); return; }
  if ((size_t)nthreads > n) nthreads = (int)n;
  pthread_t *th = (pthread_t *)(malloc)(sizeof(pthread_t) * nthreads);
  job_t *jobs = (job_t *)(malloc)(sizeof(job_t) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t].fn = fn; jobs[t].ctx = ctx;
    jobs[t].lo = n * (size_t)t / nthreads;
    jobs[t].hi = n * (size_t)(t + 1) / nthreads;
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  (free)(th); (free)(jobs);
}

static G1 g1_from(const uint8_t *b) { G1 p; p.x.value = b[0]; p.y.value = b[1]; p.infinite = b[2] != 0; return p; }
static void g1_to(uint8_t *b, G1 p) { b[0] = p.x.value; b[1] = p.y.value; b[2] = p.infinite ? 1 : 0; }
static G2 g2_from(const uint8_t *b) { G2 p; p.x.value = b[0]; p.y.value = b[1]; return p; }
static void g2_to(uint8_t *b, G2 p) { b[0] = p.x.value; b[1] = p.y.value; }
static GTP gt_from(const uint8_t *b) { GTP p; p.a.value = b[0]; p.b.value = b[1]; return p; }
static void gt_to(uint8_t *b, GTP p) { b[0] = p.a.value; b[1] = p.b.value; }

int ref_abi_version(void) { return 1; }
int ref_uses_arena(void) { return ORACLE_ARENA; }
int ref_sizeof_proof(void) { return (int)sizeof(PROOF); }

/* ------------------------------------------------------------------ family (1): fields */
/* op: 0 add 1 sub 2 mul 3 div 4 neg(a) 5 inv(a) 6 pow(a, b as exponent) ; field: 17 or 101 */
void ref_field_op(int field, int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (field == 17) {
      HF x = {a[i]}, y = {b ? b[i] : 0}, r;
      switch (op) {
        case 0: r = hf_add(x, y); break;
        case 1: r = hf_sub(x, y); break;
        case 2: r = hf_mul(x, y); break;
        case 3: r = hf_div(x, y); break;
        case 4: r = hf_neg(x); break;
        case 5: r = hf_inv(x); break;
        default: r = hf_pow(x, b[i]); break;
      }
      out[i] = r.value;
    } else {
      GF x = {a[i]}, y = {b ? b[i] : 0}, r;
      switch (op) {
        case 0: r = gf_add(x, y); break;
        case 1: r = gf_sub(x, y); break;
        case 2: r = gf_mul(x, y); break;
        case 3: r = gf_div(x, y); break;
        case 4: r = gf_neg(x); break;
        case 5: r = gf_inv(x); break;
        default: r = gf_pow(x, b[i]); break;
      }
      out[i] = r.value;
    }
  }
}
uint8_t ref_hf_new(int64_t v) { return hf_new(v).value; }
uint8_t ref_gf_new(int64_t v) { return gf_new(v).value; }

/* ------------------------------------------------------------------ family (2): polynomials */
static POLY poly_from(const uint8_t *c, size_t len) {
  HF tmp[256];
  for (size_t i = 0; i < len; i++) tmp[i].value = c[i];
  return poly_new(tmp, len);
}
static void poly_to(uint8_t *c, size_t stride, uint8_t *len, const POLY *p) {
  memset(c, 0, stride);
  for (size_t i = 0; i < p->len && i < stride; i++) c[i] = p->coeffs[i].value;
  *len = (uint8_t)p->len;
}

/* op: 0 add, 1 sub, 2 mul.  a:[n][sa] raw coefficients (poly_new trims), alen:[n] */
void ref_poly_binop(int op, const uint8_t *a, const uint8_t *alen, size_t sa,
                    const uint8_t *b, const uint8_t *blen, size_t sb,
                    uint8_t *out, uint8_t *olen, size_t so, size_t n) {
  size_t mark = arena_mark();
  for (size_t i = 0; i < n; i++) {
    POLY pa = poly_from(a + i * sa, alen[i]), pb = poly_from(b + i * sb, blen[i]);
    POLY r = op == 0 ? poly_add(&pa, &pb) : op == 1 ? poly_sub(&pa, &pb) : poly_mul(&pa, &pb);
    poly_to(out + i * so, so, olen + i, &r);
    poly_free(&pa); poly_free(&pb); poly_free(&r);
    arena_reset(mark);
  }
}

/* status[i] = 1 when the reference would exit ("Division by zero polynomial") */
void ref_poly_divide(const uint8_t *num, const uint8_t *nlen, size_t sn,
                     const uint8_t *den, const uint8_t *dlen, size_t sd,
                     uint8_t *quot, uint8_t *qlen, size_t sq,
                     uint8_t *rem, uint8_t *rlen, size_t sr, uint8_t *status, size_t n) {
  size_t mark = arena_mark();
  for (size_t i = 0; i < n; i++) {
    status[i] = 0;
    if (TRAP_BEGIN() == 0) {
      POLY pn = poly_from(num + i * sn, nlen[i]), pd = poly_from(den + i * sd, dlen[i]);
      POLY q, r;
      poly_divide(&pn, &pd, &q, &r);
      poly_to(quot + i * sq, sq, qlen + i, &q);
      poly_to(rem + i * sr, sr, rlen + i, &r);
    } else {
      status[i] = 1;
      memset(quot + i * sq, 0, sq); memset(rem + i * sr, 0, sr);
      qlen[i] = 0; rlen[i] = 0;
    }
    TRAP_END();
    arena_reset(mark);
  }
}

void ref_poly_eval(const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *x, uint8_t *out, size_t n) {
  size_t mark = arena_mark();
  for (size_t i = 0; i < n; i++) {
    POLY pp = poly_from(p + i * sp, plen[i]);
    HF xv = {x[i]};
    out[i] = poly_eval(&pp, xv).value;
    poly_free(&pp);
    arena_reset(mark);
  }
}

/* op: 0 scale(p, k[i]) 1 negate 2 shift(p, k[i]) 3 add_hf(p, k[i]) ; status 1 = exit */
void ref_poly_unop(int op, const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *k,
                   uint8_t *out, uint8_t *olen, size_t so, size_t n) {
  size_t mark = arena_mark();
  for (size_t i = 0; i < n; i++) {
    POLY pp = poly_from(p + i * sp, plen[i]);
    HF kv = {k ? k[i] : 0};
    POLY r;
    if (op == 0) r = poly_scale(&pp, kv);
    else if (op == 1) r = poly_negate(&pp);
    else if (op == 2) r = poly_shift(&pp, k[i]);
    else r = poly_add_hf(&pp, kv);
    poly_to(out + i * so, so, olen + i, &r);
    arena_reset(mark);
  }
}

void ref_poly_slice(const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *start, const uint8_t *end,
                    uint8_t *out, uint8_t *olen, size_t so, uint8_t *status, size_t n) {
  size_t mark = arena_mark();
  for (size_t i = 0; i < n; i++) {
    status[i] = 0;
    if (TRAP_BEGIN() == 0) {
      POLY pp = poly_from(p + i * sp, plen[i]);
      POLY r = poly_slice(&pp, start[i], end[i]);
      poly_to(out + i * so, so, olen + i, &r);
    } else {
      status[i] = 1; memset(out + i * so, 0, so); olen[i] = 0;
    }
    TRAP_END();
    arena_reset(mark);
  }
}

void ref_poly_z(const uint8_t *points, size_t len, uint8_t *out, uint8_t *olen, size_t so) {
  size_t mark = arena_mark();
  HF pts[256];
  for (size_t i = 0; i < len; i++) pts[i].value = points[i];
  POLY r = poly_z(pts, len);
  poly_to(out, so, olen, &r);
  arena_reset(mark);
}

/* status 1 = exit (duplicate x) */
void ref_poly_lagrange(const uint8_t *xs, const uint8_t *ys, size_t len, size_t n,
                       uint8_t *out, uint8_t *olen, size_t so, uint8_t *status) {
  size_t mark = arena_mark();
  for (size_t i = 0; i < n; i++) {
    status[i] = 0;
    if (TRAP_BEGIN() == 0) {
      HF x[64], y[64];
      for (size_t j = 0; j < len; j++) { x[j].value = xs[i * len + j]; y[j].value = ys[i * len + j]; }
      POLY r = poly_lagrange(x, y, len);
      poly_to(out + i * so, so, olen + i, &r);
    } else {
      status[i] = 1; memset(out + i * so, 0, so); olen[i] = 0;
    }
    TRAP_END();
    arena_reset(mark);
  }
}

/* matrices: row-major [m][n] bytes */
void ref_matrix_mul(const uint8_t *a, size_t am, size_t an, const uint8_t *b, size_t bm, size_t bn, uint8_t *out) {
  size_t mark = arena_mark();
  MATRIX A = matrix_zero(am, an), B = matrix_zero(bm, bn);
  for (size_t i = 0; i < am * an; i++) A.v[i].value = a[i];
  for (size_t i = 0; i < bm * bn; i++) B.v[i].value = b[i];
  MATRIX R = matrix_mul(&A, &B);
  for (size_t i = 0; i < am * bn; i++) out[i] = R.v[i].value;
  arena_reset(mark);
}
void ref_matrix_inv(const uint8_t *a, size_t n, uint8_t *out) {
  size_t mark = arena_mark();
  MATRIX A = matrix_zero(n, n);
  for (size_t i = 0; i < n * n; i++) A.v[i].value = a[i];
  MATRIX R = matrix_inv(&A);
  for (size_t i = 0; i < n * n; i++) out[i] = R.v[i].value;
  arena_reset(mark);
}
void ref_matrix_gauss_jordan(uint8_t *a, size_t m, size_t n) {
  size_t mark = arena_mark();
  MATRIX A = matrix_zero(m, n);
  for (size_t i = 0; i < m * n; i++) A.v[i].value = a[i];
  matrix_gauss_jordan(&A);
  for (size_t i = 0; i < m * n; i++) a[i] = A.v[i].value;
  arena_reset(mark);
}

/* ------------------------------------------------------------------ family (3): groups */
/* op: 0 add 1 double(a) 2 neg(a) */
void ref_g1_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G1 x = g1_from(a + 3 * i), r;
    if (op == 0) { G1 y = g1_from(b + 3 * i); r = g1_add(&x, &y); }
    else if (op == 1) r = g1_double(&x);
    else r = g1_neg(&x);
    g1_to(out + 3 * i, r);
  }
}
typedef struct { const uint8_t *p; const uint64_t *s; uint8_t *out; } g1mul_ctx;
static void g1mul_range(void *c, size_t lo, size_t hi) {
  g1mul_ctx *x = (g1mul_ctx *)c;
  for (size_t i = lo; i < hi; i++) {
    G1 p = g1_from(x->p + 3 * i);
    g1_to(x->out + 3 * i, g1_mul(&p, x->s[i]));
  }
}
void ref_g1_mul(const uint8_t *p, const uint64_t *scalars, uint8_t *out, size_t n, int nthreads) {
  g1mul_ctx c = {p, scalars, out};
  run_ranges(g1mul_range, &c, n, nthreads);
}
void ref_g1_is_on_curve(const uint8_t *p, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) { G1 x = g1_from(p + 3 * i); out[i] = g1_is_on_curve(&x) ? 1 : 0; }
}
/* op: 0 add 2 neg */
void ref_g2_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G2 x = g2_from(a + 2 * i), r;
    if (op == 0) { G2 y = g2_from(b + 2 * i); r = g2_add(&x, &y); }
    else r = g2_neg(&x);
    g2_to(out + 2 * i, r);
  }
}
/* scalar 0 is undefined behaviour in the reference (g2.h:69-83): callers must not pass it */
void ref_g2_mul(const uint8_t *p, const uint64_t *scalars, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (scalars[i] == 0) { out[2 * i] = 0xFF; out[2 * i + 1] = 0xFF; continue; }
    g2_to(out + 2 * i, g2_mul(g2_from(p + 2 * i), scalars[i]));
  }
}
void ref_gtp_mul(const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    GTP x = gt_from(a + 2 * i), y = gt_from(b + 2 * i);
    gt_to(out + 2 * i, gtp_mul(&x, &y));
  }
}
void ref_gtp_pow(const uint8_t *a, const uint64_t *e, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    GTP x = gt_from(a + 2 * i);
    gt_to(out + 2 * i, gtp_pow(&x, e[i]));
  }
}

void ref_srs_create(uint8_t secret, uint32_t n, uint8_t *g1s_out, uint8_t *g2_out) {
  size_t mark = arena_mark();
  GF s = {secret};
  SRS srs = srs_create(s, n);
  for (size_t i = 0; i < srs.len; i++) g1_to(g1s_out + 3 * i, srs.g1s[i]);
  g2_to(g2_out, srs.g2_1);
  g2_to(g2_out + 2, srs.g2_s);
  srs_free(&srs);
  arena_reset(mark);
}

static SRS srs_from(const uint8_t *g1s, uint32_t len, const uint8_t *g2) {
  SRS srs;
  srs.len = len;
  srs.g1s = (G1 *)malloc(len * sizeof(G1) + 1);
  for (uint32_t i = 0; i < len; i++) srs.g1s[i] = g1_from(g1s + 3 * i);
  srs.g2_1 = g2_from(g2);
  srs.g2_s = g2_from(g2 + 2);
  return srs;
}

typedef struct {
  const uint8_t *g1s; uint32_t srs_len; const uint8_t *g2;
  const uint8_t *polys; const uint8_t *plen; size_t sp; uint8_t *out; uint8_t *status;
} commit_ctx;
static void commit_range(void *c, size_t lo, size_t hi) {
  commit_ctx *x = (commit_ctx *)c;
  size_t mark0 = arena_mark();
  SRS srs = srs_from(x->g1s, x->srs_len, x->g2);
  size_t mark = arena_mark();
  for (size_t i = lo; i < hi; i++) {
    x->status[i] = 0;
    if (TRAP_BEGIN() == 0) {
      POLY p = poly_from(x->polys + i * x->sp, x->plen[i]);
      G1 r = srs_eval_at_s(&srs, &p);
      g1_to(x->out + 3 * i, r);
    } else {
      x->status[i] = 1; memset(x->out + 3 * i, 0, 3);
    }
    TRAP_END();
    arena_reset(mark);
  }
#if !ORACLE_ARENA
  srs_free(&srs);
#endif
  arena_reset(mark0);
}
/* KZG commit: status 1 = "exceeds SRS size" exit */
void ref_srs_eval_at_s(const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2,
                       const uint8_t *polys, const uint8_t *plen, size_t sp,
                       uint8_t *out, uint8_t *status, size_t n, int nthreads) {
  commit_ctx c = {g1s, srs_len, g2, polys, plen, sp, out, status};
  run_ranges(commit_range, &c, n, nthreads);
}

/* ------------------------------------------------------------------ family (4): pairing */
void ref_line(const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G1 x = g1_from(a + 3 * i), y = g1_from(b + 3 * i);
    LINE_EQ l = line(&x, &y);
    out[3 * i] = l.x.value; out[3 * i + 1] = l.y.value; out[3 * i + 2] = l.c.value;
  }
}
typedef struct { const uint8_t *p, *q; uint8_t *out; uint64_t r; } pair_ctx;
static void pair_range(void *c, size_t lo, size_t hi) {
  pair_ctx *x = (pair_ctx *)c;
  for (size_t i = lo; i < hi; i++) {
    G1 p = g1_from(x->p + 3 * i);
    G2 q = g2_from(x->q + 2 * i);
    GTP r = x->r ? pairing_f(x->r, &p, &q) : pairing(&p, &q);
    gt_to(x->out + 2 * i, r);
  }
}
void ref_pairing(const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n, int nthreads) {
  pair_ctx c = {p, q, out, 0};
  run_ranges(pair_range, &c, n, nthreads);
}
void ref_pairing_f(uint64_t r, const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n) {
  pair_ctx c = {p, q, out, r};
  pair_range(&c, 0, n);
}

/* ------------------------------------------------------------------ protocol */
/* circuit: 44 bytes  q_l[4] q_r[4] q_o[4] q_m[4] q_c[4] | ca_type[4] ca_idx[4] cb_type[4] cb_idx[4] cc_type[4] cc_idx[4] */
typedef struct {
  PLONK plonk;
  CONSTRAINTS cons;
} ref_ctx_t;

static void ctx_build(ref_ctx_t *c, const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2) {
  SRS srs = srs_from(g1s, srs_len, g2);
  c->plonk = plonk_new(srs, 4);
  CONSTRAINTS *k = &c->cons;
  k->num_constraints = 4;
  k->num_gates = 4;
  HF **sel[5] = {&k->q_l, &k->q_r, &k->q_o, &k->q_m, &k->q_c};
  for (int s = 0; s < 5; s++) {
    *sel[s] = (HF *)malloc(4 * sizeof(HF));
    for (int i = 0; i < 4; i++) (*sel[s])[i].value = circuit[4 * s + i];
  }
  COPY_OF **cp[3] = {&k->c_a, &k->c_b, &k->c_c};
  for (int s = 0; s < 3; s++) {
    *cp[s] = (COPY_OF *)malloc(4 * sizeof(COPY_OF));
    for (int i = 0; i < 4; i++) {
      (*cp[s])[i].type = (COPY_OF_TYPE)circuit[20 + 8 * s + i];
      (*cp[s])[i].index = circuit[20 + 8 * s + 4 + i];
    }
  }
}
static void ctx_free(ref_ctx_t *c) {
#if !ORACLE_ARENA
  constraints_free(&c->cons);
  plonk_free(&c->plonk);
#else
  (void)c;
#endif
}

/* dump what plonk_new computes: h[4] k1_h[4] k2_h[4] h_pows_inv[16] z_h[8]+len  -> out[37] */
void ref_plonk_setup_dump(uint8_t *out) {
  size_t mark = arena_mark();
  uint8_t g1s[3 * 7], g2[4];
  ref_srs_create(2, 6, g1s, g2);
  SRS srs = srs_from(g1s, 7, g2);
  PLONK p = plonk_new(srs, 4);
  for (int i = 0; i < 4; i++) { out[i] = p.h[i].value; out[4 + i] = p.k1_h[i].value; out[8 + i] = p.k2_h[i].value; }
  for (int i = 0; i < 16; i++) out[12 + i] = p.h_pows_inv.v[i].value;
  memset(out + 28, 0, 8);
  for (size_t i = 0; i < p.z_h_x.len; i++) out[28 + i] = p.z_h_x.coeffs[i].value;
  out[36] = (uint8_t)p.z_h_x.len;
  arena_reset(mark);
}

void ref_copy_constraints_to_roots(const uint8_t *types, const uint8_t *idx, size_t len, uint8_t *sigma) {
  size_t mark = arena_mark();
  uint8_t g1s[3 * 7], g2[4];
  ref_srs_create(2, 6, g1s, g2);
  SRS srs = srs_from(g1s, 7, g2);
  PLONK p = plonk_new(srs, 4);
  COPY_OF co[64]; HF s[64];
  for (size_t i = 0; i < len; i++) { co[i].type = (COPY_OF_TYPE)types[i]; co[i].index = idx[i]; }
  copy_constraints_to_roots(&p, co, len, s);
  for (size_t i = 0; i < len; i++) sigma[i] = s[i].value;
  arena_reset(mark);
}

void ref_interpolate_at_h(const uint8_t *vals, uint8_t *out, uint8_t *olen, size_t n) {
  size_t mark0 = arena_mark();
  uint8_t g1s[3 * 7], g2[4];
  ref_srs_create(2, 6, g1s, g2);
  SRS srs = srs_from(g1s, 7, g2);
  PLONK p = plonk_new(srs, 4);
  size_t mark = arena_mark();
  for (size_t i = 0; i < n; i++) {
    HF v[4];
    for (int j = 0; j < 4; j++) v[j].value = vals[4 * i + j];
    POLY r = interpolate_at_h(&p, v, 4);
    poly_to(out + 4 * i, 4, olen + i, &r);
    poly_free(&r);
    arena_reset(mark);
  }
#if !ORACLE_ARENA
  plonk_free(&p);
#endif
  arena_reset(mark0);
}

/* SURVEY.md Appendix B numbering of the reference's exit paths */
static uint8_t classify_trap(void) {
  if (T.kind == 2) {
    if (T.line == 231) return 1;              /* constraints not satisfied */
    if (T.line == 368) return 6;              /* acc_x(omega^n) != 1 */
    return 11;                                /* opening remainders (610, 617) */
  }
  const char *f = T.fmt ? T.fmt : "";
  if (strstr(f, "Non-zero remainder")) return 8;
  if (strstr(f, "Invalid slice")) return 9;
  if (strstr(f, "exceeds SRS size")) {
    if (T.commits < 3) return 5;
    if (T.commits == 3) return 7;
    if (T.commits < 7) return 10;
    return 12;
  }
  if (strstr(f, "Invalid copy_of")) return 3;
  if (strstr(f, "Length mismatch")) return 4;
  return 2;
}

typedef struct {
  const uint8_t *circuit, *g1s; uint32_t srs_len; const uint8_t *g2;
  const uint8_t *wit, *rnd, *chal; uint8_t *proofs, *status;
} prove_ctx;

static void prove_range(void *c, size_t lo, size_t hi) {
  prove_ctx *x = (prove_ctx *)c;
  size_t mark0 = arena_mark();
  ref_ctx_t ctx;
  ctx_build(&ctx, x->circuit, x->g1s, x->srs_len, x->g2);
  size_t mark = arena_mark();
  for (size_t i = lo; i < hi; i++) {
    HF a[4], b[4], cc[4], rnd[9];
    for (int j = 0; j < 4; j++) {
      a[j].value = x->wit[12 * i + j];
      b[j].value = x->wit[12 * i + 4 + j];
      cc[j].value = x->wit[12 * i + 8 + j];
    }
    for (int j = 0; j < 9; j++) rnd[j].value = x->rnd[9 * i + j];
    ASSIGNMENTS as; as.a = a; as.b = b; as.c = cc; as.len = 4;
    CHALLENGE ch;
    ch.alpha.value = x->chal[5 * i]; ch.beta.value = x->chal[5 * i + 1]; ch.gamma.value = x->chal[5 * i + 2];
    ch.z.value = x->chal[5 * i + 3]; ch.v.value = x->chal[5 * i + 4];
    T.commits = 0;
    if (TRAP_BEGIN() == 0) {
      PROOF pr = plonk_prove(&ctx.plonk, &ctx.cons, &as, &ch, rnd);
      uint8_t *o = x->proofs + 34 * i;
      G1 *g = &pr.a_s;
      for (int j = 0; j < 9; j++) g1_to(o + 3 * j, g[j]);
      HF *s = &pr.a_z;
      for (int j = 0; j < 7; j++) o[27 + j] = s[j].value;
      x->status[i] = 0;
    } else {
      memset(x->proofs + 34 * i, 0, 34);
      x->status[i] = classify_trap();
    }
    TRAP_END();
    arena_reset(mark);
  }
  ctx_free(&ctx);
  arena_reset(mark0);
}

/* plonk_prove over a batch.  wit:[n][12] (a[4] b[4] c[4]), rnd:[n][9], chal:[n][5] (alpha beta gamma z v).
 * proofs:[n][34] in PROOF field order with G1 as (x, y, infinite); status:[n] Appendix-B row (0 = completed). */
void ref_plonk_prove_batch(const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2,
                           const uint8_t *wit, const uint8_t *rnd, const uint8_t *chal, size_t n,
                           uint8_t *proofs, uint8_t *status, int nthreads) {
#if !ORACLE_ARENA
  /* libc malloc build: the reference leaks ~1.9 kB per call; callers must bound n per process */
#endif
  prove_ctx c = {circuit, g1s, srs_len, g2, wit, rnd, chal, proofs, status};
  run_ranges(prove_range, &c, n, nthreads);
}

/* --------------------------- Fiat-Shamir mode (spec: fs_spec.inc).
 * The reference takes its challenges up front, so the transcript is driven from outside: plonk_prove is
 * run up to five times per item, each pass with the challenges known so far (zero for the rest), and the
 * commitments / openings it produced on the way are read from the capture hooks above.  Everything a pass
 * computes before its first not-yet-known challenge is used is what the final run computes, so the
 * commitments and the exits of the rounds already fixed are final.  The last pass runs the unmodified
 * prover with the complete challenge set: its PROOF bytes and exit status are the expected output. */
#include "fs_spec.inc"

typedef struct {
  const uint8_t *circuit, *g1s; uint32_t srs_len; const uint8_t *g2;
  const uint8_t *wit, *rnd; uint8_t *proofs, *status, *chal;
} prove_fs_ctx;

/* one trapped run of the unmodified prover (its own frame, so that nothing the caller changes between passes
 * lives across the setjmp): returns 1 and the PROOF record in o if it completed, else 0 and the exit status */
__attribute__((noinline)) static int fs_pass(ref_ctx_t *ctx, ASSIGNMENTS *as, HF *rnd, const uint8_t *ch6, uint8_t *o, uint8_t *status) {
  CHALLENGE ch;
  ch.alpha.value = ch6[0]; ch.beta.value = ch6[1]; ch.gamma.value = ch6[2]; ch.z.value = ch6[3]; ch.v.value = ch6[4];
  T.commits = 0;
  T.n_evals = 0;
  volatile int completed = 0;
  if (TRAP_BEGIN() == 0) {
    PROOF pr = plonk_prove(&ctx->plonk, &ctx->cons, as, &ch, rnd);
    G1 *g = &pr.a_s;
    for (int j = 0; j < 9; j++) g1_to(o + 3 * j, g[j]);
    HF *s = &pr.a_z;
    for (int j = 0; j < 7; j++) o[27 + j] = s[j].value;
    completed = 1;
  } else {
    *status = classify_trap();
  }
  TRAP_END();
  return completed;
}

static void prove_fs_range(void *c, size_t lo, size_t hi) {
  prove_fs_ctx *x = (prove_fs_ctx *)c;
  size_t mark0 = arena_mark();
  ref_ctx_t ctx;
  ctx_build(&ctx, x->circuit, x->g1s, x->srs_len, x->g2);
  const fs_state seed = fs_seed(x->circuit, x->g1s, x->srs_len, x->g2);
  size_t mark = arena_mark();
  for (size_t i = lo; i < hi; i++) {
    HF a[4], b[4], cc[4], rnd[9];
    for (int j = 0; j < 4; j++) {
      a[j].value = x->wit[12 * i + j];
      b[j].value = x->wit[12 * i + 4 + j];
      cc[j].value = x->wit[12 * i + 8 + j];
    }
    for (int j = 0; j < 9; j++) rnd[j].value = x->rnd[9 * i + j];
    ASSIGNMENTS as; as.a = a; as.b = b; as.c = cc; as.len = 4;
    uint8_t ch6[6] = {0, 0, 0, 0, 0, 0};
    uint8_t known[6] = {0, 0, 0, 0, 0, 0};
    uint8_t rec[34];            /* the transcript's view of the PROOF record, filled stage by stage */
    uint8_t *o = x->proofs + 34 * i;
    fs_state st = seed;
    int stage = 0;              /* rounds whose challenges are fixed: 0 none, 1 beta/gamma, 2 alpha, 3 z, 4 v */
    memset(rec, 0, sizeof rec);
    for (;;) {
      uint8_t status = 0;
      const int completed = fs_pass(&ctx, &as, rnd, ch6, o, &status);
      arena_reset(mark);
      /* how far is this pass final?  Exits of rounds <= stage+1 depend only on fixed challenges. */
      static const uint8_t last_status_of_round[5] = {5, 7, 10, 10, 12};   /* rounds 1..5 (round 4 has no exit) */
      if (stage == 4) {          /* complete challenge set: this is the run */
        if (completed) {
          fs_round5(st, o, ch6);
          known[5] = 1;
          x->status[i] = 0;
        } else {
          memset(o, 0, 34);
          x->status[i] = status;
        }
        break;
      }
      if (!completed && status <= last_status_of_round[stage]) {   /* exits before the next challenge is drawn */
        memset(o, 0, 34);
        x->status[i] = status;
        break;
      }
      /* advance the transcript by one round from the captured values */
      if (stage == 0) {
        memcpy(rec, T.commit_bytes, 9);
        st = fs_round1(st, rec, ch6); known[1] = known[2] = 1;
      } else if (stage == 1) {
        memcpy(rec + 9, T.commit_bytes[3], 3);
        st = fs_round2(st, rec, ch6); known[0] = 1;
      } else if (stage == 2) {
        memcpy(rec + 12, T.commit_bytes[4], 9);
        st = fs_round3(st, rec, ch6); known[3] = 1;
      } else {
        /* the nine poly_eval calls after the seventh commitment: a b c s1 s2 t zw l1 r (plonk.h:527-574) */
        uint8_t ev[9]; int k = 0;
        for (int e = 0; e < T.n_evals && e < 32; e++) if (T.eval_commits[e] == 7 && k < 9) ev[k++] = T.eval_val[e];
        if (k != 9) { fputs("ref_driver: fs: unexpected evaluation count\n", stderr); abort(); }
        rec[27] = ev[0]; rec[28] = ev[1]; rec[29] = ev[2]; rec[30] = ev[3]; rec[31] = ev[4]; rec[32] = ev[8]; rec[33] = ev[6];
        st = fs_round4(st, rec, ch6); known[4] = 1;
      }
      stage++;
    }
    if (x->chal) for (int k = 0; k < 6; k++) x->chal[6 * i + k] = known[k] ? ch6[k] : 0xFF;
  }
  ctx_free(&ctx);
  arena_reset(mark0);
}

/* Fiat-Shamir plonk_prove over a batch: wit:[n][12], rnd:[n][9] -> proofs:[n][34], status:[n], and (optional)
 * chal:[n][6] = alpha beta gamma z v u as derived, 0xFF for a challenge the reference exits before drawing. */
void ref_plonk_prove_fs_batch(const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2,
                              const uint8_t *wit, const uint8_t *rnd, size_t n,
                              uint8_t *proofs, uint8_t *status, uint8_t *chal, int nthreads) {
  prove_fs_ctx c = {circuit, g1s, srs_len, g2, wit, rnd, proofs, status, chal};
  run_ranges(prove_fs_range, &c, n, nthreads);
}
void ref_fs_seed(const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2, uint32_t out[4]) {
  const fs_state st = fs_seed(circuit, g1s, srs_len, g2);
  memcpy(out, st.v, sizeof st.v);
}
/* the verifier's side of the transcript: all six challenges from complete PROOF records */
void ref_fs_derive(const uint32_t seed[4], const uint8_t *proofs, size_t n, uint8_t *chal6) {
  fs_state st;
  memcpy(st.v, seed, sizeof st.v);
  for (size_t i = 0; i < n; i++) fs_derive(st, proofs + 34 * i, chal6 + 6 * i);
}

/* --------------------------- verifier: NOT in the reference (plonk.h:656-659).  Parity unpinned.
 * The algorithm lives in oracle/verify_spec.inc and is instantiated here over the reference's
 * own primitives, so that its building blocks are the reference's. */
#define VS_HF HF
#define VS_G1 G1
#define VS_G2 G2
#define VS_GT GTP
#define VS_POLY POLY
#define vs_hf(v) hf_new(v)
#define vs_hf_val(x) ((x).value)
#define vs_hf_add hf_add
#define vs_hf_sub hf_sub
#define vs_hf_mul hf_mul
#define vs_hf_neg hf_neg
#define vs_hf_inv hf_inv
#define vs_hf_pow hf_pow
#define vs_g1_add(a, b) g1_add(&(a), &(b))
#define vs_g1_neg(a) g1_neg(&(a))
#define vs_g1_mul(a, s) g1_mul(&(a), (uint64_t)(s))
#define vs_g1_on_curve(a) g1_is_on_curve(&(a))
#define vs_pairing(p, q) pairing(&(p), &(q))
#define vs_gt_equal(a, b) gtp_equal(&(a), &(b))
#define vs_gt_a(g) ((g).a.value)
#define vs_gt_b(g) ((g).b.value)
#define vs_g1_from g1_from
#include "verify_spec.inc"

typedef struct {
  const uint8_t *circuit, *g1s; uint32_t srs_len; const uint8_t *g2;
  const uint8_t *proofs, *chal, *u; uint8_t *verdict, *gt;
} verify_ctx;

static void verify_setup(vs_key *key, ref_ctx_t *ctx) {
  /* preprocessed commitments = srs_eval_at_s of the interpolated selector / permutation polynomials */
  PLONK *p = &ctx->plonk;
  CONSTRAINTS *k = &ctx->cons;
  HF s1[4], s2[4], s3[4];
  copy_constraints_to_roots(p, k->c_a, 4, s1);
  copy_constraints_to_roots(p, k->c_b, 4, s2);
  copy_constraints_to_roots(p, k->c_c, 4, s3);
  const HF *vals[8] = {k->q_m, k->q_l, k->q_r, k->q_o, k->q_c, s1, s2, s3};
  G1 *dst[8] = {&key->qm, &key->ql, &key->qr, &key->qo, &key->qc, &key->s1, &key->s2, &key->s3};
  for (int i = 0; i < 8; i++) {
    POLY f = interpolate_at_h(p, vals[i], 4);
    *dst[i] = srs_eval_at_s(&p->srs, &f);
  }
  key->g1_one = p->srs.g1s[0];
  key->g2_one = p->srs.g2_1;
  key->g2_s = p->srs.g2_s;
}

static void verify_range(void *c, size_t lo, size_t hi) {
  verify_ctx *x = (verify_ctx *)c;
  size_t mark0 = arena_mark();
  ref_ctx_t ctx;
  ctx_build(&ctx, x->circuit, x->g1s, x->srs_len, x->g2);
  vs_key key;
  verify_setup(&key, &ctx);
  size_t mark = arena_mark();
  for (size_t i = lo; i < hi; i++) {
    uint8_t gt4[4];
    x->verdict[i] = vs_verify(&key, x->proofs + 34 * i, x->chal + 5 * i, x->u[i], gt4);
    if (x->gt) memcpy(x->gt + 4 * i, gt4, 4);
    arena_reset(mark);
  }
  ctx_free(&ctx);
  arena_reset(mark0);
}

/* verdict: 1 accept, 0 pairing mismatch, 2 a commitment is not on the curve, 3 an opening is not in F17.
 * gt (optional): [n][4] = lhs.a lhs.b rhs.a rhs.b */
void ref_plonk_verify_batch(const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2,
                            const uint8_t *proofs, const uint8_t *chal, const uint8_t *u, size_t n,
                            uint8_t *verdict, uint8_t *gt, int nthreads) {
  verify_ctx c = {circuit, g1s, srs_len, g2, proofs, chal, u, verdict, gt};
  run_ranges(verify_range, &c, n, nthreads);
}

/* the eight preprocessed commitments + [1]_1, for tests: out[9][3] */
void ref_verifier_key(const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2, uint8_t *out) {
  size_t mark0 = arena_mark();
  ref_ctx_t ctx;
  ctx_build(&ctx, circuit, g1s, srs_len, g2);
  vs_key key;
  verify_setup(&key, &ctx);
  G1 *src[9] = {&key.qm, &key.ql, &key.qr, &key.qo, &key.qc, &key.s1, &key.s2, &key.s3, &key.g1_one};
  for (int i = 0; i < 9; i++) g1_to(out + 3 * i, *src[i]);
  ctx_free(&ctx);
  arena_reset(mark0);
}

/* ------------------------------------------------------------------ timing helper for bench.py */
double ref_now_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
