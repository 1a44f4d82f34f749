/*
 * oracle/plonk_port.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A plain-C restatement ("port") of the reference's prove/verify path, written from the
 * reference's behaviour, not from its text: value-type polynomials with fixed capacity instead
 * of malloc'd ones, status codes instead of exit()/assert().  Each function cites the reference
 * file:line whose behaviour it restates (paths relative to /root/reference/).
 *
 * Pinning: tests/test_oracle.py (and the CPU halves of tests/parity_suite.py) check this file against (a) every golden vector the
 * reference's own tests hold, (b) the golden transcript of SURVEY.md
 * Appendix A, and (c) oracle/_ref (the unmodified reference compiled here) on seeded random
 * batches, including the per-item exit paths of Appendix B.  The verifier part is
 * "parity unpinned" (the reference has none): see verify_spec.inc.
 *
 * Output: oracle/libplonk_port.so.  Exports the same batch functions as ref_driver.c with the
 * prefix port_ instead of ref_.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include "fs_spec.inc"

typedef uint8_t fe; /* a field element, always stored reduced */

/* Optional operation counters (-DPORT_COUNT_OPS, single-threaded use): the algorithmic field-operation
 * counts of the reference's algorithm, from which bench.py derives `roofline.achieved` on the integer side
 * (SURVEY.md section 8(d)).  0 hf_mul 1 hf_add 2 hf_sub 3 hf_inv 4 gf_mul 5 gf_add 6 gf_sub 7 gf_inv (calls;
 * each is also 11 gf_mul, counted under 4). */
#ifdef PORT_COUNT_OPS
static uint64_t g_ops[8];
#define CNT(i) (g_ops[i]++)
#else
#define CNT(i) ((void)0)
#endif

/* ================================================================ F17  (src/hf.h) */
#define P17 17
static fe hf_new_(int64_t v) { int64_t t = v % P17; if (t < 0) t += P17; return (fe)t; }      /* hf.h:25-35 */
static fe hf_add_(fe a, fe b) { CNT(1); uint8_t s = (uint8_t)(a + b); if (s >= P17) s -= P17; return s; } /* hf.h:79-84 */
static fe hf_sub_(fe a, fe b) { CNT(2); int8_t d = (int8_t)a - (int8_t)b; if (d < 0) d += P17; return (fe)d; } /* hf.h:92-97 */
static fe hf_mul_(fe a, fe b) { CNT(0); return (fe)((uint16_t)a * (uint16_t)b % P17); }              /* hf.h:105-109 */
static fe hf_neg_(fe a) { return a == 0 ? 0 : (fe)(P17 - a); }                                /* hf.h:116-119 */
static fe hf_pow_(fe base, uint64_t e) {                                                      /* hf.h:127-137 */
  fe r = 1;
  while (e > 0) { if (e & 1) r = hf_mul_(r, base); base = hf_mul_(base, base); e >>= 1; }
  return r;
}
/* hf.h:145-191: table lookup, inv(0) = 0.  Restated as "the j with a*j = 1, else 0". */
static fe hf_inv_(fe a) { CNT(3); for (fe j = 1; j < P17; j++) if ((uint16_t)a * j % P17 == 1) return j; return 0; }
static fe hf_div_(fe a, fe b) { return hf_mul_(a, hf_inv_(b)); }                              /* hf.h:201-203 */

/* ================================================================ F101 (src/gf.h) */
#define P101 101
static fe gf_new_(int64_t v) { int64_t t = v % P101; if (t < 0) t += P101; return (fe)t; }    /* gf.h:24-34 */
static fe gf_add_(fe a, fe b) { CNT(5); uint16_t s = (uint16_t)(a + b); if (s >= P101) s -= P101; return (fe)s; } /* gf.h:87-93 */
static fe gf_sub_(fe a, fe b) { CNT(6); int16_t d = (int16_t)a - (int16_t)b; if (d < 0) d += P101; return (fe)d; } /* gf.h:101-107 */
static fe gf_mul_(fe a, fe b) { CNT(4); return (fe)((uint16_t)a * (uint16_t)b % P101); }              /* gf.h:115-120 */
static fe gf_neg_(fe a) { return a == 0 ? 0 : (fe)(P101 - a); }                               /* gf.h:127-132 */
static fe gf_pow_(fe base, uint64_t e) {                                                      /* gf.h:140-151 */
  fe r = 1;
  while (e > 0) { if (e & 1) r = gf_mul_(r, base); e >>= 1; base = gf_mul_(base, base); }
  return r;
}
static fe gf_inv_(fe a) { CNT(7); return gf_pow_(a, P101 - 2); }   /* gf.h:159-162: Fermat, so inv(0) = 0 */
static fe gf_div_(fe a, fe b) { return gf_mul_(a, gf_inv_(b)); }                              /* gf.h:170-172 */

/* ================================================================ polynomials over F17 (src/poly.h) */
#define PCAP 512
typedef struct { fe c[PCAP]; int len; } poly;

/* poly.h:20-38: every constructor drops trailing zeros while len > 1 */
static poly poly_make(const fe *c, int len) {
  poly p;
  while (len > 1 && c[len - 1] == 0) len--;
  p.len = len;
  for (int i = 0; i < len; i++) p.c[i] = c[i];
  return p;
}
static poly poly_const(fe v) { return poly_make(&v, 1); }                                     /* poly.h:45-53 */
static int poly_is_zero_(const poly *p) {                                                     /* poly.h:55-64 */
  for (int i = 0; i < p->len; i++) if (p->c[i]) return 0;
  return 1;
}
static fe coef(const poly *p, int i) { return i < p->len ? p->c[i] : 0; }
static poly poly_add_(const poly *a, const poly *b) {                                         /* poly.h:72-87 */
  fe t[PCAP]; int m = a->len > b->len ? a->len : b->len;
  for (int i = 0; i < m; i++) t[i] = hf_add_(coef(a, i), coef(b, i));
  return poly_make(t, m);
}
static poly poly_sub_(const poly *a, const poly *b) {                                         /* poly.h:89-104 */
  fe t[PCAP]; int m = a->len > b->len ? a->len : b->len;
  for (int i = 0; i < m; i++) t[i] = hf_sub_(coef(a, i), coef(b, i));
  return poly_make(t, m);
}
static poly poly_mul_(const poly *a, const poly *b) {                                         /* poly.h:106-122 */
  fe t[PCAP]; int m = a->len + b->len - 1;
  if (m < 0) m = 0;
  memset(t, 0, sizeof t);
  for (int i = 0; i < a->len; i++)
    for (int j = 0; j < b->len; j++) t[i + j] = hf_add_(t[i + j], hf_mul_(a->c[i], b->c[j]));
  return poly_make(t, m);
}
/* poly.h:67-70: in place on the constant term, no trimming; the caller's polynomial changes */
static void poly_add_const_inplace(poly *a, fe b) { a->c[0] = hf_add_(a->c[0], b); }
/* poly.h:124-177.  Returns 1 where the reference exits (zero divisor). */
static int poly_divmod_(const poly *num, const poly *den, poly *quot, poly *rem) {
  if (poly_is_zero_(den)) return 1;
  fe q[PCAP], r[PCAP];
  memset(q, 0, sizeof q); memset(r, 0, sizeof r);
  int nl = num->len, dl = den->len;
  for (int i = 0; i < nl; i++) r[i] = num->c[i];
  fe lead_inv = hf_inv_(den->c[dl - 1]);
  for (int i = nl - 1; i >= dl - 1; i--) {
    fe k = hf_mul_(r[i], lead_inv);
    q[i - (dl - 1)] = k;
    for (int j = 0; j < dl; j++) r[i - j] = hf_sub_(r[i - j], hf_mul_(k, den->c[dl - 1 - j]));
  }
  int ql = nl >= dl ? nl - dl + 1 : 1;
  while (ql > 1 && q[ql - 1] == 0) ql--;
  int rl = dl - 1;
  if (rl > nl) rl = nl;
  while (rl > 1 && r[rl - 1] == 0) rl--;
  *quot = poly_make(q, ql);
  *rem = poly_make(r, rl);       /* rl == 0 when the divisor is a constant: len-0 remainder (hazard C-4) */
  return 0;
}
static poly poly_scale_(const poly *p, fe k) {                                                /* poly.h:179-197 */
  if (k == 0) return poly_const(0);
  fe t[PCAP];
  for (int i = 0; i < p->len; i++) t[i] = hf_mul_(p->c[i], k);
  return poly_make(t, p->len);
}
static poly poly_shift_(const poly *p, int s) {                                               /* poly.h:199-216 */
  if (poly_is_zero_(p)) return poly_const(0);
  fe t[PCAP];
  memset(t, 0, sizeof t);
  for (int i = 0; i < p->len; i++) t[i + s] = p->c[i];
  return poly_make(t, p->len + s);
}
/* poly.h:218-238; returns 1 where the reference exits */
static int poly_slice_(const poly *p, int start, int end, poly *out) {
  if (start >= end || end > p->len) return 1;
  *out = poly_make(p->c + start, end - start);
  return 0;
}
static poly poly_negate_(const poly *p) {                                                     /* poly.h:240-254 */
  fe t[PCAP];
  for (int i = 0; i < p->len; i++) t[i] = hf_neg_(p->c[i]);
  return poly_make(t, p->len);
}
static fe poly_eval_(const poly *p, fe x) {                                                   /* poly.h:265-272 */
  fe y = 0;
  for (int i = p->len - 1; i >= 0; i--) y = hf_add_(hf_mul_(y, x), p->c[i]);
  return y;
}
static poly poly_z_(const fe *pts, int n) {                                                   /* poly.h:274-286 */
  poly acc = poly_const(1);
  for (int i = 0; i < n; i++) {
    fe t[2] = {hf_neg_(pts[i]), 1};
    poly term = poly_make(t, 2);
    acc = poly_mul_(&acc, &term);
  }
  return acc;
}
/* poly.h:288-321; returns 1 where the reference exits (duplicate x) */
static int poly_lagrange_(const fe *xs, const fe *ys, int n, poly *out) {
  poly l = poly_const(0);
  for (int j = 0; j < n; j++) {
    poly lj = poly_const(1);
    for (int i = 0; i < n; i++) {
      if (i == j) continue;
      fe dinv = hf_inv_(hf_sub_(xs[j], xs[i]));
      if (dinv == 0) return 1;
      fe t[2] = {hf_neg_(hf_mul_(dinv, xs[i])), dinv};
      poly term = poly_make(t, 2);
      lj = poly_mul_(&lj, &term);
    }
    poly s = poly_scale_(&lj, ys[j]);
    l = poly_add_(&l, &s);
  }
  *out = l;
  return 0;
}

/* ================================================================ matrices over F17 (src/matrix.h) */
#define MCAP 16
typedef struct { int m, n; fe v[MCAP * 2 * MCAP]; } matrix;   /* row-major v[col + row*n], matrix.h:48 */
static fe mget(const matrix *a, int r, int c) { return a->v[c + r * a->n]; }
static void mset(matrix *a, int r, int c, fe x) { a->v[c + r * a->n] = x; }
static matrix matrix_mul_(const matrix *a, const matrix *b) {                                 /* matrix.h:81-98 */
  matrix r; r.m = a->m; r.n = b->n;
  for (int i = 0; i < a->m; i++)
    for (int j = 0; j < b->n; j++) {
      fe s = 0;
      for (int k = 0; k < a->n; k++) s = hf_add_(s, hf_mul_(mget(a, i, k), mget(b, k, j)));
      mset(&r, i, j, s);
    }
  return r;
}
static void matrix_rref_(matrix *a) {                                                         /* matrix.h:100-149 */
  int lead = 0;
  for (int r = 0; r < a->m; r++) {
    if (a->n <= lead) return;
    int i = r;
    while (mget(a, i, lead) == 0) {
      i++;
      if (i == a->m) { i = r; lead++; if (lead == a->n) return; }
    }
    if (i != r)
      for (int k = 0; k < a->n; k++) { fe t = mget(a, i, k); mset(a, i, k, mget(a, r, k)); mset(a, r, k, t); }
    fe d = mget(a, r, lead);
    if (d != 0)
      for (int k = 0; k < a->n; k++) mset(a, r, k, hf_div_(mget(a, r, k), d));
    for (int i2 = 0; i2 < a->m; i2++) {
      if (i2 == r) continue;
      fe mult = mget(a, i2, lead);
      for (int k = 0; k < a->n; k++) mset(a, i2, k, hf_sub_(mget(a, i2, k), hf_mul_(mget(a, r, k), mult)));
    }
    lead++;
  }
}
static matrix matrix_inv_(const matrix *a) {                                                  /* matrix.h:151-176 */
  int n = a->n;
  matrix aug; aug.m = n; aug.n = 2 * n;
  memset(aug.v, 0, sizeof aug.v);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) mset(&aug, i, j, mget(a, i, j));
    mset(&aug, i, i + n, 1);
  }
  matrix_rref_(&aug);
  matrix inv; inv.m = n; inv.n = n;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) mset(&inv, i, j, mget(&aug, i, j + n));
  return inv;
}

/* ================================================================ G1 (src/g1.h): affine (x, y, infinite) */
typedef struct { fe x, y; uint8_t inf; } g1;
static g1 g1_mk(uint64_t x, uint64_t y) { g1 p = {gf_new_((int64_t)x), gf_new_((int64_t)y), 0}; return p; } /* g1.h:13-20 */
static g1 g1_id(void) { g1 p = {0, 0, 1}; return p; }                                          /* g1.h:33-35 */
static int g1_on_curve_(const g1 *p) {                                                         /* g1.h:26-31 */
  if (p->inf) return 1;
  return gf_pow_(p->y, 2) == gf_add_(gf_pow_(p->x, 3), 3);
}
static g1 g1_dbl_(const g1 *a) {                                                               /* g1.h:37-56 */
  if (a->inf || a->y == 0) return g1_id();
  fe m = gf_div_(gf_mul_(3, gf_mul_(a->x, a->x)), gf_mul_(2, a->y));
  fe m2 = gf_mul_(m, m);
  fe xr = gf_sub_(m2, gf_mul_(2, a->x));
  fe yr = gf_sub_(gf_mul_(m, gf_sub_(gf_mul_(3, a->x), m2)), a->y);
  return g1_mk(xr, yr);
}
static g1 g1_add_(const g1 *a, const g1 *b) {                                                  /* g1.h:59-83 */
  if (a->inf) return *b;
  if (b->inf) return *a;
  if (a->x == b->x) {
    if (gf_add_(a->y, b->y) == 0) return g1_id();
    return g1_dbl_(a);
  }
  fe m = gf_mul_(gf_sub_(b->y, a->y), gf_inv_(gf_sub_(b->x, a->x)));
  fe m2 = gf_mul_(m, m);
  fe xr = gf_sub_(gf_sub_(m2, a->x), b->x);
  fe yr = gf_sub_(gf_mul_(m, gf_sub_(a->x, xr)), a->y);
  return g1_mk(xr, yr);
}
static g1 g1_neg_(const g1 *a) { if (a->inf) return *a; return g1_mk(a->x, gf_neg_(a->y)); }   /* g1.h:85-89 */
static g1 g1_mul_(const g1 *p, uint64_t s) {                                                   /* g1.h:91-103 */
  g1 r = g1_id(), add = *p;
  while (s > 0) { if (s & 1) r = g1_add_(&r, &add); add = g1_dbl_(&add); s >>= 1; }
  return r;
}

/* ================================================================ G2 (src/g2.h): (x, y*u), u^2 = -2, no identity */
typedef struct { fe x, y; } g2;
static g2 g2_add_(const g2 *p, const g2 *q) {                                                  /* g2.h:32-66 */
  fe x, y;
  if (p->x == q->x && p->y == q->y) {
    fe m = gf_div_(gf_mul_(3, gf_mul_(p->x, p->x)), gf_mul_(2, p->y));
    fe m2 = gf_mul_(m, m);
    fe n2inv = gf_inv_(gf_neg_(2));
    fe w = gf_mul_(m2, n2inv);
    x = gf_sub_(w, gf_mul_(2, p->x));
    y = gf_sub_(gf_mul_(gf_mul_(n2inv, m), gf_sub_(gf_mul_(3, p->x), w)), p->y);
  } else {
    fe m = gf_div_(gf_sub_(q->y, p->y), gf_sub_(q->x, p->x));
    fe w = gf_mul_(gf_mul_(m, m), gf_neg_(2));
    x = gf_sub_(gf_sub_(w, p->x), q->x);
    y = gf_sub_(gf_mul_(m, gf_sub_(p->x, x)), p->y);
  }
  g2 r = {x, y};
  return r;
}
static g2 g2_neg_(const g2 *p) { g2 r = {p->x, gf_neg_(p->y)}; return r; }                     /* g2.h:27-30 */
/* g2.h:68-84; scalar 0 is undefined in the reference -- callers exclude it */
static g2 g2_mul_(g2 base, uint64_t s) {
  int have = 0; g2 r = {0xFF, 0xFF};
  while (s > 0) {
    if (s & 1) { if (have) r = g2_add_(&r, &base); else { r = base; have = 1; } }
    s >>= 1;
    base = g2_add_(&base, &base);
  }
  return r;
}

/* ================================================================ GT (src/gt.h): a + b u, u^2 = -2 */
typedef struct { fe a, b; } gt;
static gt gt_conj(const gt *p) { gt r = {p->a, gf_neg_(p->b)}; return r; }                     /* gt.h:19-21 */
static gt gt_mul_(const gt *x, const gt *y) {                                                  /* gt.h:23-28 */
  gt r;
  r.a = gf_sub_(gf_mul_(x->a, y->a), gf_mul_(gf_mul_(2, x->b), y->b));
  r.b = gf_add_(gf_mul_(x->a, y->b), gf_mul_(x->b, y->a));
  return r;
}
static gt gt_pow_(const gt *base, uint64_t e) {                                                /* gt.h:30-51 */
  gt p;
  if (e >= 101) { gt t = gt_pow_(base, e / 101); p = gt_conj(&t); e %= 101; }
  else { p.a = 1; p.b = 0; }
  gt cur = *base;
  while (e > 0) { if (e & 1) p = gt_mul_(&p, &cur); e >>= 1; cur = gt_mul_(&cur, &cur); }
  return p;
}

/* ================================================================ pairing (src/pairing.h) */
typedef struct { fe x, y, c; } line_eq;
static line_eq line_(const g1 *a, const g1 *b) {                                               /* pairing.h:19-29 */
  fe m = gf_sub_(b->x, a->x), n = gf_sub_(b->y, a->y);
  line_eq l = {n, gf_neg_(m), gf_sub_(gf_mul_(m, a->y), gf_mul_(n, a->x))};
  return l;
}
static gt line_at(const line_eq *l, const g2 *q) {                                             /* pairing.h:41-44,57-60 */
  gt t = {gf_add_(gf_mul_(q->x, l->x), l->c), gf_mul_(q->y, l->y)};
  return t;
}
static gt miller_(uint64_t r, const g1 *p, const g2 *q) {                                      /* pairing.h:31-64 */
  if (r == 1) { gt one = {1, 0}; return one; }
  if (r & 1) {
    g1 rp = g1_mul_(p, r - 1);
    line_eq l = line_(&rp, p);
    gt prev = miller_(r - 1, p, q);
    gt term = line_at(&l, q);
    return gt_mul_(&prev, &term);
  }
  g1 rp = g1_mul_(p, r / 2);
  g1 nrp = g1_neg_(&rp);
  g1 two_nrp = g1_mul_(&nrp, 2);
  line_eq l = line_(&rp, &two_nrp);
  gt prev = miller_(r / 2, p, q);
  gt sq = gt_mul_(&prev, &prev);
  gt term = line_at(&l, q);
  return gt_mul_(&sq, &term);
}
static gt pairing_(const g1 *p, const g2 *q) {                                                 /* pairing.h:66-83 */
  gt f = miller_(17, p, q);
  return gt_pow_(&f, (101ull * 101ull - 1ull) / 17ull);
}

/* ================================================================ KZG SRS (src/srs.h) */
#define SRS_CAP 64
typedef struct { g1 g1s[SRS_CAP]; int len; g2 g2_1, g2_s; } srs_t;
static srs_t srs_create_(fe secret, int n) {                                                   /* srs.h:18-43 */
  srs_t s; s.len = n + 1;
  g1 id = g1_id();
  fe sp = secret;                                   /* starts at secret, not 1 (srs.h:32) */
  for (int i = 0; i < s.len; i++) { s.g1s[i] = g1_mul_(&id, sp); sp = gf_mul_(sp, secret); }   /* multiples of the IDENTITY (srs.h:27,34) */
  g2 h = {36, 31};                                  /* g2.h:19-21 */
  s.g2_1 = h;
  s.g2_s = g2_mul_(h, secret);
  return s;
}
/* srs.h:53-68; returns 1 where the reference exits (poly longer than the SRS) */
static int srs_commit_(const srs_t *s, const poly *p, g1 *out) {
  if (p->len > s->len) return 1;
  g1 acc = g1_id();
  for (int i = 0; i < p->len; i++) { g1 t = g1_mul_(&s->g1s[i], p->c[i]); acc = g1_add_(&acc, &t); }
  *out = acc;
  return 0;
}

/* ================================================================ protocol (src/constraints.h, src/plonk.h) */
typedef struct { fe q_l[4], q_r[4], q_o[4], q_m[4], q_c[4]; uint8_t c_type[3][4], c_idx[3][4]; } circuit_t;
typedef struct { srs_t srs; fe h[4], k1h[4], k2h[4]; matrix vinv; poly zh; } plonk_t;

static void circuit_from(circuit_t *c, const uint8_t *b) {
  memcpy(c->q_l, b, 4); memcpy(c->q_r, b + 4, 4); memcpy(c->q_o, b + 8, 4); memcpy(c->q_m, b + 12, 4); memcpy(c->q_c, b + 16, 4);
  for (int s = 0; s < 3; s++) { memcpy(c->c_type[s], b + 20 + 8 * s, 4); memcpy(c->c_idx[s], b + 24 + 8 * s, 4); }
}
/* plonk.h:53-119 with OMEGA=4, K1=2, K2=3 (plonk.h:12-14); the coset checks cannot fire for these */
static void plonk_new_(plonk_t *p, const srs_t *srs) {
  p->srs = *srs;
  for (int i = 0; i < 4; i++) p->h[i] = hf_pow_(4, (uint64_t)i);
  for (int i = 0; i < 4; i++) { p->k1h[i] = hf_mul_(p->h[i], 2); p->k2h[i] = hf_mul_(p->h[i], 3); }
  matrix v; v.m = 4; v.n = 4;
  for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) mset(&v, r, c, hf_pow_(p->h[r], (uint64_t)c));
  p->vinv = matrix_inv_(&v);
  p->zh = poly_z_(p->h, 4);
}
/* plonk.h:142-160 (1-based index, plonk.h:144); returns 1 where the reference exits */
static int copy_to_roots_(const plonk_t *p, const uint8_t *type, const uint8_t *idx, int n, fe *sigma) {
  for (int i = 0; i < n; i++) {
    int j = idx[i] - 1;
    if (type[i] == 0) sigma[i] = p->h[j];
    else if (type[i] == 1) sigma[i] = p->k1h[j];
    else if (type[i] == 2) sigma[i] = p->k2h[j];
    else return 1;
  }
  return 0;
}
/* plonk.h:162-195: h_pows_inv (4x4) times the value column, then trimmed */
static poly interp_h_(const plonk_t *p, const fe *vals) {
  matrix col; col.m = 4; col.n = 1;
  for (int i = 0; i < 4; i++) col.v[i] = vals[i];
  matrix r = matrix_mul_(&p->vinv, &col);
  return poly_make(r.v, 4);
}
/* constraints.h:145-171 */
static int satisfies_(const circuit_t *c, const fe *a, const fe *b, const fe *w) {
  for (int i = 0; i < 4; i++) {
    fe lhs = 0;
    lhs = hf_add_(lhs, hf_mul_(c->q_l[i], a[i]));
    lhs = hf_add_(lhs, hf_mul_(c->q_r[i], b[i]));
    lhs = hf_add_(lhs, hf_mul_(c->q_o[i], w[i]));
    lhs = hf_add_(lhs, hf_mul_(c->q_m[i], hf_mul_(a[i], b[i])));
    lhs = hf_add_(lhs, c->q_c[i]);
    if (lhs != 0) return 0;
  }
  return 1;
}

static poly lin2(fe c0, fe c1) { fe t[2] = {c0, c1}; return poly_make(t, 2); }
static void put_g1(uint8_t *o, g1 p) { o[0] = p.x; o[1] = p.y; o[2] = p.inf ? 1 : 0; }

/* plonk.h:223-656.  Returns the SURVEY.md Appendix-B row of the first exit that fires (0 = done). */
/* Fiat-Shamir mode (fs != NULL; spec: fs_spec.inc): ch is ignored, each challenge is drawn from the transcript at
 * the point where the protocol fixes it, and fs->ch / fs->known report what was drawn before the first exit. */
typedef struct { fs_state st; uint8_t ch[6], known[6]; } fs_t;
static int plonk_prove_(const plonk_t *pk, const circuit_t *cs, const fe *wa, const fe *wb, const fe *wc,
                        const fe ch[5], const fe rnd[9], uint8_t proof[34], fs_t *fs) {
  const int n = 4;
  if (!satisfies_(cs, wa, wb, wc)) return 1;                                    /* plonk.h:231 */
  fe alpha = 0, beta = 0, gamma = 0, z = 0, v = 0;
  if (!fs) { alpha = ch[0]; beta = ch[1]; gamma = ch[2]; z = ch[3]; v = ch[4]; }
  const fe omega = 4, k1 = 2, k2 = 3;
  fe sg1[4], sg2[4], sg3[4];
  if (copy_to_roots_(pk, cs->c_type[0], cs->c_idx[0], n, sg1)) return 3;        /* plonk.h:254-256 */
  if (copy_to_roots_(pk, cs->c_type[1], cs->c_idx[1], n, sg2)) return 3;
  if (copy_to_roots_(pk, cs->c_type[2], cs->c_idx[2], n, sg3)) return 3;
  poly fa = interp_h_(pk, wa), fb = interp_h_(pk, wb), fc = interp_h_(pk, wc);  /* plonk.h:265-275 */
  poly qo = interp_h_(pk, cs->q_o), qm = interp_h_(pk, cs->q_m), ql = interp_h_(pk, cs->q_l);
  poly qr = interp_h_(pk, cs->q_r), qc = interp_h_(pk, cs->q_c);
  poly s1 = interp_h_(pk, sg1), s2 = interp_h_(pk, sg2), s3 = interp_h_(pk, sg3);

  /* round 1 (plonk.h:279-301) */
  poly bl, t;
  bl = lin2(rnd[1], rnd[0]); t = poly_mul_(&bl, &pk->zh); poly a = poly_add_(&t, &fa);
  bl = lin2(rnd[3], rnd[2]); t = poly_mul_(&bl, &pk->zh); poly b = poly_add_(&t, &fb);
  bl = lin2(rnd[5], rnd[4]); t = poly_mul_(&bl, &pk->zh); poly c = poly_add_(&t, &fc);
  g1 a_s, b_s, c_s;
  if (srs_commit_(&pk->srs, &a, &a_s)) return 5;
  if (srs_commit_(&pk->srs, &b, &b_s)) return 5;
  if (srs_commit_(&pk->srs, &c, &c_s)) return 5;
  if (fs) {
    put_g1(proof, a_s); put_g1(proof + 3, b_s); put_g1(proof + 6, c_s);
    fs->st = fs_round1(fs->st, proof, fs->ch);
    beta = fs->ch[1]; gamma = fs->ch[2]; fs->known[1] = fs->known[2] = 1;
  }

  /* round 2 (plonk.h:320-379) */
  fe acc[4];
  acc[0] = 1;
  for (int i = 1; i < n; i++) {
    fe w = hf_pow_(omega, (uint64_t)(i - 1));
    fe den = hf_mul_(hf_mul_(hf_add_(wa[i - 1], hf_add_(hf_mul_(beta, w), gamma)),
                             hf_add_(wb[i - 1], hf_add_(hf_mul_(beta, hf_mul_(k1, w)), gamma))),
                     hf_add_(wc[i - 1], hf_add_(hf_mul_(beta, hf_mul_(k2, w)), gamma)));
    fe e1 = poly_eval_(&s1, w), e2 = poly_eval_(&s2, w), e3 = poly_eval_(&s3, w);
    fe num = hf_mul_(hf_mul_(hf_add_(wa[i - 1], hf_add_(hf_mul_(beta, e1), gamma)),
                             hf_add_(wb[i - 1], hf_add_(hf_mul_(beta, e2), gamma))),
                     hf_add_(wc[i - 1], hf_add_(hf_mul_(beta, e3), gamma)));
    acc[i] = hf_mul_(acc[i - 1], hf_div_(den, num));     /* den/num with 1/0 = 0 (plonk.h:357) */
  }
  poly accx = interp_h_(pk, acc);
  if (poly_eval_(&accx, hf_pow_(omega, (uint64_t)n)) != 1) return 6;            /* plonk.h:366-368 */
  fe zb[3] = {rnd[8], rnd[7], rnd[6]};
  bl = poly_make(zb, 3); t = poly_mul_(&bl, &pk->zh);
  poly zx = poly_add_(&t, &accx);
  g1 z_s;
  if (srs_commit_(&pk->srs, &zx, &z_s)) return 7;
  if (fs) {
    put_g1(proof + 9, z_s);
    fs->st = fs_round2(fs->st, proof, fs->ch);
    alpha = fs->ch[0]; fs->known[0] = 1;
  }

  /* round 3 (plonk.h:385-524) */
  fe lv[4] = {1, 0, 0, 0};
  poly l1 = interp_h_(pk, lv);
  poly pi = poly_const(0);
  poly ab = poly_mul_(&a, &b), abm = poly_mul_(&ab, &qm);
  poly al = poly_mul_(&a, &ql), br = poly_mul_(&b, &qr), co = poly_mul_(&c, &qo);
  poly u1 = poly_add_(&abm, &al), u2 = poly_add_(&br, &co);
  poly t1 = poly_add_(&u1, &u2);
  t1 = poly_add_(&t1, &pi);
  t1 = poly_add_(&t1, &qc);

  poly g;
  g = lin2(gamma, beta);                 poly xa = poly_add_(&a, &g);  xa = poly_scale_(&xa, alpha);
  g = lin2(gamma, hf_mul_(beta, k1));    poly xb = poly_add_(&b, &g);
  g = lin2(gamma, hf_mul_(beta, k2));    poly xc = poly_add_(&c, &g);
  poly t2 = poly_mul_(&xa, &xb); t2 = poly_mul_(&t2, &xc); t2 = poly_mul_(&t2, &zx);

  poly w;
  w = poly_scale_(&s1, beta); poly ya = poly_add_(&a, &w); poly_add_const_inplace(&ya, gamma); ya = poly_scale_(&ya, alpha);
  w = poly_scale_(&s2, beta); poly yb = poly_add_(&b, &w); poly_add_const_inplace(&yb, gamma);
  w = poly_scale_(&s3, beta); poly yc = poly_add_(&c, &w); poly_add_const_inplace(&yc, gamma);
  fe zw[PCAP];
  for (int i = 0; i < zx.len; i++) zw[i] = hf_mul_(zx.c[i], hf_pow_(omega, (uint64_t)i));
  poly zwx = poly_make(zw, zx.len);                                             /* z(omega x), plonk.h:466-470 */
  poly t3 = poly_mul_(&ya, &yb); t3 = poly_mul_(&t3, &yc); t3 = poly_mul_(&t3, &zwx);

  poly m1 = poly_const(hf_neg_(1));
  poly zm1 = poly_add_(&zx, &m1);
  zm1 = poly_scale_(&zm1, hf_pow_(alpha, 2));
  poly t4 = poly_mul_(&zm1, &l1);

  poly tn = poly_add_(&t1, &t2);
  tn = poly_sub_(&tn, &t3);
  tn = poly_add_(&tn, &t4);
  poly tx, rem;
  if (poly_divmod_(&tn, &pk->zh, &tx, &rem)) return 2;
  if (!poly_is_zero_(&rem)) return 8;                                           /* plonk.h:507-510 */
  poly tlo, tmid, thi;
  if (poly_slice_(&tx, 0, n + 2, &tlo)) return 9;                               /* plonk.h:517-519 */
  if (poly_slice_(&tx, n + 2, 2 * (n + 2), &tmid)) return 9;
  if (poly_slice_(&tx, 2 * (n + 2), tx.len, &thi)) return 9;
  g1 tlo_s, tmid_s, thi_s;
  if (srs_commit_(&pk->srs, &tlo, &tlo_s)) return 10;
  if (srs_commit_(&pk->srs, &tmid, &tmid_s)) return 10;
  if (srs_commit_(&pk->srs, &thi, &thi_s)) return 10;
  if (fs) {
    put_g1(proof + 12, tlo_s); put_g1(proof + 15, tmid_s); put_g1(proof + 18, thi_s);
    fs->st = fs_round3(fs->st, proof, fs->ch);
    z = fs->ch[3]; fs->known[3] = 1;
  }

  /* round 4 (plonk.h:527-574) */
  fe a_z = poly_eval_(&a, z), b_z = poly_eval_(&b, z), c_z = poly_eval_(&c, z);
  fe s1_z = poly_eval_(&s1, z), s2_z = poly_eval_(&s2, z);
  fe t_z = poly_eval_(&tx, z), zw_z = poly_eval_(&zwx, z);
  poly r1, rr;
  r1 = poly_scale_(&qm, hf_mul_(a_z, b_z));
  rr = poly_scale_(&ql, a_z); r1 = poly_add_(&r1, &rr);
  rr = poly_scale_(&qr, b_z); r1 = poly_add_(&r1, &rr);
  rr = poly_scale_(&qo, c_z); r1 = poly_add_(&r1, &rr);                         /* q_C is left out (hazard C-9) */
  fe e_a = hf_add_(hf_add_(a_z, hf_mul_(beta, z)), gamma);
  fe e_b = hf_add_(hf_add_(b_z, hf_mul_(hf_mul_(beta, k1), z)), gamma);
  fe e_c = hf_add_(hf_add_(c_z, hf_mul_(hf_mul_(beta, k2), z)), gamma);
  poly r2 = poly_scale_(&zx, hf_mul_(hf_mul_(hf_mul_(e_a, e_b), e_c), alpha));
  poly s3s = poly_scale_(&s3, hf_mul_(beta, zw_z));
  fe f_a = hf_add_(a_z, hf_add_(hf_mul_(beta, s1_z), gamma));
  fe f_b = hf_add_(b_z, hf_add_(hf_mul_(beta, s2_z), gamma));
  poly r3 = poly_mul_(&zx, &s3s);                                               /* times the POLYNOMIAL z(x), and added */
  r3 = poly_scale_(&r3, hf_mul_(hf_mul_(f_a, f_b), alpha));
  poly r4 = poly_scale_(&zx, hf_mul_(poly_eval_(&l1, z), hf_pow_(alpha, 2)));
  poly rx = poly_add_(&r1, &r2);
  rx = poly_add_(&rx, &r3);
  rx = poly_add_(&rx, &r4);
  fe r_z = poly_eval_(&rx, z);
  if (fs) {
    fe sc4[7] = {a_z, b_z, c_z, s1_z, s2_z, r_z, zw_z};
    memcpy(proof + 27, sc4, 7);
    fs->st = fs_round4(fs->st, proof, fs->ch);
    v = fs->ch[4]; fs->known[4] = 1;
  }

  /* round 5 (plonk.h:582-621) */
  poly tm = poly_scale_(&tmid, hf_pow_(z, (uint64_t)(n + 2)));
  poly th = poly_scale_(&thi, hf_pow_(z, (uint64_t)(2 * n + 4)));
  poly wz = poly_add_(&tlo, &tm);
  wz = poly_add_(&wz, &th);
  poly_add_const_inplace(&wz, hf_neg_(t_z));
  poly q;
  poly_add_const_inplace(&rx, hf_neg_(r_z));  q = poly_scale_(&rx, v);                 poly o1 = q;
  poly_add_const_inplace(&a, hf_neg_(a_z));   q = poly_scale_(&a, hf_pow_(v, 2));      poly o2 = q;
  poly_add_const_inplace(&b, hf_neg_(b_z));   q = poly_scale_(&b, hf_pow_(v, 3));      poly o3 = q;
  poly_add_const_inplace(&c, hf_neg_(c_z));   q = poly_scale_(&c, hf_pow_(v, 4));      poly o4 = q;
  poly_add_const_inplace(&s1, hf_neg_(s1_z)); q = poly_scale_(&s1, hf_pow_(v, 5));     poly o5 = q;
  poly_add_const_inplace(&s2, hf_neg_(s2_z)); q = poly_scale_(&s2, hf_pow_(v, 6));     poly o6 = q;
  wz = poly_add_(&wz, &o1); wz = poly_add_(&wz, &o2); wz = poly_add_(&wz, &o3);
  wz = poly_add_(&wz, &o4); wz = poly_add_(&wz, &o5); wz = poly_add_(&wz, &o6);
  poly d1 = lin2(hf_neg_(z), 1), wzq, rem1;
  if (poly_divmod_(&wz, &d1, &wzq, &rem1)) return 2;
  if (!poly_is_zero_(&rem1)) return 11;                                         /* plonk.h:610 */
  poly_add_const_inplace(&zx, hf_neg_(zw_z));
  poly d2 = lin2(hf_mul_(hf_neg_(z), omega), 1), wzwq, rem2;
  if (poly_divmod_(&zx, &d2, &wzwq, &rem2)) return 2;
  if (!poly_is_zero_(&rem2)) return 11;                                         /* plonk.h:617 */
  g1 wz_s, wzw_s;
  if (srs_commit_(&pk->srs, &wzq, &wz_s)) return 12;                            /* plonk.h:620-621 */
  if (srs_commit_(&pk->srs, &wzwq, &wzw_s)) return 12;

  g1 pts[9] = {a_s, b_s, c_s, z_s, tlo_s, tmid_s, thi_s, wz_s, wzw_s};          /* PROOF field order, plonk.h:24-41 */
  for (int i = 0; i < 9; i++) put_g1(proof + 3 * i, pts[i]);
  fe sc[7] = {a_z, b_z, c_z, s1_z, s2_z, r_z, zw_z};
  memcpy(proof + 27, sc, 7);
  if (fs) { fs->st = fs_round5(fs->st, proof, fs->ch); fs->known[5] = 1; }
  return 0;
}

/* ================================================================ exported batch functions */
typedef void (*range_fn)(void *ctx, size_t lo, size_t hi);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->ctx, j->lo, j->hi); return NULL; }
static void run_ranges(range_fn fn, void *ctx, size_t n, int nthreads) {
  if (nthreads <= 1 || n < 2) { fn(ctx, 0, n); return; }
  if ((size_t)nthreads > n) nthreads = (int)n;
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
  job_t *jobs = malloc(sizeof(job_t) * (size_t)nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t].fn = fn; jobs[t].ctx = ctx;
    jobs[t].lo = n * (size_t)t / (size_t)nthreads;
    jobs[t].hi = n * (size_t)(t + 1) / (size_t)nthreads;
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

static g1 g1_from(const uint8_t *b) { g1 p = {b[0], b[1], b[2] != 0}; return p; }
static g2 g2_from(const uint8_t *b) { g2 p = {b[0], b[1]}; return p; }
static gt gt_from(const uint8_t *b) { gt p = {b[0], b[1]}; return p; }
static void poly_out(uint8_t *c, size_t stride, uint8_t *len, const poly *p) {
  memset(c, 0, stride);
  for (int i = 0; i < p->len && (size_t)i < stride; i++) c[i] = p->c[i];
  *len = (uint8_t)p->len;
}
static srs_t srs_from(const uint8_t *g1s, uint32_t len, const uint8_t *g2b) {
  srs_t s; s.len = (int)len;
  for (uint32_t i = 0; i < len && i < SRS_CAP; i++) s.g1s[i] = g1_from(g1s + 3 * i);
  s.g2_1 = g2_from(g2b); s.g2_s = g2_from(g2b + 2);
  return s;
}

int port_abi_version(void) { return 1; }
#ifdef PORT_COUNT_OPS
void port_ops_reset(void) { memset(g_ops, 0, sizeof g_ops); }
void port_ops_get(uint64_t out[8]) { memcpy(out, g_ops, sizeof g_ops); }
#endif

void port_field_op(int field, int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    fe x = a[i], y = b ? b[i] : 0, r;
    if (field == 17) {
      switch (op) {
        case 0: r = hf_add_(x, y); break;  case 1: r = hf_sub_(x, y); break;
        case 2: r = hf_mul_(x, y); break;  case 3: r = hf_div_(x, y); break;
        case 4: r = hf_neg_(x); break;     case 5: r = hf_inv_(x); break;
        default: r = hf_pow_(x, y); break;
      }
    } else {
      switch (op) {
        case 0: r = gf_add_(x, y); break;  case 1: r = gf_sub_(x, y); break;
        case 2: r = gf_mul_(x, y); break;  case 3: r = gf_div_(x, y); break;
        case 4: r = gf_neg_(x); break;     case 5: r = gf_inv_(x); break;
        default: r = gf_pow_(x, y); break;
      }
    }
    out[i] = r;
  }
}
uint8_t port_hf_new(int64_t v) { return hf_new_(v); }
uint8_t port_gf_new(int64_t v) { return gf_new_(v); }

void port_poly_binop(int op, const uint8_t *a, const uint8_t *alen, size_t sa, const uint8_t *b, const uint8_t *blen, size_t sb,
                     uint8_t *out, uint8_t *olen, size_t so, size_t n) {
  for (size_t i = 0; i < n; i++) {
    poly pa = poly_make(a + i * sa, alen[i]), pb = poly_make(b + i * sb, blen[i]);
    poly r = op == 0 ? poly_add_(&pa, &pb) : op == 1 ? poly_sub_(&pa, &pb) : poly_mul_(&pa, &pb);
    poly_out(out + i * so, so, olen + i, &r);
  }
}
void port_poly_divide(const uint8_t *num, const uint8_t *nlen, size_t sn, const uint8_t *den, const uint8_t *dlen, size_t sd,
                      uint8_t *quot, uint8_t *qlen, size_t sq, uint8_t *rem, uint8_t *rlen, size_t sr, uint8_t *status, size_t n) {
  for (size_t i = 0; i < n; i++) {
    poly pn = poly_make(num + i * sn, nlen[i]), pd = poly_make(den + i * sd, dlen[i]), q, r;
    status[i] = (uint8_t)poly_divmod_(&pn, &pd, &q, &r);
    if (status[i]) { memset(quot + i * sq, 0, sq); memset(rem + i * sr, 0, sr); qlen[i] = rlen[i] = 0; continue; }
    poly_out(quot + i * sq, sq, qlen + i, &q);
    poly_out(rem + i * sr, sr, rlen + i, &r);
  }
}
void port_poly_eval(const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *x, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) { poly pp = poly_make(p + i * sp, plen[i]); out[i] = poly_eval_(&pp, x[i]); }
}
void port_poly_unop(int op, const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *k, uint8_t *out, uint8_t *olen, size_t so, size_t n) {
  for (size_t i = 0; i < n; i++) {
    poly pp = poly_make(p + i * sp, plen[i]), r;
    fe kv = k ? k[i] : 0;
    if (op == 0) r = poly_scale_(&pp, kv);
    else if (op == 1) r = poly_negate_(&pp);
    else if (op == 2) r = poly_shift_(&pp, kv);
    else { poly_add_const_inplace(&pp, kv); r = pp; }
    poly_out(out + i * so, so, olen + i, &r);
  }
}
void port_poly_slice(const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *start, const uint8_t *end,
                     uint8_t *out, uint8_t *olen, size_t so, uint8_t *status, size_t n) {
  for (size_t i = 0; i < n; i++) {
    poly pp = poly_make(p + i * sp, plen[i]), r;
    status[i] = (uint8_t)poly_slice_(&pp, start[i], end[i], &r);
    if (status[i]) { memset(out + i * so, 0, so); olen[i] = 0; continue; }
    poly_out(out + i * so, so, olen + i, &r);
  }
}
void port_poly_z(const uint8_t *points, size_t len, uint8_t *out, uint8_t *olen, size_t so) {
  poly r = poly_z_(points, (int)len);
  poly_out(out, so, olen, &r);
}
void port_poly_lagrange(const uint8_t *xs, const uint8_t *ys, size_t len, size_t n, uint8_t *out, uint8_t *olen, size_t so, uint8_t *status) {
  for (size_t i = 0; i < n; i++) {
    poly r;
    status[i] = (uint8_t)poly_lagrange_(xs + i * len, ys + i * len, (int)len, &r);
    if (status[i]) { memset(out + i * so, 0, so); olen[i] = 0; continue; }
    poly_out(out + i * so, so, olen + i, &r);
  }
}
void port_matrix_mul(const uint8_t *a, size_t am, size_t an, const uint8_t *b, size_t bm, size_t bn, uint8_t *out) {
  matrix A, B; A.m = (int)am; A.n = (int)an; B.m = (int)bm; B.n = (int)bn;
  memcpy(A.v, a, am * an); memcpy(B.v, b, bm * bn);
  matrix R = matrix_mul_(&A, &B);
  memcpy(out, R.v, am * bn);
}
void port_matrix_inv(const uint8_t *a, size_t n, uint8_t *out) {
  matrix A; A.m = A.n = (int)n; memcpy(A.v, a, n * n);
  matrix R = matrix_inv_(&A);
  memcpy(out, R.v, n * n);
}
void port_matrix_gauss_jordan(uint8_t *a, size_t m, size_t n) {
  matrix A; A.m = (int)m; A.n = (int)n; memcpy(A.v, a, m * n);
  matrix_rref_(&A);
  memcpy(a, A.v, m * n);
}

void port_g1_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    g1 x = g1_from(a + 3 * i), r;
    if (op == 0) { g1 y = g1_from(b + 3 * i); r = g1_add_(&x, &y); }
    else if (op == 1) r = g1_dbl_(&x);
    else r = g1_neg_(&x);
    put_g1(out + 3 * i, r);
  }
}
typedef struct { const uint8_t *p; const uint64_t *s; uint8_t *out; } g1mul_ctx;
static void g1mul_range(void *c, size_t lo, size_t hi) {
  g1mul_ctx *x = c;
  for (size_t i = lo; i < hi; i++) { g1 p = g1_from(x->p + 3 * i); put_g1(x->out + 3 * i, g1_mul_(&p, x->s[i])); }
}
void port_g1_mul(const uint8_t *p, const uint64_t *scalars, uint8_t *out, size_t n, int nthreads) {
  g1mul_ctx c = {p, scalars, out};
  run_ranges(g1mul_range, &c, n, nthreads);
}
void port_g1_is_on_curve(const uint8_t *p, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) { g1 x = g1_from(p + 3 * i); out[i] = (uint8_t)g1_on_curve_(&x); }
}
void port_g2_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    g2 x = g2_from(a + 2 * i), r;
    if (op == 0) { g2 y = g2_from(b + 2 * i); r = g2_add_(&x, &y); } else r = g2_neg_(&x);
    out[2 * i] = r.x; out[2 * i + 1] = r.y;
  }
}
void port_g2_mul(const uint8_t *p, const uint64_t *scalars, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) { g2 r = g2_mul_(g2_from(p + 2 * i), scalars[i]); out[2 * i] = r.x; out[2 * i + 1] = r.y; }
}
void port_gtp_mul(const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) { gt x = gt_from(a + 2 * i), y = gt_from(b + 2 * i), r = gt_mul_(&x, &y); out[2 * i] = r.a; out[2 * i + 1] = r.b; }
}
void port_gtp_pow(const uint8_t *a, const uint64_t *e, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) { gt x = gt_from(a + 2 * i), r = gt_pow_(&x, e[i]); out[2 * i] = r.a; out[2 * i + 1] = r.b; }
}
void port_srs_create(uint8_t secret, uint32_t n, uint8_t *g1s_out, uint8_t *g2_out) {
  srs_t s = srs_create_(secret, (int)n);
  for (int i = 0; i < s.len; i++) put_g1(g1s_out + 3 * i, s.g1s[i]);
  g2_out[0] = s.g2_1.x; g2_out[1] = s.g2_1.y; g2_out[2] = s.g2_s.x; g2_out[3] = s.g2_s.y;
}
typedef struct { srs_t srs; const uint8_t *polys, *plen; size_t sp; uint8_t *out, *status; } commit_ctx;
static void commit_range(void *c, size_t lo, size_t hi) {
  commit_ctx *x = c;
  for (size_t i = lo; i < hi; i++) {
    poly p = poly_make(x->polys + i * x->sp, x->plen[i]);
    g1 r;
    x->status[i] = (uint8_t)srs_commit_(&x->srs, &p, &r);
    if (x->status[i]) memset(x->out + 3 * i, 0, 3); else put_g1(x->out + 3 * i, r);
  }
}
void port_srs_eval_at_s(const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2b, const uint8_t *polys, const uint8_t *plen, size_t sp,
                        uint8_t *out, uint8_t *status, size_t n, int nthreads) {
  commit_ctx c; c.srs = srs_from(g1s, srs_len, g2b); c.polys = polys; c.plen = plen; c.sp = sp; c.out = out; c.status = status;
  run_ranges(commit_range, &c, n, nthreads);
}
void port_line(const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    g1 x = g1_from(a + 3 * i), y = g1_from(b + 3 * i);
    line_eq l = line_(&x, &y);
    out[3 * i] = l.x; out[3 * i + 1] = l.y; out[3 * i + 2] = l.c;
  }
}
typedef struct { const uint8_t *p, *q; uint8_t *out; uint64_t r; } pair_ctx;
static void pair_range(void *c, size_t lo, size_t hi) {
  pair_ctx *x = c;
  for (size_t i = lo; i < hi; i++) {
    g1 p = g1_from(x->p + 3 * i); g2 q = g2_from(x->q + 2 * i);
    gt r = x->r ? miller_(x->r, &p, &q) : pairing_(&p, &q);
    x->out[2 * i] = r.a; x->out[2 * i + 1] = r.b;
  }
}
void port_pairing(const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n, int nthreads) {
  pair_ctx c = {p, q, out, 0};
  run_ranges(pair_range, &c, n, nthreads);
}
void port_pairing_f(uint64_t r, const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n) {
  pair_ctx c = {p, q, out, r};
  pair_range(&c, 0, n);
}

static void default_plonk(plonk_t *pk) { srs_t s = srs_create_(2, 6); plonk_new_(pk, &s); }
void port_plonk_setup_dump(uint8_t *out) {
  plonk_t pk; default_plonk(&pk);
  memcpy(out, pk.h, 4); memcpy(out + 4, pk.k1h, 4); memcpy(out + 8, pk.k2h, 4);
  memcpy(out + 12, pk.vinv.v, 16);
  memset(out + 28, 0, 8);
  memcpy(out + 28, pk.zh.c, (size_t)pk.zh.len);
  out[36] = (uint8_t)pk.zh.len;
}
void port_copy_constraints_to_roots(const uint8_t *types, const uint8_t *idx, size_t len, uint8_t *sigma) {
  plonk_t pk; default_plonk(&pk);
  copy_to_roots_(&pk, types, idx, (int)len, sigma);
}
void port_interpolate_at_h(const uint8_t *vals, uint8_t *out, uint8_t *olen, size_t n) {
  plonk_t pk; default_plonk(&pk);
  for (size_t i = 0; i < n; i++) { poly r = interp_h_(&pk, vals + 4 * i); poly_out(out + 4 * i, 4, olen + i, &r); }
}

typedef struct {
  plonk_t pk; circuit_t cs;
  const uint8_t *wit, *rnd, *chal; uint8_t *proofs, *status;
  int fs_mode; fs_state fs_seed; uint8_t *chal_out;
} prove_ctx;
static void prove_range(void *c, size_t lo, size_t hi) {
  prove_ctx *x = c;
  for (size_t i = lo; i < hi; i++) {
    const uint8_t *w = x->wit + 12 * i;
    uint8_t *o = x->proofs + 34 * i;
    int st;
    if (x->fs_mode) {
      fs_t fs;
      memset(&fs, 0, sizeof fs);
      fs.st = x->fs_seed;
      st = plonk_prove_(&x->pk, &x->cs, w, w + 4, w + 8, NULL, x->rnd + 9 * i, o, &fs);
      if (x->chal_out) for (int k = 0; k < 6; k++) x->chal_out[6 * i + k] = fs.known[k] ? fs.ch[k] : 0xFF;
    } else {
      st = plonk_prove_(&x->pk, &x->cs, w, w + 4, w + 8, x->chal + 5 * i, x->rnd + 9 * i, o, NULL);
    }
    if (st) memset(o, 0, 34);
    x->status[i] = (uint8_t)st;
  }
}
void port_plonk_prove_batch(const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2b,
                            const uint8_t *wit, const uint8_t *rnd, const uint8_t *chal, size_t n,
                            uint8_t *proofs, uint8_t *status, int nthreads) {
  prove_ctx *c = malloc(sizeof *c);
  srs_t s = srs_from(g1s, srs_len, g2b);
  plonk_new_(&c->pk, &s);
  circuit_from(&c->cs, circuit);
  c->wit = wit; c->rnd = rnd; c->chal = chal; c->proofs = proofs; c->status = status;
  c->fs_mode = 0; memset(&c->fs_seed, 0, sizeof c->fs_seed); c->chal_out = NULL;
  run_ranges(prove_range, c, n, nthreads);
  free(c);
}
/* Fiat-Shamir mode (spec: fs_spec.inc): chal_out (optional) [n][6] = alpha beta gamma z v u, 0xFF where not drawn */
void port_plonk_prove_fs_batch(const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2b,
                               const uint8_t *wit, const uint8_t *rnd, size_t n,
                               uint8_t *proofs, uint8_t *status, uint8_t *chal_out, int nthreads) {
  prove_ctx *c = malloc(sizeof *c);
  srs_t s = srs_from(g1s, srs_len, g2b);
  plonk_new_(&c->pk, &s);
  circuit_from(&c->cs, circuit);
  c->wit = wit; c->rnd = rnd; c->chal = NULL; c->proofs = proofs; c->status = status;
  c->fs_mode = 1; c->fs_seed = fs_seed(circuit, g1s, srs_len, g2b); c->chal_out = chal_out;
  run_ranges(prove_range, c, n, nthreads);
  free(c);
}
void port_fs_seed(const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2b, uint32_t out[4]) {
  const fs_state st = fs_seed(circuit, g1s, srs_len, g2b);
  memcpy(out, st.v, sizeof st.v);
}
void port_fs_derive(const uint32_t seed[4], const uint8_t *proofs, size_t n, uint8_t *chal6) {
  fs_state st;
  memcpy(st.v, seed, sizeof st.v);
  for (size_t i = 0; i < n; i++) fs_derive(st, proofs + 34 * i, chal6 + 6 * i);
}

/* ---------------- verifier (parity unpinned: see verify_spec.inc) over the restated primitives */
#define VS_HF fe
#define VS_G1 g1
#define VS_G2 g2
#define VS_GT gt
#define vs_hf(v) hf_new_((int64_t)(v))
#define vs_hf_val(x) (x)
#define vs_hf_add hf_add_
#define vs_hf_sub hf_sub_
#define vs_hf_mul hf_mul_
#define vs_hf_neg hf_neg_
#define vs_hf_inv hf_inv_
#define vs_hf_pow hf_pow_
#define vs_g1_add(a, b) g1_add_(&(a), &(b))
#define vs_g1_neg(a) g1_neg_(&(a))
#define vs_g1_mul(a, s) g1_mul_(&(a), (uint64_t)(s))
#define vs_g1_on_curve(a) g1_on_curve_(&(a))
#define vs_pairing(p, q) pairing_(&(p), &(q))
#define vs_gt_equal(x, y) ((x).a == (y).a && (x).b == (y).b)                                   /* pairing.h:9-11 */
#define vs_gt_a(g) ((g).a)
#define vs_gt_b(g) ((g).b)
#define vs_g1_from g1_from
#include "verify_spec.inc"

static void verify_key_(vs_key *k, const plonk_t *pk, const circuit_t *cs) {
  fe sg[3][4];
  for (int s = 0; s < 3; s++) copy_to_roots_(pk, cs->c_type[s], cs->c_idx[s], 4, sg[s]);
  const fe *vals[8] = {cs->q_m, cs->q_l, cs->q_r, cs->q_o, cs->q_c, sg[0], sg[1], sg[2]};
  g1 *dst[8] = {&k->qm, &k->ql, &k->qr, &k->qo, &k->qc, &k->s1, &k->s2, &k->s3};
  for (int i = 0; i < 8; i++) { poly f = interp_h_(pk, vals[i]); srs_commit_(&pk->srs, &f, dst[i]); }
  k->g1_one = pk->srs.g1s[0];
  k->g2_one = pk->srs.g2_1;
  k->g2_s = pk->srs.g2_s;
}
typedef struct { vs_key key; const uint8_t *proofs, *chal, *u; uint8_t *verdict, *gt; } verify_ctx;
static void verify_range(void *c, size_t lo, size_t hi) {
  verify_ctx *x = c;
  for (size_t i = lo; i < hi; i++) {
    uint8_t g4[4];
    x->verdict[i] = vs_verify(&x->key, x->proofs + 34 * i, x->chal + 5 * i, x->u[i], g4);
    if (x->gt) memcpy(x->gt + 4 * i, g4, 4);
  }
}
void port_plonk_verify_batch(const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2b,
                             const uint8_t *proofs, const uint8_t *chal, const uint8_t *u, size_t n,
                             uint8_t *verdict, uint8_t *gtout, int nthreads) {
  plonk_t *pk = malloc(sizeof *pk);
  circuit_t cs;
  srs_t s = srs_from(g1s, srs_len, g2b);
  plonk_new_(pk, &s);
  circuit_from(&cs, circuit);
  verify_ctx c;
  verify_key_(&c.key, pk, &cs);
  c.proofs = proofs; c.chal = chal; c.u = u; c.verdict = verdict; c.gt = gtout;
  run_ranges(verify_range, &c, n, nthreads);
  free(pk);
}
void port_verifier_key(const uint8_t *circuit, const uint8_t *g1s, uint32_t srs_len, const uint8_t *g2b, uint8_t *out) {
  plonk_t *pk = malloc(sizeof *pk);
  circuit_t cs;
  srs_t s = srs_from(g1s, srs_len, g2b);
  plonk_new_(pk, &s);
  circuit_from(&cs, circuit);
  vs_key k;
  verify_key_(&k, pk, &cs);
  g1 *src[9] = {&k.qm, &k.ql, &k.qr, &k.qo, &k.qc, &k.s1, &k.s2, &k.s3, &k.g1_one};
  for (int i = 0; i < 9; i++) put_g1(out + 3 * i, *src[i]);
  free(pk);
}
