/* tests/c/dropin_check.c -- acceptance program for the drop-in headers (include/*.h + libplonk_b200.so).
 * It is written against the REFERENCE's API exactly as a plonk.c user would write it (one translation unit
 * including plonk.h and pairing.h) and checks the known answers the reference's own tests assert
 * (g1-test.c, g2-test.c, gt-test.c, pairing-test.c, poly-test.c, srs-test.c, plonk-test.c) plus the golden
 * transcript of SURVEY.md Appendix A.  Needs a GPU: every arithmetic call below runs a CUDA kernel. */
#include <assert.h>
#include <stdio.h>
#include <string.h>
#include "plonk.h"
#include "pairing.h"

static int failures = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
  } while (0)

static int poly_is(const POLY *p, const int *want, size_t n) {
  if (p->len != n) return 0;
  for (size_t i = 0; i < n; i++)
    if (p->coeffs[i].value != (uint8_t)want[i]) return 0;
  return 1;
}
static int g1_is(G1 p, int x, int y, int inf) { return p.x.value == x && p.y.value == y && (int)p.infinite == inf; }

static void check_poly(void) {
  HF a[] = {f17(5), f17(0), f17(10), f17(6)}, b[] = {f17(1), f17(2), f17(4)};
  POLY pa = poly_new(a, 4), pb = poly_new(b, 3);
  POLY m = poly_mul(&pa, &pb);                         /* poly-test.c:101-115 */
  int want_m[] = {5, 10, 13, 9, 1, 7};
  CHECK(poly_is(&m, want_m, 6));
  POLY s = poly_add(&pa, &pb), d = poly_sub(&pa, &pb);
  int want_s[] = {6, 2, 14, 6}, want_d[] = {4, 15, 6, 6};
  CHECK(poly_is(&s, want_s, 4) && poly_is(&d, want_d, 4));
  /* (x-3)(x-5) / (x-3) = x-5 (poly-test.c:148-169) */
  HF r3[] = {f17(-3), f17(1)}, r5[] = {f17(-5), f17(1)};
  POLY x3 = poly_new(r3, 2), x5 = poly_new(r5, 2), prod = poly_mul(&x3, &x5), q, rem;
  poly_divide(&prod, &x3, &q, &rem);
  int want_q[] = {12, 1};
  CHECK(poly_is(&q, want_q, 2) && poly_is_zero(&rem) && rem.len == 1);
  CHECK(poly_eval(&prod, f17(3)).value == 0 && poly_eval(&prod, f17(4)).value == 16);
  HF pts[] = {f17(1), f17(5)};
  POLY z = poly_z(pts, 2);                             /* poly-test.c:180-189 */
  int want_z[] = {5, 11, 1};
  CHECK(poly_is(&z, want_z, 3));
  HF xs[] = {f17(1), f17(4), f17(16), f17(13)}, ys[] = {f17(3), f17(4), f17(0), f17(0)};
  POLY l = poly_lagrange(xs, ys, 4);                   /* agrees with interpolate_at_h on H: 6 + x + 4x^2 + 9x^3 */
  int want_l[] = {6, 1, 4, 9};
  CHECK(poly_is(&l, want_l, 4));
  POLY sc = poly_scale(&pa, f17(0));                   /* scalar 0 -> [0] (hazard C-5) */
  CHECK(sc.len == 1 && sc.coeffs[0].value == 0);
  POLY alias = poly_add_hf(&pa, f17(12));              /* in place, alias (hazard C-3) */
  CHECK(alias.coeffs == pa.coeffs && pa.coeffs[0].value == 0 && pa.len == 4);
  POLY sl = poly_slice(&m, 1, 4), ng = poly_negate(&pb), sh = poly_shift(&pb, 2);
  int want_sl[] = {10, 13, 9}, want_ng[] = {16, 15, 13}, want_sh[] = {0, 0, 1, 2, 4};
  CHECK(poly_is(&sl, want_sl, 3) && poly_is(&ng, want_ng, 3) && poly_is(&sh, want_sh, 5));
  /* constant divisor -> remainder of length 0 (hazard C-4) */
  HF two[] = {f17(2)};
  POLY c2 = poly_new(two, 1), q2, r2;
  poly_divide(&pb, &c2, &q2, &r2);
  CHECK(r2.len == 0 && q2.len == 3 && q2.coeffs[0].value == 9);
  POLY all[] = {pa, pb, m, s, d, x3, x5, prod, q, rem, z, l, sc, sl, ng, sh, c2, q2, r2};
  for (size_t i = 0; i < sizeof all / sizeof all[0]; i++) poly_free(&all[i]);
}

static void check_matrix(void) {
  HF v[] = {f17(1), f17(1), f17(1), f17(1), f17(1), f17(4), f17(16), f17(13), f17(1), f17(16), f17(1), f17(16), f17(1), f17(13), f17(16), f17(4)};
  MATRIX V = matrix_new(v, 4, 4), inv = matrix_inv(&V), id = matrix_mul(&V, &inv), sum = matrix_add(&V, &V);
  int want[] = {13, 13, 13, 13, 13, 16, 4, 1, 13, 4, 13, 4, 13, 1, 4, 16};     /* plonk-test.c:38-40 */
  for (int i = 0; i < 16; i++) CHECK(inv.v[i].value == want[i]);
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) CHECK(matrix_get(&id, r, c).value == (r == c));
  CHECK(matrix_get(&sum, 1, 1).value == 8);
  HF g[] = {f17(2), f17(4), f17(1), f17(0), f17(0), f17(3)};
  MATRIX G = matrix_new(g, 2, 3);
  matrix_gauss_jordan(&G);
  CHECK(G.v[0].value == 1 && G.v[1].value == 2 && G.v[2].value == 0 && G.v[5].value == 1);
  matrix_free(&V); matrix_free(&inv); matrix_free(&id); matrix_free(&sum); matrix_free(&G);
}

static void check_groups(void) {
  G1 g = g1_generator(), two = g1_add(&g, &g), three = g1_add(&two, &g), four = g1_double(&two), eight = g1_add(&four, &four);
  CHECK(g1_is(two, 68, 74, 0) && g1_is(three, 26, 45, 0) && g1_is(four, 65, 98, 0) && g1_is(eight, 18, 49, 0));   /* g1-test.c:26-41 */
  G1 sixteen = g1_add(&eight, &eight), neg = g1_neg(&g), zero = g1_add(&g, &neg), six = g1_mul(&g, 6);
  G1 five = g1_add(&four, &g), six2 = g1_add(&five, &g);
  CHECK(g1_is(sixteen, 1, 99, 0) && g1_is(neg, 1, 99, 0) && g1_is(zero, 0, 0, 1) && six.x.value == six2.x.value && six.y.value == six2.y.value);
  G1 big = g1_mul(&g, 17ull * 1000003ull + 5);         /* raw 64-bit scalars, never reduced mod 17 */
  CHECK(g1_is(big, 12, 32, 0) && g1_is_on_curve(&big));
  G1 off = g1_new(5, 7);
  CHECK(!g1_is_on_curve(&off));
  G2 h = g2_generator(), h2 = g2_add(&h, &h), h3 = g2_add(&h2, &h), h4 = g2_add(&h2, &h2), h4b = g2_add(&h3, &h), h6 = g2_mul(h, 6), h6b = g2_add(&h4, &h2);
  CHECK(h2.x.value == 90 && h2.y.value == 82 && h4.x.value == h4b.x.value && h4.y.value == h4b.y.value && h6.x.value == h6b.x.value && h6.y.value == h6b.y.value);
  GTP a = gtp_new(f101(26), f101(97)), b = gtp_new(f101(93), f101(76)), ab = gtp_mul(&a, &b);                      /* gt-test.c:11-26 */
  CHECK(ab.a.value == 97 && ab.b.value == 89);
  GTP p6 = gtp_new(f101(42), f101(49)), r6 = gtp_pow(&p6, 6), conj = gtp_pow(&b, 101), c600 = gtp_new(f101(68), f101(47)), r600 = gtp_pow(&c600, 600);
  CHECK(r6.a.value == 97 && r6.b.value == 89 && conj.a.value == 93 && conj.b.value == 25 && r600.a.value == 97 && r600.b.value == 89);
  /* pairing-test.c:5-27 */
  G1 r = g1_mul(&g, 4), p5 = g1_mul(&g, 5), pr = g1_add(&g, &r);
  G2 q = g2_mul(h, 3), q5 = g2_mul(q, 5);
  GTP left = pairing(&p5, &q), right = pairing(&g, &q5), pq = pairing(&g, &q), pq5 = gtp_pow(&pq, 5), rq = pairing(&r, &q), prod = gtp_mul(&pq, &rq), sum = pairing(&pr, &q);
  CHECK(gtp_equal(&left, &right) && gtp_equal(&left, &pq5) && gtp_equal(&sum, &prod));
  GTP e = pairing(&g, &h), f = pairing_f(17, &g, &h);                                                                /* SURVEY.md A.3 */
  G1 id = g1_identity();
  GTP e0 = pairing(&id, &h);
  CHECK(e.a.value == 7 && e.b.value == 28 && f.a.value == 15 && f.b.value == 26 && e0.a.value == 0 && e0.b.value == 0);
  LINE_EQ l = line(&g, &two);
  CHECK(l.x.value == 72 && l.y.value == 34 && l.c.value == 62);
}

static void fill_test_circuit(CONSTRAINTS *k, ASSIGNMENTS *as) {
  static HF ql[4], qr[4], qo[4], qm[4], qc[4], a[4], b[4], c[4];
  static COPY_OF ca[4], cb[4], cc[4];
  int vql[] = {0, 0, 0, 1}, vqm[] = {1, 1, 1, 0}, va[] = {3, 4, 5, 9}, vb[] = {3, 4, 5, 16}, vc[] = {9, 16, 25, 25};
  for (int i = 0; i < 4; i++) {
    ql[i] = f17(vql[i]); qr[i] = f17(vql[i]); qo[i] = f17(-1); qm[i] = f17(vqm[i]); qc[i] = f17(0);
    a[i] = f17(va[i]); b[i] = f17(vb[i]); c[i] = f17(vc[i]);
  }
  COPY_OF t_ca[4] = {{COPYOF_B, 1}, {COPYOF_B, 2}, {COPYOF_B, 3}, {COPYOF_C, 1}};
  COPY_OF t_cb[4] = {{COPYOF_A, 1}, {COPYOF_A, 2}, {COPYOF_A, 3}, {COPYOF_C, 2}};
  COPY_OF t_cc[4] = {{COPYOF_A, 4}, {COPYOF_B, 4}, {COPYOF_C, 4}, {COPYOF_C, 3}};
  memcpy(ca, t_ca, sizeof ca); memcpy(cb, t_cb, sizeof cb); memcpy(cc, t_cc, sizeof cc);
  k->q_l = ql; k->q_r = qr; k->q_o = qo; k->q_m = qm; k->q_c = qc; k->num_gates = 4;
  k->c_a = ca; k->c_b = cb; k->c_c = cc; k->num_constraints = 4;
  as->a = a; as->b = b; as->c = c; as->len = 4;
}

static void check_protocol(void) {
  SRS srs = srs_create(f101(2), 6);                                             /* srs-test.c:14-17: the degenerate SRS */
  CHECK(srs.len == 7 && g1_is(srs.g1s[0], 0, 0, 1) && g1_is(srs.g1s[6], 0, 0, 1) && srs.g2_s.x.value == 90 && srs.g2_s.y.value == 82);
  PLONK pk = plonk_new(srs, 4);
  int h[] = {1, 4, 16, 13}, k1[] = {2, 8, 15, 9}, k2[] = {3, 12, 14, 5}, zh[] = {16, 0, 0, 0, 1};
  for (int i = 0; i < 4; i++) CHECK(pk.h[i].value == h[i] && pk.k1_h[i].value == k1[i] && pk.k2_h[i].value == k2[i]);
  CHECK(poly_is(&pk.z_h_x, zh, 5) && pk.h_pows_inv.v[5].value == 16);
  HF vec[] = {f17(3), f17(4), f17(0), f17(0)};
  POLY ip = interpolate_at_h(&pk, vec, 4);                                      /* plonk-test.c:52-57 */
  int want_ip[] = {6, 1, 4, 9};
  CHECK(poly_is(&ip, want_ip, 4));
  CONSTRAINTS k; ASSIGNMENTS as;
  fill_test_circuit(&k, &as);
  HF sig[4];
  copy_constraints_to_roots(&pk, k.c_a, 4, sig);                                /* plonk-test.c:105-112 */
  CHECK(sig[0].value == 2 && sig[1].value == 8 && sig[2].value == 15 && sig[3].value == 3);
  CHECK(constraints_satisfy(&k, &as));
  HF rnd[9] = {f17(7), f17(4), f17(11), f17(12), f17(16), f17(2), f17(14), f17(11), f17(7)};
  CHALLENGE ch = {f17(15), f17(12), f17(13), f17(5), f17(12)};
  PROOF pr = plonk_prove(&pk, &k, &as, &ch, rnd);                               /* SURVEY.md Appendix A.1 */
  CHECK(g1_is(pr.a_s, 0, 0, 1) && g1_is(pr.w_z_omega_s, 0, 0, 1));
  CHECK(pr.a_z.value == 15 && pr.b_z.value == 13 && pr.c_z.value == 5 && pr.s_sigma_1_z.value == 1 && pr.s_sigma_2_z.value == 12 &&
        pr.r_z.value == 15 && pr.z_omega_z.value == 15);
  /* Appendix A.2: a generator SRS handed over through the public struct */
  G1 g = g1_generator();
  for (size_t i = 0; i < pk.srs.len; i++) pk.srs.g1s[i] = g1_mul(&g, 1ull << i);
  HF pc[] = {f17(1), f17(2), f17(3), f17(4), f17(5), f17(6)};
  POLY p6 = poly_new(pc, 6);
  G1 cm = srs_eval_at_s(&pk.srs, &p6);
  CHECK(g1_is(cm, 68, 27, 0));
  PROOF p2 = plonk_prove(&pk, &k, &as, &ch, rnd);
  CHECK(g1_is(p2.a_s, 91, 66, 0) && g1_is(p2.b_s, 26, 45, 0) && g1_is(p2.c_s, 91, 35, 0) && g1_is(p2.z_s, 32, 59, 0) && g1_is(p2.t_lo_s, 12, 32, 0) &&
        g1_is(p2.t_mid_s, 26, 45, 0) && g1_is(p2.t_hi_s, 91, 66, 0) && g1_is(p2.w_z_s, 91, 35, 0) && g1_is(p2.w_z_omega_s, 65, 98, 0));
  CHECK(p2.r_z.value == 15 && sizeof(PROOF) == 34);
  poly_free(&ip); poly_free(&p6);
  plonk_free(&pk);
}

int main(void) {
  check_poly();
  check_matrix();
  check_groups();
  check_protocol();
  printf(failures ? "dropin_check: %d FAILURES\n" : "dropin_check: all checks passed\n", failures);
  return failures ? 1 : 0;
}
