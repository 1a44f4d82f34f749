// curve.cuh -- G1 (src/g1.h), G2 (src/g2.h), GT (src/gt.h) and the pairing (src/pairing.h) as
// device functions over 32-bit registers.
//
// Bit-exactness is with the reference's *structs*, not only with the mathematics: a G1 value is the
// triple (x, y, infinite) and every function below returns exactly the triple the reference
// returns, for ANY input bytes -- off-curve points, identities with non-zero coordinates,
// x-collisions of unrelated points (hazards 6-8, 10-12).
#pragma once
#include "field.cuh"

namespace pb {

struct G1 { uint32_t x, y, inf; };
struct G2 { uint32_t x, y; };
struct GT { uint32_t a, b; };
struct Line { uint32_t x, y, c; };

PB_HD G1 g1_identity() { return G1{0u, 0u, 1u}; }  // g1.h:33-35

// g1.h:37-56.  Tangent slope m = 3x^2 / 2y (1/0 = 0 cannot occur: y == 0 returns the identity).
PB_HD G1 g1_double(const FieldTables& t, G1 a) {
  uint32_t m = red101(red101(3u * a.x * a.x) * inv101(t, red101(2u * a.y)));
  uint32_t m2 = m * m;                                   // raw < 101^2
  uint32_t xr = red101(m2 + 2u * P101 - 2u * a.x);
  uint32_t yr = red101(m * (a.x + P101 - xr) + P101 - a.y);
  G1 r{xr, yr, 0u};
  if (a.inf || a.y == 0u) r = g1_identity();
  return r;
}

// g1.h:59-83.  One slope computation serves both the chord and the tangent branch: when the x
// coordinates collide and the points are not mutual inverses the reference doubles `a` (ignoring
// b), and x_r = m^2 - x_a - x_b equals the tangent formula's m^2 - 2 x_a because x_b = x_a.
PB_HD G1 g1_add(const FieldTables& t, G1 a, G1 b) {
  bool same_x = a.x == b.x;
  uint32_t sum_y = a.y + b.y;
  bool to_id = same_x && (sum_y == 0u || sum_y == P101 || a.y == 0u);   // P + (-P), or doubling a 2-torsion point
  uint32_t num = same_x ? red101(3u * a.x * a.x) : (b.y + P101 - a.y);  // chord numerator kept raw (< 202)
  uint32_t den = same_x ? red101(2u * a.y) : sub101(b.x, a.x);
  uint32_t m = red101(num * inv101(t, den));
  uint32_t xr = red101(m * m + 2u * P101 - a.x - b.x);
  uint32_t yr = red101(m * (a.x + P101 - xr) + P101 - a.y);
  G1 r{xr, yr, 0u};
  if (to_id) r = g1_identity();
  if (b.inf) r = a;
  if (a.inf) r = b;
  return r;
}

// Leaner variants for the fast paths, where every operand is a canonically encoded point of E(F_101) (coordinates < 101,
// infinite in {0, 1}, identity = {0, 0, 1}) -- checked before they are used (verifier.cuh step 1, context creation).
// Same results as g1_add / g1_double on that domain; the difference is bookkeeping: the slope's denominator indexes the
// inverse table unreduced (2y <= 200, x_b + 101 - x_a <= 201; the table repeats mod 101) and the tangent numerator
// 3x^2 goes into the product raw (3 * 100^2 * 100 < 2^26).
PB_HD G1 g1_double_c(const FieldTables& t, G1 a) {
  const uint32_t m = red101(3u * a.x * a.x * inv101(t, 2u * a.y));
  const uint32_t xr = red101(m * m + 2u * P101 - 2u * a.x);
  const uint32_t yr = red101(m * (a.x + P101 - xr) + P101 - a.y);
  G1 r{xr, yr, 0u};
  if (a.inf || a.y == 0u) r = g1_identity();
  return r;
}
PB_HD G1 g1_add_c(const FieldTables& t, G1 a, G1 b) {
  const bool same_x = a.x == b.x;
  const uint32_t sum_y = a.y + b.y;
  // canonical curve points with equal x are equal or mutually inverse, so a.y == 0 implies b.y == 0: the reference's third
  // test (doubling a 2-torsion point, g1.h:38) is covered by sum_y == 0 on this domain
  const bool to_id = same_x && (sum_y == 0u || sum_y == P101);
  const uint32_t num = same_x ? 3u * a.x * a.x : (b.y + P101 - a.y);
  const uint32_t den = same_x ? 2u * a.y : (b.x + P101 - a.x);
  const uint32_t m = red101(num * inv101(t, den));
  const uint32_t xr = red101(m * m + 2u * P101 - a.x - b.x);
  const uint32_t yr = red101(m * (a.x + P101 - xr) + P101 - a.y);
  G1 r{xr, yr, 0u};
  if (to_id) r = g1_identity();
  if (b.inf) r = a;
  if (a.inf) r = b;
  return r;
}

PB_HD G1 g1_neg(G1 a) {  // g1.h:85-89: the identity comes back untouched
  G1 r{a.x, neg101(a.y), 0u};
  if (a.inf) r = a;
  return r;
}

// g1.h:91-103: LSB-first double-and-add over the raw 64-bit scalar (never reduced mod 17).
PB_HD G1 g1_mul(const FieldTables& t, G1 p, uint64_t s) {
  G1 r = g1_identity();
  G1 added = p;
  while (s) {
    if (s & 1ull) r = g1_add(t, r, added);
    added = g1_double(t, added);
    s >>= 1;
  }
  return r;
}

PB_HD bool g1_is_on_curve(G1 p) {  // g1.h:26-31
  uint32_t lhs = red101(p.y * p.y);
  uint32_t rhs = red101(red101(p.x * p.x) * p.x + 3u);
  return p.inf || lhs == rhs;
}

// g2.h:32-66.  No identity, no P = -Q handling: a zero denominator silently gives slope 0.
PB_HD G2 g2_add(const FieldTables& t, G2 p, G2 q) {
  constexpr uint32_t NEG2 = 99u;      // -2
  constexpr uint32_t NEG2_INV = 50u;  // (-2)^-1 = 99^99 mod 101
  uint32_t x, y;
  if (p.x == q.x && p.y == q.y) {
    uint32_t m = red101(red101(3u * p.x * p.x) * inv101(t, red101(2u * p.y)));
    uint32_t w = red101(red101(m * m) * NEG2_INV);
    x = red101(w + 2u * P101 - 2u * p.x);
    uint32_t k = red101(NEG2_INV * m);
    y = red101(k * red101(3u * p.x + P101 - w) + P101 - p.y);
  } else {
    uint32_t m = red101((q.y + P101 - p.y) * inv101(t, sub101(q.x, p.x)));
    uint32_t w = red101(red101(m * m) * NEG2);
    x = red101(w + 2u * P101 - p.x - q.x);
    y = red101(m * (p.x + P101 - x) + P101 - p.y);
  }
  return G2{x, y};
}
PB_HD G2 g2_neg(G2 p) { return G2{p.x, neg101(p.y)}; }  // g2.h:27-30

// g2.h:68-84.  Scalar 0 is undefined behaviour in the reference; here it yields (0xFF, 0xFF).
PB_HD G2 g2_mul(const FieldTables& t, G2 base, uint64_t s) {
  G2 r{0xFFu, 0xFFu};
  bool have = false;
  while (s) {
    if (s & 1ull) {
      if (have) r = g2_add(t, r, base);
      else { r = base; have = true; }
    }
    s >>= 1;
    base = g2_add(t, base, base);
  }
  return r;
}

// gt.h:23-28: (a + b u)(c + d u) with u^2 = -2
PB_HD GT gt_mul(GT x, GT y) {
  uint32_t a = red101(x.a * y.a + 2u * P101 * P101 - 2u * x.b * y.b);
  uint32_t b = red101(x.a * y.b + x.b * y.a);
  return GT{a, b};
}
PB_HD GT gt_sqr(GT x) {
  return GT{red101(x.a * x.a + 2u * P101 * P101 - 2u * x.b * x.b), red101(2u * x.a * x.b)};
}
PB_HD GT gt_conj(GT x) { return GT{x.a, neg101(x.b)}; }  // gtp_neg, gt.h:19-21

// gt.h:30-51 for arbitrary exponents (generic entry point): the Frobenius split at 101 is part of
// the function's *value* only through commutative ring arithmetic, so it is evaluated iteratively:
// exp = sum d_k 101^k  ->  prod conj^k(base)^{d_k}; conj is applied once per level as the reference
// nests it.
PB_HD GT gt_pow(GT base, uint64_t e) {
  // digits of e in base 101, most significant first, as the recursion unwinds
  uint32_t digits[10];
  int nd = 0;
  do { digits[nd++] = (uint32_t)(e % 101ull); e /= 101ull; } while (e);
  GT p{1u, 0u};
  for (int k = nd - 1; k >= 0; --k) {
    if (k != nd - 1) p = gt_conj(p);
    uint32_t d = digits[k];
    GT cur = base;
    while (d) { if (d & 1u) p = gt_mul(p, cur); d >>= 1; cur = gt_sqr(cur); }
  }
  return p;
}

// pairing.h:19-29
PB_HD Line line_through(G1 a, G1 b) {
  uint32_t m = sub101(b.x, a.x), n = sub101(b.y, a.y);
  return Line{n, neg101(m), red101(m * a.y + P101 * P101 - n * a.x)};
}
// pairing.h:41-44 / 57-60: the line evaluated at Q = (qx, qy u)
PB_HD GT line_at(Line l, G2 q) { return GT{red101(q.x * l.x + l.c), red101(q.y * l.y)}; }

// pairing.h:31-64 for general r (generic entry point; the recursion is unrolled into a bit scan
// from the top bit of r).  Values, not redundancy, are reproduced: k*P is carried along instead of
// being recomputed by g1_mul at every level; g1_mul(p, k) is still the
// reference's double-and-add, evaluated exactly.
PB_HD GT miller(const FieldTables& t, uint64_t r, G1 p, G2 q) {
  // walk r's recursion: r -> r-1 (odd) or r/2 (even) until 1; replay it backwards
  uint8_t ops[128];
  int n = 0;
  uint64_t k = r;
  while (k > 1) {
    if (k & 1ull) { ops[n++] = 1; k -= 1; } else { ops[n++] = 0; k >>= 1; }
  }
  GT f{1u, 0u};
  k = 1;
  for (int i = n - 1; i >= 0; --i) {
    if (ops[i]) {           // f_{k+1} = f_k * l_{kP, P}(Q)
      G1 kp = g1_mul(t, p, k);
      f = gt_mul(f, line_at(line_through(kp, p), q));
      k += 1;
    } else {                // f_{2k} = f_k^2 * l_{kP, 2(-kP)}(Q)
      G1 kp = g1_mul(t, p, k);
      G1 two_neg = g1_mul(t, g1_neg(kp), 2);
      f = gt_mul(gt_mul(f, f), line_at(line_through(kp, two_neg), q));
      k <<= 1;
    }
  }
  return f;
}

// pairing.h:66-83 specialised to r = 17, exponent (101^2 - 1)/17 = 600 -- the hot path.
// 17 -> 16 -> 8 -> 4 -> 2 -> 1: g1_mul(p, 2^j) is j doublings of p (the accumulator starts at the
// identity, whose addition returns the doubled point unchanged), and g1_mul(-kP, 2) = double(-kP)
// = -(2kP) as triples, so one doubling chain P, 2P, 4P, 8P, 16P feeds all five line functions.
// Final exponentiation: f^600 = conj(f^5) * f^95 (gt.h:32-38 splits at 101), evaluated with shared
// squarings in the commutative ring F_101[u]/(u^2+2).
// f_17,P(Q) before the final exponentiation
PB_HD GT miller17(const FieldTables& t, G1 p, G2 q) {
  G1 p2 = g1_double(t, p), p4 = g1_double(t, p2), p8 = g1_double(t, p4), p16 = g1_double(t, p8);
  GT f = line_at(line_through(p, g1_neg(p2)), q);                    // f_2  (f_1 = 1)
  f = gt_mul(gt_sqr(f), line_at(line_through(p2, g1_neg(p4)), q));   // f_4
  f = gt_mul(gt_sqr(f), line_at(line_through(p4, g1_neg(p8)), q));   // f_8
  f = gt_mul(gt_sqr(f), line_at(line_through(p8, g1_neg(p16)), q));  // f_16
  f = gt_mul(f, line_at(line_through(p16, p), q));                   // f_17
  return f;
}
// the same for a canonically encoded point of E(F_101) (fast verifier): the doubling chain without the any-input bookkeeping
// line through canonical points a, b evaluated at Q, with the differences kept unreduced (<= 201; same residues as
// line_through + line_at): slope parts m = x_b - x_a, n = y_b - y_a, value (q.x n + m y_a - n x_a) + (-m q.y) u
PB_HD GT line_eval_c(G1 a, G1 b, G2 q) {
  const uint32_t m = b.x + P101 - a.x, n = b.y + P101 - a.y;
  const uint32_t c = m * a.y + 200u * P101 - n * a.x;                 // >= 0: n x_a <= 201 * 100
  return GT{red101(q.x * n + c), red101(q.y * (2u * P101 - m))};
}
PB_HD GT miller17_c(const FieldTables& t, G1 p, G2 q) {
  G1 p2 = g1_double_c(t, p), p4 = g1_double_c(t, p2), p8 = g1_double_c(t, p4), p16 = g1_double_c(t, p8);
  GT f = line_eval_c(p, g1_neg(p2), q);
  f = gt_mul(gt_sqr(f), line_eval_c(p2, g1_neg(p4), q));
  f = gt_mul(gt_sqr(f), line_eval_c(p4, g1_neg(p8), q));
  f = gt_mul(gt_sqr(f), line_eval_c(p8, g1_neg(p16), q));
  f = gt_mul(f, line_eval_c(p16, p, q));
  return f;
}
PB_HD GT final_exp600(GT f) {
  GT f2 = gt_sqr(f), f4 = gt_sqr(f2), f8 = gt_sqr(f4), f16 = gt_sqr(f8), f32 = gt_sqr(f16), f64 = gt_sqr(f32);
  GT f5 = gt_mul(f4, f);
  GT f95 = gt_mul(gt_mul(gt_mul(f64, f16), gt_mul(f8, f4)), gt_mul(f2, f));
  return gt_mul(gt_conj(f5), f95);
}
PB_HD GT pairing17(const FieldTables& t, G1 p, G2 q) { return final_exp600(miller17(t, p, q)); }

// gtp_equal(pairing(p1, q1), pairing(p2, q2)) with ONE final exponentiation.  GT = F_101[u]/(u^2 + 2) is a field (-2 is
// a non-residue mod 101), so for f2 != 0:  f1^600 == f2^600  <=>  (f1 / f2)^600 == 1, and 1/f2 = conj(f2) / N(f2) with
// N(f2) in F_101^*, whose 600th power is (N^100)^6 = 1 -- hence (f1 conj(f2))^600 == 1.  A zero Miller value (the
// reference's pairing of the identity is (0, 0)) stays zero under the power: equal iff both are zero.
PB_HD bool pairings_equal17_c(const FieldTables& t, G1 p1, G2 q1, G1 p2, G2 q2) {
  const GT f1 = miller17_c(t, p1, q1), f2 = miller17_c(t, p2, q2);
  const bool z1 = (f1.a | f1.b) == 0u, z2 = (f2.a | f2.b) == 0u;
  const GT e = final_exp600(gt_mul(f1, gt_conj(f2)));
  const bool one = e.a == 1u && e.b == 0u;
  return (z1 || z2) ? (z1 && z2) : one;
}

}  // namespace pb
