// verifier.cuh -- plonk_verify.  NOT in the reference (plonk.h ends at plonk_prove, plonk.h:656-659):
// parity is unpinned against the reference and pinned against oracle/verify_spec.inc, the textbook
// PLONK verifier stated over the reference's primitives.  This file follows that specification step
// by step and performs the same group operations in the same order, each of them the exact
// emulation of g1_mul / g1_add / g1_neg / pairing from curve.cuh.
#pragma once
#include <type_traits>
#include "curve.cuh"
#include "transcript.cuh"
#include "prover.cuh"   // pack_g1 / unpack_g1

namespace pb {

// Preprocessed verifier key, passed by value (uniform, constant bank).
struct VerifyKey {
  G1 qm, ql, qr, qo, qc, s1, s2, s3;   // srs_eval_at_s of the interpolated selector / permutation polynomials
  G1 g1_one;                           // srs.g1s[0]
  G2 g2_one, g2_s;                     // srs.g2_1, srs.g2_s
  FsState fs_seed;                     // Fiat-Shamir mode only: the transcript state after absorbing circuit and SRS (transcript.cuh)
};

struct VerifyOut {
  uint32_t verdict;   // 1 accept, 0 pairing mismatch, 2 bad commitment encoding / off curve, 3 bad scalar byte
  GT lhs, rhs;
};

// proof: 34 bytes already unpacked into registers (27 commitment bytes, 7 openings)
PB_HD void verify_one(const VerifyKey& k, const FieldTables& ft, const uint32_t (&pb)[27], const uint32_t (&op)[7],
                     const uint32_t (&ch)[5], uint32_t u, VerifyOut& out) {
  out.lhs = GT{0u, 0u};
  out.rhs = GT{0u, 0u};
  // step 1: encodings and curve membership (y^2 = x^3 + 3, g1.h:26-31)
  bool bad_pt = false;
  G1 P[9];
#pragma unroll
  for (int j = 0; j < 9; j++) {
    uint32_t x = pb[3 * j], y = pb[3 * j + 1], f = pb[3 * j + 2];
    bad_pt |= x > 100u || y > 100u || f > 1u || (f == 1u && (x | y) != 0u);
    P[j] = G1{x > 100u ? 0u : x, y > 100u ? 0u : y, f != 0u ? 1u : 0u};   // clamp so that table look-ups stay in range
    bad_pt |= !g1_is_on_curve(P[j]);
  }
  // step 2: openings and challenges are field elements
  bool bad_sc = u > 16u;
#pragma unroll
  for (int j = 0; j < 7; j++) bad_sc |= op[j] > 16u;
#pragma unroll
  for (int j = 0; j < 5; j++) bad_sc |= ch[j] > 16u;
  if (bad_pt) { out.verdict = 2u; return; }
  if (bad_sc) { out.verdict = 3u; return; }

  const uint32_t a_z = op[0], b_z = op[1], c_z = op[2], s1_z = op[3], s2_z = op[4], r_z = op[5], zw_z = op[6];
  const uint32_t alpha = ch[0], beta = ch[1], gamma = ch[2], z = ch[3], v = ch[4];
  constexpr uint32_t K1 = 2u, K2 = 3u, OMEGA = 4u;

  // steps 4-6
  const uint32_t z2 = red17(z * z), z3 = red17(z2 * z), z4 = red17(z2 * z2);
  const uint32_t zh_z = sub17(z4, 1u);
  const uint32_t l1_z = red17(13u * (1u + z + z2 + z3));
  const uint32_t alpha2 = red17(alpha * alpha);
  // step 7
  const uint32_t pa = red17(a_z + beta * s1_z + gamma), pbb = red17(b_z + beta * s2_z + gamma);
  const uint32_t pab = red17(pa * pbb);
  const uint32_t perm = red17(red17(red17(pab * red17(c_z + gamma)) * zw_z) * alpha);
  const uint32_t t_num = red17(r_z + 2u * P17 - perm - red17(l1_z * alpha2));
  const uint32_t t_z = red17(t_num * inv17(ft, zh_z));
  // step 8 scalars
  const uint32_t bz = red17(beta * z);
  const uint32_t ga = red17(a_z + bz + gamma), gb = red17(b_z + K1 * bz + gamma), gc = red17(c_z + K2 * bz + gamma);
  const uint32_t d_z = red17(red17(red17(red17(red17(ga * gb) * gc) * alpha) * v) + red17(red17(l1_z * alpha2) * v) + u);
  const uint32_t d_s3 = red17(red17(red17(red17(pab * alpha) * v) * beta) * zw_z);

  G1 D = g1_add(ft, g1_mul(ft, k.qm, red17(red17(a_z * b_z) * v)), g1_mul(ft, k.ql, red17(a_z * v)));
  D = g1_add(ft, D, g1_mul(ft, k.qr, red17(b_z * v)));
  D = g1_add(ft, D, g1_mul(ft, k.qo, red17(c_z * v)));
  D = g1_add(ft, D, g1_mul(ft, k.qc, v));
  D = g1_add(ft, D, g1_mul(ft, P[3], d_z));
  D = g1_add(ft, D, g1_neg(g1_mul(ft, k.s3, d_s3)));

  // step 9
  const uint32_t v2 = red17(v * v), v3 = red17(v2 * v), v4 = red17(v3 * v), v5 = red17(v4 * v), v6 = red17(v5 * v);
  const uint32_t z6 = red17(z4 * z2), z12 = red17(z6 * z6);
  G1 F = g1_add(ft, P[4], g1_mul(ft, P[5], z6));
  F = g1_add(ft, F, g1_mul(ft, P[6], z12));
  F = g1_add(ft, F, D);
  F = g1_add(ft, F, g1_mul(ft, P[0], v2));
  F = g1_add(ft, F, g1_mul(ft, P[1], v3));
  F = g1_add(ft, F, g1_mul(ft, P[2], v4));
  F = g1_add(ft, F, g1_mul(ft, k.s1, v5));
  F = g1_add(ft, F, g1_mul(ft, k.s2, v6));

  // step 10
  const uint32_t e = red17(t_z + v * r_z + v2 * a_z + v3 * b_z + v4 * c_z + v5 * s1_z + v6 * s2_z + u * zw_z);
  G1 E = g1_mul(ft, k.g1_one, e);

  // step 11
  G1 lhs_p = g1_add(ft, P[7], g1_mul(ft, P[8], u));
  G1 rhs_p = g1_add(ft, g1_mul(ft, P[7], z), g1_mul(ft, P[8], red17(red17(u * z) * OMEGA)));
  rhs_p = g1_add(ft, rhs_p, F);
  rhs_p = g1_add(ft, rhs_p, g1_neg(E));

  out.lhs = pairing17(ft, lhs_p, k.g2_s);
  out.rhs = pairing17(ft, rhs_p, k.g2_one);
  out.verdict = (out.lhs.a == out.rhs.a && out.lhs.b == out.rhs.b) ? 1u : 0u;   // gtp_equal, pairing.h:9-11
}

// ---------------------------------------------------------------------------------------------------------------
// Fast path.  Preconditions (checked at context creation / per proof): every key point and srs.g1s[0] is a canonically
// encoded point of E(F_101), and the proof's nine commitments passed step 1.  Every value below is then an element of
// one finite abelian group with a canonical encoding, the reference's g1_add / g1_mul / g1_neg ARE that group's law,
// and the two points handed to the pairing are the same triples whatever the order of the additions.  That allows:
//   - fixed-base pair tables for the nine preprocessed points (two scalars -> one look-up + one addition),
//   - one joint double-and-add (Straus, 2 points per 4-entry sub-table) for the eight proof points of step 11's
//     right-hand side instead of eight separate g1_mul calls.
struct alignas(16) VerifyTables {   // 4704 bytes: a multiple of 16, so that one bulk copy stages it
  uint32_t P2[4][289];     // [0] a*qM + b*qL   [1] a*qR + b*qO   [2] a*qC - b*S3   [3] a*S1 + b*S2   (index a*17 + b)
  uint32_t one_neg[17];    // -(c * g1s[0])
};

PB_HD G1 pick4(uint32_t sel, G1 p, G1 q, G1 pq) {   // sel: 0 -> identity, 1 -> p, 2 -> q, 3 -> p + q
  G1 r = g1_identity();
  if (sel == 1u) r = p;
  if (sel == 2u) r = q;
  if (sel == 3u) r = pq;
  return r;
}

// The joint double-and-add picks one of {identity, A, B, A + B} per pair and bit.  PICK = a per-thread look-up table of packed
// points (x | y << 8 | infinite << 16) in shared memory: one load + unpack per pick instead of three compares and nine selects.
// Layout [pair][entry][lane]: a warp's 32 loads of one instruction hit 32 consecutive words -- no bank conflicts.
#ifndef PB_VERIFY_SMEM_PICK
#define PB_VERIFY_SMEM_PICK 1
#endif
struct PickNone {};
template <int NT>
struct PickSmem {
  uint32_t* base;                                  // this thread's column: word (pair * 4 + entry) * NT
  PB_HD void set(int pair, int entry, G1 p) const { base[(pair * 4 + entry) * NT] = p.x | p.y << 8 | p.inf << 16; }
  PB_HD G1 get(int pair, uint32_t entry) const { return unpack_g1(base[(pair * 4 + entry) * NT]); }
};

// WANT_GT = false: the caller does not read the two pairing values, so the check shares one final exponentiation
// (pairings_equal17_c, curve.cuh) and out.lhs / out.rhs stay zero.
template <bool WANT_GT = true, typename Pick = PickNone>
PB_HD void verify_one_fast(const VerifyKey& k, const VerifyTables& vt, const FieldTables& ft, const uint32_t (&pb)[27],
                          const uint32_t (&op)[7], const uint32_t (&ch)[5], uint32_t u, VerifyOut& out, Pick pick = Pick()) {
  out.lhs = GT{0u, 0u};
  out.rhs = GT{0u, 0u};
  bool bad_pt = false;
  G1 P[9];
#pragma unroll
  for (int j = 0; j < 9; j++) {
    uint32_t x = pb[3 * j], y = pb[3 * j + 1], f = pb[3 * j + 2];
    bad_pt |= x > 100u || y > 100u || f > 1u || (f == 1u && (x | y) != 0u);
    P[j] = G1{x > 100u ? 0u : x, y > 100u ? 0u : y, f != 0u ? 1u : 0u};
    bad_pt |= !g1_is_on_curve(P[j]);
  }
  bool bad_sc = u > 16u;
#pragma unroll
  for (int j = 0; j < 7; j++) bad_sc |= op[j] > 16u;
#pragma unroll
  for (int j = 0; j < 5; j++) bad_sc |= ch[j] > 16u;
  if (bad_pt) { out.verdict = 2u; return; }
  if (bad_sc) { out.verdict = 3u; return; }

  const uint32_t a_z = op[0], b_z = op[1], c_z = op[2], s1_z = op[3], s2_z = op[4], r_z = op[5], zw_z = op[6];
  const uint32_t alpha = ch[0], beta = ch[1], gamma = ch[2], z = ch[3], v = ch[4];
  constexpr uint32_t K1 = 2u, K2 = 3u, OMEGA = 4u;
  const uint32_t z2 = red17(z * z), z3 = red17(z2 * z), z4 = red17(z2 * z2);
  const uint32_t zh_z = sub17(z4, 1u);
  const uint32_t l1_z = red17(13u * (1u + z + z2 + z3));
  const uint32_t alpha2 = red17(alpha * alpha);
  const uint32_t pa = red17(a_z + beta * s1_z + gamma), pbb = red17(b_z + beta * s2_z + gamma);
  const uint32_t pab = red17(pa * pbb);
  const uint32_t perm = red17(red17(red17(pab * red17(c_z + gamma)) * zw_z) * alpha);
  const uint32_t t_num = red17(r_z + 2u * P17 - perm - red17(l1_z * alpha2));
  const uint32_t t_z = red17(t_num * inv17(ft, zh_z));
  const uint32_t bz = red17(beta * z);
  const uint32_t ga = red17(a_z + bz + gamma), gb = red17(b_z + K1 * bz + gamma), gc = red17(c_z + K2 * bz + gamma);
  const uint32_t d_z = red17(red17(red17(red17(red17(ga * gb) * gc) * alpha) * v) + red17(red17(l1_z * alpha2) * v) + u);
  const uint32_t d_s3 = red17(red17(red17(red17(pab * alpha) * v) * beta) * zw_z);
  const uint32_t v2 = red17(v * v), v3 = red17(v2 * v), v4 = red17(v3 * v), v5 = red17(v4 * v), v6 = red17(v5 * v);
  const uint32_t z6 = red17(z4 * z2), z12 = red17(z6 * z6);
  const uint32_t e = red17(t_z + v * r_z + v2 * a_z + v3 * b_z + v4 * c_z + v5 * s1_z + v6 * s2_z + u * zw_z);

  // the preprocessed part of [D] + [F] - [E]: five look-ups
  G1 acc = unpack_g1(vt.P2[0][red17(red17(a_z * b_z) * v) * 17u + red17(a_z * v)]);
  acc = g1_add_c(ft, acc, unpack_g1(vt.P2[1][red17(b_z * v) * 17u + red17(c_z * v)]));
  acc = g1_add_c(ft, acc, unpack_g1(vt.P2[2][v * 17u + d_s3]));
  acc = g1_add_c(ft, acc, unpack_g1(vt.P2[3][v5 * 17u + v6]));
  acc = g1_add_c(ft, acc, unpack_g1(vt.one_neg[e]));
  acc = g1_add_c(ft, acc, P[4]);                                         // [t_lo]

  // z[W_z] + u z omega [W_zw] + z^6[t_mid] + z^12[t_hi] + v^2[a] + v^3[b] + v^4[c] + d_z[z]: joint double-and-add,
  // most significant bit first, two points per 4-entry sub-table
  const G1 A0 = P[7], B0 = P[8], A1 = P[5], B1 = P[6], A2 = P[0], B2 = P[1], A3 = P[2], B3 = P[3];
  const uint32_t sa0 = z, sb0 = red17(red17(u * z) * OMEGA), sa1 = z6, sb1 = z12, sa2 = v2, sb2 = v3, sa3 = v4, sb3 = d_z;
  const G1 S0 = g1_add_c(ft, A0, B0), S1 = g1_add_c(ft, A1, B1), S2 = g1_add_c(ft, A2, B2), S3 = g1_add_c(ft, A3, B3);
  constexpr bool SMEM = !std::is_same<Pick, PickNone>::value;
  if constexpr (SMEM) {
    const G1 id = g1_identity();
    pick.set(0, 0, id); pick.set(0, 1, A0); pick.set(0, 2, B0); pick.set(0, 3, S0);
    pick.set(1, 0, id); pick.set(1, 1, A1); pick.set(1, 2, B1); pick.set(1, 3, S1);
    pick.set(2, 0, id); pick.set(2, 1, A2); pick.set(2, 2, B2); pick.set(2, 3, S2);
    pick.set(3, 0, id); pick.set(3, 1, A3); pick.set(3, 2, B3); pick.set(3, 3, S3);
  }
  auto sel = [&](int pair, uint32_t sa, uint32_t sb, int bit, const G1& A, const G1& B, const G1& S) -> G1 {
    const uint32_t e = ((sa >> bit) & 1u) | (((sb >> bit) & 1u) << 1);
    if constexpr (SMEM) return pick.get(pair, e);
    else return pick4(e, A, B, S);
  };
  // top bit first, peeled: the accumulators start at the identity, whose doubling is the identity and to which an
  // addition returns the other operand, so the first doubling and the first addition of each chain are a plain pick
  G1 s = sel(0, sa0, sb0, 4, A0, B0, S0);
  s = g1_add_c(ft, s, sel(1, sa1, sb1, 4, A1, B1, S1));
  s = g1_add_c(ft, s, sel(2, sa2, sb2, 4, A2, B2, S2));
  s = g1_add_c(ft, s, sel(3, sa3, sb3, 4, A3, B3, S3));
  G1 l = pick4((u >> 4) & 1u, B0, B0, B0);
  for (int bit = 3; bit >= 0; --bit) {
    s = g1_double_c(ft, s);
    s = g1_add_c(ft, s, sel(0, sa0, sb0, bit, A0, B0, S0));
    s = g1_add_c(ft, s, sel(1, sa1, sb1, bit, A1, B1, S1));
    s = g1_add_c(ft, s, sel(2, sa2, sb2, bit, A2, B2, S2));
    s = g1_add_c(ft, s, sel(3, sa3, sb3, bit, A3, B3, S3));
    l = g1_double_c(ft, l);                                              // u [W_zw] for the left-hand side
    l = g1_add_c(ft, l, pick4((u >> bit) & 1u, B0, B0, B0));
  }
  const G1 rhs_p = g1_add_c(ft, acc, s);
  const G1 lhs_p = g1_add_c(ft, P[7], l);
  if constexpr (WANT_GT) {
    out.lhs = final_exp600(miller17_c(ft, lhs_p, k.g2_s));
    out.rhs = final_exp600(miller17_c(ft, rhs_p, k.g2_one));
    out.verdict = (out.lhs.a == out.rhs.a && out.lhs.b == out.rhs.b) ? 1u : 0u;
  } else {
    out.verdict = pairings_equal17_c(ft, lhs_p, k.g2_s, rhs_p, k.g2_one) ? 1u : 0u;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Table path ("log" verifier).  Same preconditions as the fast path.  E(F_101) has 102 points and 102 is squarefree, so
// the group is cyclic: with a generator Q every point is k Q, the reference's g1_add is addition of the k's mod 102 and
// g1_mul(P, s) is s * k (the scalar is never reduced, g1.h:91-103; here s < 17).  The right-hand point of step 11 is then
// ONE sum of 14 products of small integers, reduced mod 102 -- no group operation at all -- and, since the verifier pairs
// with just two G2 points (both in the key), pairing(k Q, g2_s) and pairing(k Q, g2_one) are two 102-entry tables built
// at context creation with the reference's own pairing (pairing17: miller17 + final exponentiation).  The check
// gtp_equal(pairing(lhs, g2_s), pairing(rhs, g2_one)) is two look-ups and a compare.  This is memoisation made possible
// by the toy parameters, exactly like the prover's 17^6-entry commitment table; PB_VERIFY_TABLES=0 at context creation
// keeps the Straus + Miller-loop kernel above, and bench.py reports that configuration too.
// Curve membership (step 1) is a look-up as well: (x, y) is on the curve iff it IS the point at its index in the list.
constexpr uint32_t VLT_MOD17_RANGE = 512u;
struct alignas(16) VerifyLogTables {
  uint32_t pw[104];        // index -> the point as x | y << 8 | infinite << 16 ([0] = 0x10000); [102..103] = 0xFFFFFFFF (no point)
  uint8_t cbase[104];      // x -> index of the first point with abscissa x (wire.cuh: CurveIndexImage); [101] = 102
  uint8_t dlog[104];       // index -> k with point = k Q; dlog[0] = 0
  uint8_t L2[4][292];      // dlog of VerifyTables.P2[t][i]
  uint8_t Lneg[24];        // dlog of VerifyTables.one_neg[e]
  uint8_t inv17[32];       // hf_inverses (FieldTables.inv17): the only field table this path needs
  uint16_t gt_s[104];      // k -> pairing(k Q, g2_s) as a | b << 8
  uint16_t gt_1[104];      // k -> pairing(k Q, g2_one)
  uint8_t mod17[VLT_MOD17_RANGE];   // x mod 17 for x < 512: every reduction of the scalar part but one is of a product of two residues
};
constexpr uint32_t GROUP_ORDER = 102u;
PB_HD uint32_t vlt_px(const VerifyLogTables& lt, uint32_t i) { return lt.pw[i] & 0xFFu; }
PB_HD uint32_t vlt_py(const VerifyLogTables& lt, uint32_t i) { return (lt.pw[i] >> 8) & 0xFFu; }

// index of a canonical curve point in the list (curve_index of wire.cuh, restated here so that this header stands alone)
PB_HD uint32_t vlt_index(const VerifyLogTables& lt, G1 p) {
  const uint32_t i = lt.cbase[p.x < 101u ? p.x : 0u] + (2u * p.y > 101u ? 1u : 0u);
  return p.inf ? 0u : i;
}
// step A of the construction (sequential): the point list, a generator, the discrete logarithms.  Returns false if no
// point has order 102 (cannot happen on this curve; the caller then keeps the Straus kernel).
PB_HD bool vlt_group(const FieldTables& ft, VerifyLogTables& lt, uint8_t (&alog)[104]) {
  uint32_t n = 1;
  lt.pw[0] = 1u << 16;
  for (uint32_t x = 0; x < 101u; x++) {
    lt.cbase[x] = (uint8_t)n;
    const uint32_t rhs = red101(red101(x * x) * x + 3u);
    for (uint32_t y = 0; y < 101u; y++)
      if (red101(y * y) == rhs && n < 104u) { lt.pw[n] = x | y << 8; n++; }
  }
  if (n != GROUP_ORDER) return false;
  for (uint32_t k = 101; k < 104; k++) lt.cbase[k] = (uint8_t)n;
  for (uint32_t k = n; k < 104; k++) lt.pw[k] = 0xFFFFFFFFu;
  for (uint32_t k = 0; k < 32; k++) lt.inv17[k] = ft.inv17[k];
  for (uint32_t k = 0; k < VLT_MOD17_RANGE; k++) lt.mod17[k] = (uint8_t)(k % 17u);
  for (uint32_t cand = 1; cand < GROUP_ORDER; cand++) {
    const G1 q{vlt_px(lt, cand), vlt_py(lt, cand), 0u};
    G1 p = q;
    uint32_t ord = 1;
    while (!p.inf && ord <= GROUP_ORDER) { p = g1_add(ft, p, q); ord++; }     // the reference's addition, g1.h:42-83
    if (ord != GROUP_ORDER) continue;
    for (uint32_t k = 0; k < 104; k++) { lt.dlog[k] = 0; alog[k] = 0; }
    p = q;
    for (uint32_t k = 1; k < GROUP_ORDER; k++) {
      const uint32_t i = vlt_index(lt, p);
      lt.dlog[i] = (uint8_t)k;
      alog[k] = (uint8_t)i;
      p = g1_add(ft, p, q);
    }
    return true;
  }
  return false;
}
// step B (independent entries t < VLT_ENTRIES): the two pairing tables, then the logarithms of the key tables
constexpr uint32_t VLT_ENTRIES = GROUP_ORDER + 4u * 289u + 17u;
PB_HD void vlt_entry(const FieldTables& ft, const VerifyKey& key, const VerifyTables& vt, const uint8_t (&alog)[104], VerifyLogTables& lt, uint32_t t) {
  if (t < GROUP_ORDER) {
    const uint32_t i = alog[t];
    const G1 p{vlt_px(lt, i), vlt_py(lt, i), i == 0u ? 1u : 0u};
    const GT a = pairing17(ft, p, key.g2_s), b = pairing17(ft, p, key.g2_one);
    lt.gt_s[t] = (uint16_t)(a.a | a.b << 8);
    lt.gt_1[t] = (uint16_t)(b.a | b.b << 8);
    if (t < 2u) { lt.gt_s[GROUP_ORDER + t] = 0; lt.gt_1[GROUP_ORDER + t] = 0; }
  } else if (t < GROUP_ORDER + 4u * 289u) {
    const uint32_t j = (t - GROUP_ORDER) / 289u, i = (t - GROUP_ORDER) % 289u;
    lt.L2[j][i] = lt.dlog[vlt_index(lt, unpack_g1(vt.P2[j][i]))];
  } else {
    const uint32_t e = t - GROUP_ORDER - 4u * 289u;
    lt.Lneg[e] = lt.dlog[vlt_index(lt, unpack_g1(vt.one_neg[e]))];
  }
}

template <bool WANT_GT = true>
PB_HD void verify_one_log(const VerifyLogTables& lt, const uint32_t (&pb)[27], const uint32_t (&op)[7], const uint32_t (&ch)[5],
                         uint32_t u, VerifyOut& out) {
  out.lhs = GT{0u, 0u};
  out.rhs = GT{0u, 0u};
  // step 1: encodings and curve membership; lg[j] = discrete logarithm of commitment j.  With w = x | y << 8 | f << 16 the
  // four tests of verify_one (x, y <= 100; f <= 1; f == 1 => x = y = 0; on the curve) are ONE compare: w must BE the
  // listed point at its own index (0 for f != 0, else first index of abscissa min(x, 100), + 1 for the larger root).
  bool bad_pt = false;
  uint32_t lg[9];
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const uint32_t x = pb[3 * j], y = pb[3 * j + 1], f = pb[3 * j + 2];
    uint32_t i = lt.cbase[x > 100u ? 100u : x] + (2u * y > 101u ? 1u : 0u);
    i = f != 0u ? 0u : i;
    bad_pt |= lt.pw[i] != (x | y << 8 | f << 16);
    lg[j] = lt.dlog[i];
  }
  // openings, challenges, u are field elements: all <= 16  <=>  their maximum is
  uint32_t top = u;
#pragma unroll
  for (int j = 0; j < 7; j++) top = umax(top, op[j]);
#pragma unroll
  for (int j = 0; j < 5; j++) top = umax(top, ch[j]);
  if (bad_pt) { out.verdict = 2u; return; }
  if (top > 16u) { out.verdict = 3u; return; }

  // steps 4-10: the scalars, exactly as in verify_one; m() = mod 17 by look-up (arguments < 512, bounds at the sites)
  auto m = [&](uint32_t x) -> uint32_t { PB_BOUND(x, VLT_MOD17_RANGE, "verifier mod17 index"); return lt.mod17[x]; };
  const uint32_t a_z = op[0], b_z = op[1], c_z = op[2], s1_z = op[3], s2_z = op[4], r_z = op[5], zw_z = op[6];
  const uint32_t alpha = ch[0], beta = ch[1], gamma = ch[2], z = ch[3], v = ch[4];
  constexpr uint32_t K1 = 2u, K2 = 3u, OMEGA = 4u;
  const uint32_t z2 = m(z * z), z3 = m(z2 * z), z4 = m(z2 * z2);                       // products of residues: <= 256
  const uint32_t zh_z = sub17(z4, 1u);
  const uint32_t l1_z = m(13u * m(1u + z + z2 + z3));                                  // <= 49, then <= 208
  const uint32_t alpha2 = m(alpha * alpha);
  const uint32_t pa = m(a_z + beta * s1_z + gamma), pbb = m(b_z + beta * s2_z + gamma); // <= 288
  const uint32_t pab = m(pa * pbb);
  const uint32_t perm = m(m(m(pab * m(c_z + gamma)) * zw_z) * alpha);
  const uint32_t l1a2 = m(l1_z * alpha2);
  const uint32_t t_num = m(r_z + 2u * P17 - perm - l1a2);                               // <= 50
  const uint32_t t_z = m(t_num * lt.inv17[zh_z]);
  const uint32_t bz = m(beta * z);
  const uint32_t ga = m(a_z + bz + gamma), gb = m(b_z + K1 * bz + gamma), gc = m(c_z + K2 * bz + gamma);   // <= 80
  const uint32_t d_z = m(m(m(m(m(ga * gb) * gc) * alpha) * v) + m(l1a2 * v) + u);      // <= 48
  const uint32_t d_s3 = m(m(m(m(pab * alpha) * v) * beta) * zw_z);
  const uint32_t v2 = m(v * v), v3 = m(v2 * v), v4 = m(v3 * v), v5 = m(v4 * v), v6 = m(v5 * v);
  const uint32_t z6 = m(z4 * z2), z12 = m(z6 * z6);
  const uint32_t e = red17(t_z + v * r_z + v2 * a_z + v3 * b_z + v4 * c_z + v5 * s1_z + v6 * s2_z + u * zw_z);   // <= 1808: Barrett
  const uint32_t uzw = m(m(u * z) * OMEGA);

  // step 11 in the exponent: [D] + [F] - [E] + z [W_z] + u z omega [W_zw] on the right, [W_z] + u [W_zw] on the left.
  // k_r <= 5 * 101 + 101 + 8 * 16 * 101 < 2^14; floor(k / 102) = (k * 41121) >> 22 is exact below 110 376
  uint32_t k_r = (uint32_t)lt.L2[0][m(m(a_z * b_z) * v) * 17u + m(a_z * v)] + lt.L2[1][m(b_z * v) * 17u + m(c_z * v)] +
                 lt.L2[2][v * 17u + d_s3] + lt.L2[3][v5 * 17u + v6] + lt.Lneg[e] + lg[4] +
                 z * lg[7] + uzw * lg[8] + z6 * lg[5] + z12 * lg[6] + v2 * lg[0] + v3 * lg[1] + v4 * lg[2] + d_z * lg[3];
  k_r -= GROUP_ORDER * ((k_r * 41121u) >> 22);
  uint32_t k_l = lg[7] + u * lg[8];
  k_l -= GROUP_ORDER * ((k_l * 41121u) >> 22);
  const uint32_t gl = lt.gt_s[k_l], gr = lt.gt_1[k_r];
  if constexpr (WANT_GT) {
    out.lhs = GT{gl & 0xFFu, gl >> 8};
    out.rhs = GT{gr & 0xFFu, gr >> 8};
  }
  out.verdict = gl == gr ? 1u : 0u;                                  // gtp_equal, pairing.h:9-11
}

}  // namespace pb
