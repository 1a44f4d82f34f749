// prover.cuh -- plonk_prove (src/plonk.h:223-656) as ONE straight-line device function: one proof per
// thread, every polynomial a fixed-width array of 32-bit registers, all loops unrolled.
//
// Why this is bit-exact although it does not mimic the reference's malloc'd, trimmed POLYs:
// every poly.h constructor trims trailing zeros (poly.h:20-24), so each reference POLY is the
// canonical form of a polynomial over F17 and `len` is deg+1 (1 for the zero polynomial).  F17[x]
// is an integral domain, so the reference's add/sub/mul/scale/divide are the exact ring operations;
// fixed-width zero-padded arithmetic computes the same coefficients, and wherever `len` feeds
// control flow (poly_slice bounds plonk.h:517-519, the SRS guard srs.h:54-57) it is recomputed here
// as the canonical length.  The in-place, untrimmed poly_add_hf (poly.h:67-70) can only zero the
// constant term, which keeps canonical form.  Quirks reproduced on purpose: 1/0 = 0 in the grand
// product (plonk.h:357), r1 without q_C, r3 multiplied by the polynomial z(x) and added
// (plonk.h:537-571), exit paths in the reference's order (Appendix B).
//
// Arithmetic: products are accumulated raw in 32 bits and reduced once per polynomial with
// red17 (field.cuh).  Bounds are stated where they matter; all are far below 2^28.
#pragma once
#include "curve.cuh"
#include "transcript.cuh"

namespace pb {

constexpr int PROVER_SRS_ROWS = 9;   // the prover never touches more than nine SRS points (|W_z| <= 9)

// Small-range reduction table: mod17[x] = x mod 17 for x < 2048.  About three quarters of the prover's reductions are of
// values below 2048 (single products, sums of a few); a shared-memory look-up replaces their IMAD.HI + IMAD pair, and the
// integer multiplier is the pipe this kernel loads most.
constexpr uint32_t MOD17_RANGE = 2048u;
#ifndef PB_PROVE_MODTAB
#define PB_PROVE_MODTAB 1
#endif

// Per-circuit constants, passed to kernels BY VALUE (kernel parameters live in the constant bank,
// so a uniform read is an instruction operand, not a load).
struct CircuitConst {
  uint32_t qv[5][4];    // selector values at the four gates: q_l q_r q_o q_m q_c      (constraints.h:35-41)
  uint32_t QP[5][4];    // their interpolations over H, zero padded: QL QR QO QM QC    (plonk.h:268-272)
  uint32_t sig[3][4];   // sigma_1..3 = copy constraints mapped to H, k1 H, k2 H         (plonk.h:142-160)
  uint32_t SP[3][4];    // S_sigma1..3(x)                                                 (plonk.h:273-275)
  uint32_t vinv[4][4];  // h_pows_inv, the inverse Vandermonde matrix of H               (plonk.h:105-112)
  uint32_t l1[4];       // L1(x) = interpolate([1,0,0,0])                                 (plonk.h:390-391)
  uint32_t srs_len;     // SRS.len (srs.h:13)
  uint32_t bad_copy;    // a COPY_OF.type outside {A,B,C}: every proof exits at plonk.h:155-157
  FsState fs_seed;      // Fiat-Shamir mode only: the transcript state after absorbing circuit and SRS (transcript.cuh)
};

// Tables that are indexed per lane (so they live in shared memory, not in the constant bank).
struct ProverTables {
  FieldTables ft;
  uint8_t pow17[17][20];   // pow17[z][k] = z^k (hf_pow, 0^0 = 1): the powers of the evaluation point are look-ups, not a 17-step chain
  // fixed-base table: T[i][c] = g1_mul(srs.g1s[i], c) for c in [0,17), packed x | y<<8 | inf<<16,
  // computed with the reference's own double-and-add on the device at context creation.
  // Rows >= srs_len are the identity {0,0,1}.
  uint32_t T[PROVER_SRS_ROWS][17];
  uint8_t mod17[MOD17_RANGE];
};

// Fast-path tables, used when every SRS point is a canonically encoded point of E(F_101) (context creation checks
// it; true for both benchmark SRS modes).  The reference's additions are then the exact group law of a finite
// abelian group with a canonical representation per element, so the ORDER of the additions no longer matters and
// neither does adding the (canonical) identity.  T2[j][c0*17 + c1] = c0*g1s[2j] + c1*g1s[2j+1], built on the device
// with the reference's own g1_add from the single-point rows -- for a 2-term polynomial it is literally the
// reference's srs_eval_at_s.  Halves the number of additions per commitment.
constexpr int PROVER_PAIR_ROWS = (PROVER_SRS_ROWS + 1) / 2;
struct ProverPairTables {
  FieldTables ft;
  uint8_t pow17[17][20];
  uint32_t T2[PROVER_PAIR_ROWS][289];
  uint8_t mod17[MOD17_RANGE];
};

// Widest fast path (same eligibility and the same exactness argument as the pair tables): the commitment of a polynomial
// of up to six coefficients is ONE look-up.  T6[c0 + 17 c1 + ... + 17^5 c5] = sum_{i<6} c_i * g1s[i] has 17^6 entries of
// 16 bits (x | y << 7 | infinite << 14: coordinates are < 101) = 48 MB in global memory, built on the device at context
// creation from the single-point rows with the reference's own g1_add.  It fits in the 126 MB L2, and every sector of it
// is hit ~12 times per 2^21-proof launch, so the gathers are L2 hits.  T3 covers SRS rows 6..8 the same way (4913
// entries) for the two longer polynomials z(x) (7 coefficients) and W_z (up to 9): one look-up each and ONE g1_add.
// 2 additions per proof instead of 21.
constexpr uint32_t WIDE_T3_ENTRIES = 17u * 17u * 17u;
constexpr uint32_t WIDE_T6_ENTRIES = WIDE_T3_ENTRIES * WIDE_T3_ENTRIES;
struct ProverWideTables {
  FieldTables ft;
  uint8_t pow17[17][20];
  const uint16_t* T6;   // [17^6], rows 0..5
  const uint16_t* T3;   // [17^3], rows 6..8
  uint8_t mod17[MOD17_RANGE];
};
PB_HD uint32_t pack_g1_16(const G1& p) { return p.x | p.y << 7 | p.inf << 14; }
PB_HD G1 unpack_g1_16(uint32_t w) { return G1{w & 0x7Fu, (w >> 7) & 0x7Fu, w >> 14}; }
PB_HD uint32_t wide_load(const uint16_t* t, uint32_t idx) {
#ifdef __CUDA_ARCH__
  return __ldg(t + idx);
#else
  return t[idx];
#endif
}

PB_HD G1 unpack_g1(uint32_t w) { return G1{w & 0xFFu, (w >> 8) & 0xFFu, (w >> 16) & 1u}; }
PB_HD uint32_t pack_g1(uint32_t x, uint32_t y, uint32_t inf) { return x | (y << 8) | (inf << 16); }

template <int N>
PB_HD void zero(uint32_t (&a)[N]) {
#pragma unroll
  for (int i = 0; i < N; i++) a[i] = 0u;
}
// out += a * b (raw)
template <int NA, int NB>
PB_HD void mul_acc(uint32_t* out, const uint32_t* a, const uint32_t* b) {
#pragma unroll
  for (int i = 0; i < NA; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) out[i + j] += a[i] * b[j];
}
template <int N>
PB_HD void reduce(uint32_t (&a)[N]) {
#pragma unroll
  for (int i = 0; i < N; i++) a[i] = red17(a[i]);
}
// canonical length: deg + 1, and 1 for the zero polynomial (poly.h:20-24)
template <int N>
PB_HD uint32_t canon_len(const uint32_t (&a)[N]) {
  uint32_t len = 1u;
#pragma unroll
  for (int i = 1; i < N; i++) len = a[i] ? (uint32_t)(i + 1) : len;
  return len;
}
template <int N>
PB_HD uint32_t dot(const uint32_t (&a)[N], const uint32_t* zp) {
  uint32_t s = 0u;
#pragma unroll
  for (int i = 0; i < N; i++) s += a[i] * zp[i];
  return red17(s);
}
// h_pows_inv * values (plonk.h:162-195)
PB_HD void interpolate(const CircuitConst& cc, const uint32_t (&v)[4], uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    uint32_t s = 0u;
#pragma unroll
    for (int c = 0; c < 4; c++) s += cc.vinv[r][c] * v[c];
    out[r] = red17(s);
  }
}
// the same with the small-range reducer (row sums are < 4 * 16 * 16 = 1024)
template <typename Tables>
PB_HD uint32_t red_small(const Tables& tb, uint32_t x) {
#if PB_PROVE_MODTAB
  PB_BOUND(x, MOD17_RANGE, "mod17 index");
  return tb.mod17[x];
#else
  return red17(x);
#endif
}
template <typename Tables>
PB_HD void interpolate_s(const Tables& tb, const CircuitConst& cc, const uint32_t (&v)[4], uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    uint32_t s = 0u;
#pragma unroll
    for (int c = 0; c < 4; c++) s += cc.vinv[r][c] * v[c];
    out[r] = red_small(tb, s);
  }
}
template <typename Tables, int N>
PB_HD uint32_t dot_s(const Tables& tb, const uint32_t (&a)[N], const uint32_t* zp) {
  static_assert(N <= 7, "more than seven terms can exceed the table range");
  uint32_t s = 0u;
#pragma unroll
  for (int i = 0; i < N; i++) s += a[i] * zp[i];
  return red_small(tb, s);
}
// KZG commitment (srs.h:53-68) against the fixed-base table: acc = sum_{i < len} T[i][c_i], same order
// of additions as the reference's loop, and the same NUMBER of additions: the loop must stop at the
// canonical length, because adding the identity is not a no-op on the reference's structs when the
// accumulator is an "identity with coordinates" (g1_add returns *b when a is infinite, g1.h:60 --
// reachable with an SRS that holds such points).
template <int N>
PB_HD G1 commit(const ProverTables& tb, const uint32_t (&c)[N], uint32_t len) {
  static_assert(N <= PROVER_SRS_ROWS, "prover polynomial wider than the table");
  G1 acc = g1_identity();
#pragma unroll
  for (int i = 0; i < N; i++) {
    G1 nxt = g1_add(tb.ft, acc, unpack_g1(tb.T[i][c[i]]));
    if ((uint32_t)i < len) acc = nxt;
  }
  return acc;
}

template <int N>
PB_HD G1 commit(const ProverPairTables& tb, const uint32_t (&c)[N], uint32_t /*len: zero coefficients are no-ops here*/) {
  static_assert(N >= 2 && N <= PROVER_SRS_ROWS, "prover polynomial shape");
  G1 acc = unpack_g1(tb.T2[0][c[0] * 17u + c[1]]);
#pragma unroll
  for (int j = 1; j < (N + 1) / 2; j++) {
    const uint32_t idx = c[2 * j] * 17u + (2 * j + 1 < N ? c[2 * j + 1] : 0u);
    acc = g1_add_c(tb.ft, acc, unpack_g1(tb.T2[j][idx]));   // canonical table entries (the pair tables' eligibility)
  }
  return acc;
}

template <int N>
PB_HD G1 commit(const ProverWideTables& tb, const uint32_t (&c)[N], uint32_t /*len: zero coefficients are no-ops here*/) {
  static_assert(N >= 1 && N <= PROVER_SRS_ROWS, "prover polynomial shape");
  constexpr int LO = N < 6 ? N : 6;
  uint32_t idx = c[LO - 1];
#pragma unroll
  for (int i = LO - 2; i >= 0; i--) idx = idx * 17u + c[i];
  G1 acc = unpack_g1_16(wide_load(tb.T6, idx));
  if constexpr (N > 6) {
    uint32_t hi = c[N - 1];
#pragma unroll
    for (int i = N - 2; i >= 6; i--) hi = hi * 17u + c[i];
    acc = g1_add_c(tb.ft, acc, unpack_g1_16(wide_load(tb.T3, hi)));   // both operands are canonical table entries
  }
  return acc;
}

// The wide tables' commitment split in two: the gathers (issue) and unpack + addition (finish).  prove_one can hand the
// LAST commitment, [W_z], back unfinished (DEFER_WZ): its status byte does not depend on the point, so the kernel issues
// the dense-list atomic and writes the rest of the record while the two gathers are in flight (they were 5.6 % of the
// kernel's stall samples when unpacked on the spot, profiles/r2/NOTES.md).
struct WzPending { uint32_t lo, hi; };
PB_HD WzPending wz_issue(const ProverWideTables& tb, const uint32_t (&c)[9]) {
  uint32_t idx = c[5], hi = c[8];
#pragma unroll
  for (int i = 4; i >= 0; i--) idx = idx * 17u + c[i];
#pragma unroll
  for (int i = 7; i >= 6; i--) hi = hi * 17u + c[i];
  return WzPending{wide_load(tb.T6, idx), wide_load(tb.T3, hi)};
}
PB_HD WzPending z_issue(const ProverWideTables& tb, const uint32_t (&c)[7]) {
  uint32_t idx = c[5];
#pragma unroll
  for (int i = 4; i >= 0; i--) idx = idx * 17u + c[i];
  return WzPending{wide_load(tb.T6, idx), wide_load(tb.T3, c[6])};
}
#ifndef PB_PROVE_DEFER_Z
#define PB_PROVE_DEFER_Z 1     // 1: finish [z] after round 3; 2: at the end of prove_one (explicit-challenge mode, wide tables); 0: on the spot
#endif
PB_HD G1 wz_finish(const ProverWideTables& tb, const WzPending& q) { return g1_add_c(tb.ft, unpack_g1_16(q.lo), unpack_g1_16(q.hi)); }

struct ProofOut {
  WzPending wz;     // DEFER_WZ only: pts[7] = wz_finish(tb, wz) is left to the caller
  G1 pts[9];        // a b c z t_lo t_mid t_hi W_z W_zw   (PROOF field order, plonk.h:24-41)
  uint32_t sc[7];   // a_z b_z c_z s_sigma_1_z s_sigma_2_z r_z z_omega_z
  uint32_t status;  // SURVEY.md Appendix B row of the first exit that fires, 0 = completed
  uint32_t ch[6];   // Fiat-Shamir mode only: alpha beta gamma z v u as drawn from the transcript
};

// FS = false: the reference's interface, challenges given by the caller (plonk.h:227).
// FS = true: the five challenge arguments are ignored and drawn from the transcript (transcript.cuh).
template <bool FS = false, bool DEFER_WZ = false, typename Tables>
PB_HD void prove_one(const CircuitConst& cc, const Tables& tb,
                    const uint32_t (&wa)[4], const uint32_t (&wb)[4], const uint32_t (&wc)[4],
                    const uint32_t (&rnd)[9], uint32_t alpha, uint32_t beta, uint32_t gamma,
                    uint32_t z, uint32_t v, ProofOut& out) {
  const uint32_t srs_len = cc.srs_len;
  constexpr uint32_t K1 = 2u, K2 = 3u;                  // plonk.h:13-14
  constexpr uint32_t H[4] = {1u, 4u, 16u, 13u};         // omega^i, omega = 4 (plonk.h:12,69-71)
  constexpr uint32_t OMEGA_POW[7] = {1u, 4u, 16u, 13u, 1u, 4u, 16u};
  auto rs = [&](uint32_t x) -> uint32_t { return red_small(tb, x); };   // x < 2048 at every use (bounds at the sites)

  // ---- step 1: constraints_satisfy (constraints.h:145-171), inside the assert of plonk.h:231
  bool unsat = false;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t lhs = cc.qv[0][i] * wa[i] + cc.qv[1][i] * wb[i] + cc.qv[2][i] * wc[i] +
                   cc.qv[3][i] * (wa[i] * wb[i]) + cc.qv[4][i];        // < 16^3 + 3*256 + 17
    unsat |= red17(lhs) != 0u;
  }

  // ---- round 1 (plonk.h:265-301): a = f_a + (b2 + b1 x) Z_H, Z_H = x^4 - 1 (poly_z(H), plonk.h:116)
  uint32_t fa[4], fb[4], fc[4];
  interpolate_s(tb, cc, wa, fa);
  interpolate_s(tb, cc, wb, fb);
  interpolate_s(tb, cc, wc, fc);
  uint32_t A[6] = {sub17(fa[0], rnd[1]), sub17(fa[1], rnd[0]), fa[2], fa[3], rnd[1], rnd[0]};
  uint32_t B[6] = {sub17(fb[0], rnd[3]), sub17(fb[1], rnd[2]), fb[2], fb[3], rnd[3], rnd[2]};
  uint32_t C[6] = {sub17(fc[0], rnd[5]), sub17(fc[1], rnd[4]), fc[2], fc[3], rnd[5], rnd[4]};
  const uint32_t len_a = canon_len(A), len_b = canon_len(B), len_c = canon_len(C);
  const uint32_t len_abc = umax(umax(len_a, len_b), len_c);
  out.pts[0] = commit(tb, A, len_a);
  out.pts[1] = commit(tb, B, len_b);
  out.pts[2] = commit(tb, C, len_c);
  Transcript tr{cc.fs_seed};
  if constexpr (FS) tr.round1(out.pts[0], out.pts[1], out.pts[2], beta, gamma);

  // ---- round 2 (plonk.h:320-379): grand-product accumulator.  S_sigma_k evaluated at omega^(i-1)
  // is sigma_k[i-1] (S_sigma_k interpolates sigma_k over H), so no Horner evaluation is needed.
  uint32_t acc[4];
  acc[0] = 1u;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    uint32_t w = H[i];
    uint32_t d0 = rs(wa[i] + beta * w + gamma);             // every value in this loop is < 16 + 16 * 48 + 16
    uint32_t d1 = rs(wb[i] + beta * (K1 * w) + gamma);
    uint32_t d2 = rs(wc[i] + beta * (K2 * w) + gamma);
    uint32_t n0 = rs(wa[i] + beta * cc.sig[0][i] + gamma);
    uint32_t n1 = rs(wb[i] + beta * cc.sig[1][i] + gamma);
    uint32_t n2 = rs(wc[i] + beta * cc.sig[2][i] + gamma);
    uint32_t den = rs(rs(d0 * d1) * d2);
    uint32_t num = rs(rs(n0 * n1) * n2);
    uint32_t frac = rs(den * inv17(tb.ft, num));        // hf_div: 1/0 = 0 (hf.h:188-203)
    acc[i + 1] = rs(acc[i] * frac);
  }
  uint32_t accx[4];
  interpolate_s(tb, cc, acc, accx);
  // plonk.h:366-368: acc_x(omega^n) == 1, omega^4 = 1
  const bool bad_acc = rs(accx[0] + accx[1] + accx[2] + accx[3]) != 1u;
  uint32_t Z[7] = {sub17(accx[0], rnd[8]), sub17(accx[1], rnd[7]), sub17(accx[2], rnd[6]), accx[3],
                   rnd[8], rnd[7], rnd[6]};
  const uint32_t len_z = canon_len(Z);
  constexpr bool DEFER_Z = PB_PROVE_DEFER_Z != 0 && DEFER_WZ;
  WzPending zq{0u, 0u};
  if constexpr (DEFER_Z) zq = z_issue(tb, Z);
  else out.pts[3] = commit(tb, Z, len_z);
  if constexpr (FS) tr.round2(out.pts[3], alpha);

  // ---- round 3 (plonk.h:385-511): t_numer = t1 + t2 - t3 + t4, all raw, one reduction at the end
  uint32_t tn[22];
  zero(tn);
  {  // t1 = a b q_M + a q_L + b q_R + c q_O + PI + q_C      (PI = 0, plonk.h:398)
    uint32_t ab[11];
    zero(ab);
    mul_acc<6, 6>(ab, A, B);            // raw < 6 * 2^8; times q_M below < 2^17: no reduction needed in between
    mul_acc<11, 4>(tn, ab, cc.QP[3]);
    mul_acc<6, 4>(tn, A, cc.QP[0]);
    mul_acc<6, 4>(tn, B, cc.QP[1]);
    mul_acc<6, 4>(tn, C, cc.QP[2]);
#pragma unroll
    for (int i = 0; i < 4; i++) tn[i] += cc.QP[4][i];
  }
  {  // t2 = alpha (a + beta x + gamma)(b + beta k1 x + gamma)(c + beta k2 x + gamma) z(x)
    uint32_t xa[6], xb[6], xc[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { xa[i] = A[i]; xb[i] = B[i]; xc[i] = C[i]; }
    xa[0] += gamma; xa[1] += beta;
    xb[0] += gamma; xb[1] += rs(beta * K1);
    xc[0] += gamma; xc[1] += rs(beta * K2);
#pragma unroll
    for (int i = 0; i < 6; i++) xa[i] = rs(xa[i] * alpha);      // < 33 * 16
    xb[0] = rs(xb[0]); xb[1] = rs(xb[1]);
    xc[0] = rs(xc[0]); xc[1] = rs(xc[1]);
    uint32_t p1[11], p2[16];
    zero(p1); zero(p2);
    mul_acc<6, 6>(p1, xa, xb);      // < 6 * 2^8
    mul_acc<11, 6>(p2, p1, xc);     // < 6 * 2^4 * 6 * 2^8 < 2^18
    mul_acc<16, 7>(tn, p2, Z);      // < 7 * 2^4 * 2^18 < 2^25
  }
  uint32_t Zw[7];                   // z(omega x), plonk.h:466-470
#pragma unroll
  for (int i = 0; i < 7; i++) Zw[i] = rs(Z[i] * OMEGA_POW[i]);
  // The second opening polynomial W_zw = (z(x) - z(omega z)) / (x - z omega) (plonk.h:612-617, 621) needs only z(x) and
  // the evaluation point.  With caller-supplied challenges it is computed and committed HERE, so that its addition
  // chain overlaps the polynomial products of round 3; in Fiat-Shamir mode z exists only after round 3.
  uint32_t zw_z = 0u, len_wzw = 0u;
  bool bad_rem2 = false;
  auto open_zw = [&]() {
    uint32_t zq[7];
#pragma unroll
    for (int i = 0; i < 7; i++) zq[i] = tb.pow17[z][i];
    zw_z = dot_s(tb, Zw, zq);                                // 7 terms < 7 * 256
    const uint32_t zo = rs(z * 4u);
    uint32_t Wzw[6];
    Wzw[5] = Z[6];
#pragma unroll
    for (int j = 5; j >= 1; j--) Wzw[j - 1] = rs(Z[j] + zo * Wzw[j]);
    bad_rem2 = rs(Z[0] + 17u - zw_z + zo * Wzw[0]) != 0u;
    len_wzw = canon_len(Wzw);
    out.pts[8] = commit(tb, Wzw, len_wzw);
  };
  if constexpr (!FS) open_zw();
  {  // t3 = alpha (a + beta S1 + gamma)(b + beta S2 + gamma)(c + beta S3 + gamma) z(omega x)
    uint32_t ya[6], yb[6], yc[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { ya[i] = A[i]; yb[i] = B[i]; yc[i] = C[i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      ya[i] += beta * cc.SP[0][i];
      yb[i] += beta * cc.SP[1][i];
      yc[i] += beta * cc.SP[2][i];
    }
    ya[0] += gamma; yb[0] += gamma; yc[0] += gamma;
#pragma unroll
    for (int i = 0; i < 6; i++) ya[i] = rs(rs(ya[i]) * alpha);    // ya < 16 + 256 + 16
#pragma unroll
    for (int i = 0; i < 4; i++) { yb[i] = rs(yb[i]); yc[i] = rs(yc[i]); }
    uint32_t p1[11], p2[16], t3[22];
    zero(p1); zero(p2); zero(t3);
    mul_acc<6, 6>(p1, ya, yb);
    mul_acc<11, 6>(p2, p1, yc);
    mul_acc<16, 7>(t3, p2, Zw);     // < 2^25
    constexpr uint32_t BIG = 17u << 21;   // a multiple of 17 above any t3 coefficient
#pragma unroll
    for (int i = 0; i < 22; i++) tn[i] += BIG - t3[i];
  }
  {  // t4 = alpha^2 (z(x) - 1) L1(x)
    uint32_t a2 = rs(alpha * alpha);
    uint32_t zm[7];
#pragma unroll
    for (int i = 0; i < 7; i++) zm[i] = Z[i];
    zm[0] += 16u;
#pragma unroll
    for (int i = 0; i < 7; i++) zm[i] = rs(zm[i] * a2);         // < 32 * 16
    mul_acc<7, 4>(tn, zm, cc.l1);
  }
  reduce(tn);

  // t = t_numer / Z_H (poly_divide, poly.h:124-177, with the divisor x^4 - 1): t[j] = tn[j+4] + t[j+4]
  uint32_t T[18];
#pragma unroll
  for (int j = 17; j >= 0; j--) T[j] = tn[j + 4] + (j + 4 < 18 ? T[j + 4] : 0u);   // < 5 * 17
#pragma unroll
  for (int j = 0; j < 18; j++) T[j] = rs(T[j]);
  bool bad_rem = false;                                     // plonk.h:507-510
#pragma unroll
  for (int k = 0; k < 4; k++) bad_rem |= add17(tn[k], T[k]) != 0u;
  const uint32_t len_t = canon_len(T);
  const bool bad_slice = len_t < 13u;                       // poly_slice(t, 12, len) needs 12 < len (plonk.h:517-519)
  uint32_t tlo[6], tmid[6], thi[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { tlo[i] = T[i]; tmid[i] = T[6 + i]; thi[i] = T[12 + i]; }
  const uint32_t len_lo = canon_len(tlo), len_mid = canon_len(tmid), len_hi = canon_len(thi);
  const uint32_t len_tparts = umax(umax(len_lo, len_mid), len_hi);
  out.pts[4] = commit(tb, tlo, len_lo);
  out.pts[5] = commit(tb, tmid, len_mid);
  out.pts[6] = commit(tb, thi, len_hi);
  if constexpr (FS) {
    tr.round3(out.pts[4], out.pts[5], out.pts[6], z);
    open_zw();
  }

  if constexpr (DEFER_Z && PB_PROVE_DEFER_Z == 1) out.pts[3] = wz_finish(tb, zq);
  // ---- round 4 (plonk.h:527-574): openings at z and the (non-standard) linearisation r(x)
  uint32_t zp[18];
#pragma unroll
  for (int i = 0; i < 18; i++) zp[i] = tb.pow17[z][i];
  const uint32_t a_z = dot_s(tb, A, zp), b_z = dot_s(tb, B, zp), c_z = dot_s(tb, C, zp);   // 6 terms < 6 * 256
  const uint32_t s1_z = dot_s(tb, cc.SP[0], zp), s2_z = dot_s(tb, cc.SP[1], zp);
  const uint32_t t_z = dot(T, zp), l1_z = dot_s(tb, cc.l1, zp);
  const uint32_t a2 = rs(alpha * alpha);
  const uint32_t bz = rs(beta * z);
  const uint32_t e_a = rs(a_z + bz + gamma);
  const uint32_t e_b = rs(b_z + K1 * bz + gamma);
  const uint32_t e_c = rs(c_z + K2 * bz + gamma);
  const uint32_t k2 = rs(rs(rs(e_a * e_b) * e_c) * alpha);                      // r2 scalar
  const uint32_t f_a = rs(a_z + beta * s1_z + gamma), f_b = rs(b_z + beta * s2_z + gamma);
  const uint32_t k3 = rs(rs(rs(f_a * f_b) * alpha) * rs(beta * zw_z));       // r3 scalar
  const uint32_t k4 = rs(l1_z * a2);                                                   // r4 scalar
  const uint32_t abz = rs(a_z * b_z);
  uint32_t R[10];
  zero(R);
  {
    uint32_t zs3[10];
    zero(zs3);
    mul_acc<7, 4>(zs3, Z, cc.SP[2]);                        // z(x) S_sigma3(x), < 4 * 2^8
#pragma unroll
    for (int i = 0; i < 10; i++) R[i] = zs3[i] * k3;
#pragma unroll
    for (int i = 0; i < 7; i++) R[i] += Z[i] * (k2 + k4);
#pragma unroll
    for (int i = 0; i < 4; i++)
      R[i] += abz * cc.QP[3][i] + a_z * cc.QP[0][i] + b_z * cc.QP[1][i] + c_z * cc.QP[2][i];
  }
  reduce(R);
  const uint32_t r_z = dot(R, zp);
  if constexpr (FS) {
    const uint32_t sc[7] = {a_z, b_z, c_z, s1_z, s2_z, r_z, zw_z};
    tr.round4(sc, v);
  }

  // ---- round 5 (plonk.h:582-621): opening polynomials
  const uint32_t v2 = rs(v * v), v3 = rs(v2 * v), v4 = rs(v3 * v), v5 = rs(v4 * v), v6 = rs(v5 * v);
  uint32_t wn[10];
#pragma unroll
  for (int i = 0; i < 10; i++) wn[i] = v * R[i];
#pragma unroll
  for (int i = 0; i < 6; i++)
    wn[i] += tlo[i] + zp[6] * tmid[i] + zp[12] * thi[i] + v2 * A[i] + v3 * B[i] + v4 * C[i];
#pragma unroll
  for (int i = 0; i < 4; i++) wn[i] += v5 * cc.SP[0][i] + v6 * cc.SP[1][i];
  // constant term: - t(z) - v r(z) - v^2 a(z) - ... (the in-place poly_add_hf calls of plonk.h:586-598)
  wn[0] += 17u * 1536u - (t_z + v * r_z + v2 * a_z + v3 * b_z + v4 * c_z + v5 * s1_z + v6 * s2_z);
  reduce(wn);
  // divide by (x - z): q[8] = wn[9], q[j-1] = wn[j] + z q[j]; remainder wn[0] + z q[0]  (plonk.h:606-610)
  uint32_t Wz[9];
  Wz[8] = wn[9];
#pragma unroll
  for (int j = 8; j >= 1; j--) Wz[j - 1] = rs(wn[j] + z * Wz[j]);
  const bool bad_rem1 = rs(wn[0] + z * Wz[0]) != 0u;
  const uint32_t len_wz = canon_len(Wz);
  const uint32_t len_w = umax(len_wz, len_wzw);
  if constexpr (DEFER_WZ) out.wz = wz_issue(tb, Wz);
  else out.pts[7] = commit(tb, Wz, len_wz);
  if constexpr (FS) {
    out.ch[0] = alpha; out.ch[1] = beta; out.ch[2] = gamma; out.ch[3] = z; out.ch[4] = v;
    tr.round5(out.pts[7], out.pts[8], out.ch[5]);
  }

  if constexpr (DEFER_Z && PB_PROVE_DEFER_Z == 2) out.pts[3] = wz_finish(tb, zq);
  out.sc[0] = a_z; out.sc[1] = b_z; out.sc[2] = c_z; out.sc[3] = s1_z; out.sc[4] = s2_z;
  out.sc[5] = r_z; out.sc[6] = zw_z;

  // first exit that fires, in the reference's execution order
  uint32_t st = 0u;
  st = len_w > srs_len ? 12u : st;
  st = (bad_rem1 || bad_rem2) ? 11u : st;
  st = len_tparts > srs_len ? 10u : st;
  st = bad_slice ? 9u : st;
  st = bad_rem ? 8u : st;
  st = len_z > srs_len ? 7u : st;
  st = bad_acc ? 6u : st;
  st = len_abc > srs_len ? 5u : st;
  st = cc.bad_copy ? 3u : st;
  st = unsat ? 1u : st;
  out.status = st;
}

}  // namespace pb
