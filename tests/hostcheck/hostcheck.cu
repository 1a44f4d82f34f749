// tests/hostcheck/hostcheck.cu -- TEST INFRASTRUCTURE.  Compiles the product's __host__ __device__ source
// (field.cuh, curve.cuh, prover.cuh, verifier.cuh) for the HOST, so that the exact code the GPU runs
// can be diffed against the oracle in this GPU-less container.  Never linked into the product library and
// never used as a fallback: libplonk_b200.so has no host execution path.
#include <vector>
#include <type_traits>
#include <cstdint>
#include <cstring>
#include "../../plonk.c_b200/csrc/prover.cuh"
#include "../../plonk.c_b200/csrc/verifier.cuh"
#include "../../plonk.c_b200/csrc/wire.cuh"

using namespace pb;

static FieldTables make_ft() {
  FieldTables ft;
  memset(&ft, 0, sizeof ft);
  for (uint32_t i = 0; i < 256; i++) ft.inv101[i] = (uint8_t)pow101(i % 101u, 99);
  for (uint32_t i = 0; i < 17; i++) ft.inv17[i] = (uint8_t)pow17(i, 15);
  return ft;
}
static G1 ld(const uint8_t* p) { return G1{p[0], p[1], p[2] ? 1u : 0u}; }
static void st(uint8_t* p, G1 g) { p[0] = (uint8_t)g.x; p[1] = (uint8_t)g.y; p[2] = (uint8_t)g.inf; }

// one batch through prove_one; chal == NULL selects the Fiat-Shamir instantiation, whose chal_out (optional) follows the
// rule of prove_kernel: a challenge exists only if the reference's execution reaches the point where it is drawn
template <typename Tables>
static void prove_batch(const CircuitConst& cc, const Tables& tb, const uint8_t* wit, const uint8_t* rnd, const uint8_t* chal,
                        uint8_t* proofs, uint8_t* status, uint8_t* chal_out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    uint32_t wa[4], wb[4], wc[4], r[9];
    for (int k = 0; k < 4; k++) { wa[k] = wit[12 * i + k]; wb[k] = wit[12 * i + 4 + k]; wc[k] = wit[12 * i + 8 + k]; }
    for (int k = 0; k < 9; k++) r[k] = rnd[9 * i + k];
    ProofOut o;
    if (chal) {
      const uint8_t* ch = chal + 5 * i;
      if constexpr (std::is_same<Tables, ProverWideTables>::value) {   // as prove_kernel runs it: [W_z] handed back unfinished
        prove_one<false, true>(cc, tb, wa, wb, wc, r, ch[0], ch[1], ch[2], ch[3], ch[4], o);
        o.pts[7] = wz_finish(tb, o.wz);
      } else {
        prove_one<false>(cc, tb, wa, wb, wc, r, ch[0], ch[1], ch[2], ch[3], ch[4], o);
      }
    } else {
      prove_one<true>(cc, tb, wa, wb, wc, r, 0u, 0u, 0u, 0u, 0u, o);
      if (chal_out) {
        const uint32_t s = o.status;
        const bool k1 = s == 0u || (s >= 6u && s <= 12u), k2 = s == 0u || (s >= 8u && s <= 12u);
        const bool k3 = s == 0u || (s >= 11u && s <= 12u), k5 = s == 0u;
        const bool known[6] = {k2, k1, k1, k3, k3, k5};
        for (int k = 0; k < 6; k++) chal_out[6 * i + k] = known[k] ? (uint8_t)o.ch[k] : 0xFF;
      }
    }
    uint8_t* po = proofs + 34 * i;
    memset(po, 0, 34);
    if (o.status == 0) {
      for (int j = 0; j < 9; j++) st(po + 3 * j, o.pts[j]);
      for (int j = 0; j < 7; j++) po[27 + j] = (uint8_t)o.sc[j];
    }
    status[i] = (uint8_t)o.status;
  }
}
extern "C" {

void hc_g1_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  FieldTables ft = make_ft();
  for (size_t i = 0; i < n; i++) {
    G1 r = op == 0 ? g1_add(ft, ld(a + 3 * i), ld(b + 3 * i)) : op == 1 ? g1_double(ft, ld(a + 3 * i)) : g1_neg(ld(a + 3 * i));
    st(out + 3 * i, r);
  }
}
void hc_g1_mul(const uint8_t* p, const uint64_t* s, uint8_t* out, size_t n) {
  FieldTables ft = make_ft();
  for (size_t i = 0; i < n; i++) st(out + 3 * i, g1_mul(ft, ld(p + 3 * i), s[i]));
}
void hc_g1_on_curve(const uint8_t* p, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = g1_is_on_curve(ld(p + 3 * i)) ? 1 : 0;
}
void hc_g2_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  FieldTables ft = make_ft();
  for (size_t i = 0; i < n; i++) {
    G2 r = op == 0 ? g2_add(ft, G2{a[2 * i], a[2 * i + 1]}, G2{b[2 * i], b[2 * i + 1]}) : g2_neg(G2{a[2 * i], a[2 * i + 1]});
    out[2 * i] = (uint8_t)r.x; out[2 * i + 1] = (uint8_t)r.y;
  }
}
void hc_g2_mul(const uint8_t* p, const uint64_t* s, uint8_t* out, size_t n) {
  FieldTables ft = make_ft();
  for (size_t i = 0; i < n; i++) { G2 r = g2_mul(ft, G2{p[2 * i], p[2 * i + 1]}, s[i]); out[2 * i] = (uint8_t)r.x; out[2 * i + 1] = (uint8_t)r.y; }
}
void hc_gtp_mul(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) { GT r = gt_mul(GT{a[2 * i], a[2 * i + 1]}, GT{b[2 * i], b[2 * i + 1]}); out[2 * i] = (uint8_t)r.a; out[2 * i + 1] = (uint8_t)r.b; }
}
void hc_gtp_pow(const uint8_t* a, const uint64_t* e, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) { GT r = gt_pow(GT{a[2 * i], a[2 * i + 1]}, e[i]); out[2 * i] = (uint8_t)r.a; out[2 * i + 1] = (uint8_t)r.b; }
}
void hc_line(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) { Line l = line_through(ld(a + 3 * i), ld(b + 3 * i)); out[3 * i] = (uint8_t)l.x; out[3 * i + 1] = (uint8_t)l.y; out[3 * i + 2] = (uint8_t)l.c; }
}
void hc_pairing(const uint8_t* p, const uint8_t* q, uint8_t* out, size_t n) {
  FieldTables ft = make_ft();
  for (size_t i = 0; i < n; i++) { GT r = pairing17(ft, ld(p + 3 * i), G2{q[2 * i], q[2 * i + 1]}); out[2 * i] = (uint8_t)r.a; out[2 * i + 1] = (uint8_t)r.b; }
}
void hc_pairing_f(uint64_t r, const uint8_t* p, const uint8_t* q, uint8_t* out, size_t n) {
  FieldTables ft = make_ft();
  for (size_t i = 0; i < n; i++) { GT f = miller(ft, r, ld(p + 3 * i), G2{q[2 * i], q[2 * i + 1]}); out[2 * i] = (uint8_t)f.a; out[2 * i + 1] = (uint8_t)f.b; }
}
// cc: CircuitConst as uint32 words (struct order); table: [9][17] packed.  chal NULL = Fiat-Shamir mode.
void hc_prove(const uint32_t* cc_words, const uint32_t* table, const uint8_t* wit, const uint8_t* rnd, const uint8_t* chal,
              uint8_t* proofs, uint8_t* status, uint8_t* chal_out, size_t n) {
  CircuitConst cc;
  memcpy(&cc, cc_words, sizeof cc);
  ProverTables tb;
  tb.ft = make_ft();
  for (uint32_t zz = 0; zz < 17; zz++) for (uint32_t k = 0; k < 20; k++) tb.pow17[zz][k] = (uint8_t)pow17(zz, k);
  for (uint32_t i = 0; i < MOD17_RANGE; i++) tb.mod17[i] = (uint8_t)(i % 17u);
  memcpy(tb.T, table, sizeof tb.T);
  prove_batch(cc, tb, wit, rnd, chal, proofs, status, chal_out, n);
}
// fast path: pair tables built from the single-point rows exactly as pair_table_kernel does
void hc_prove_pairs(const uint32_t* cc_words, const uint32_t* table, const uint8_t* wit, const uint8_t* rnd, const uint8_t* chal,
                    uint8_t* proofs, uint8_t* status, uint8_t* chal_out, size_t n) {
  CircuitConst cc;
  memcpy(&cc, cc_words, sizeof cc);
  static ProverPairTables tb;
  tb.ft = make_ft();
  for (uint32_t zz = 0; zz < 17; zz++) for (uint32_t k = 0; k < 20; k++) tb.pow17[zz][k] = (uint8_t)pow17(zz, k);
  for (uint32_t i = 0; i < MOD17_RANGE; i++) tb.mod17[i] = (uint8_t)(i % 17u);
  for (uint32_t k = 0; k < PROVER_PAIR_ROWS * 289u; k++) {
    const uint32_t j = k / 289u, c0 = (k % 289u) / 17u, c1 = k % 17u;
    const G1 p = unpack_g1(table[(2 * j) * 17 + c0]);
    const G1 q = 2 * j + 1 < (uint32_t)PROVER_SRS_ROWS ? unpack_g1(table[(2 * j + 1) * 17 + c1]) : g1_identity();
    const G1 r = g1_add(tb.ft, p, q);
    tb.T2[j][k % 289u] = pack_g1(r.x, r.y, r.inf);
  }
  prove_batch(cc, tb, wit, rnd, chal, proofs, status, chal_out, n);
}
// widest fast path: T3 / T6 built exactly as wide_t3_kernel / wide_t6_kernel do (cached per single-point table)
static std::vector<uint16_t> g_wide_store;
static void wide_build(const uint32_t* table) {
  static uint32_t cached[PROVER_SRS_ROWS * 17];
  static bool have = false;
  const FieldTables ft = make_ft();
  if (have && memcmp(cached, table, sizeof cached) == 0) return;
  g_wide_store.assign(3u * (size_t)WIDE_T3_ENTRIES + WIDE_T6_ENTRIES, 0);
  for (uint32_t k = 0; k < 3u * WIDE_T3_ENTRIES; k++) {
    const uint32_t t = k / WIDE_T3_ENTRIES, e = k % WIDE_T3_ENTRIES;
    const uint32_t c[3] = {e % 17u, (e / 17u) % 17u, e / 289u};
    G1 acc = g1_identity();
    for (uint32_t i = 0; i < 3; i++) acc = g1_add(ft, acc, unpack_g1(table[(3u * t + i) * 17u + c[i]]));
    g_wide_store[k] = (uint16_t)pack_g1_16(acc);
  }
  uint16_t* t6 = g_wide_store.data() + 3u * (size_t)WIDE_T3_ENTRIES;
  for (uint32_t k = 0; k < WIDE_T6_ENTRIES; k++) {
    const G1 r = g1_add(ft, unpack_g1_16(g_wide_store[k % WIDE_T3_ENTRIES]), unpack_g1_16(g_wide_store[WIDE_T3_ENTRIES + k / WIDE_T3_ENTRIES]));
    t6[k] = (uint16_t)pack_g1_16(r);
  }
  memcpy(cached, table, sizeof cached);
  have = true;
}
// every entry of T6 (17^6) and of the rows-6..8 T3 (17^3) unpacked to (x, y, infinite) triples, for comparison with srs_eval_at_s
void hc_wide_tables(const uint32_t* table, uint8_t* t6_out /*[17^6][3]*/, uint8_t* t3_out /*[17^3][3]*/) {
  wide_build(table);
  const uint16_t* t6 = g_wide_store.data() + 3u * (size_t)WIDE_T3_ENTRIES;
  const uint16_t* t3 = g_wide_store.data() + 2u * (size_t)WIDE_T3_ENTRIES;
  for (uint32_t k = 0; k < WIDE_T6_ENTRIES; k++) st(t6_out + 3 * (size_t)k, unpack_g1_16(t6[k]));
  for (uint32_t k = 0; k < WIDE_T3_ENTRIES; k++) st(t3_out + 3 * (size_t)k, unpack_g1_16(t3[k]));
}
void hc_prove_wide(const uint32_t* cc_words, const uint32_t* table, const uint8_t* wit, const uint8_t* rnd, const uint8_t* chal,
                   uint8_t* proofs, uint8_t* status, uint8_t* chal_out, size_t n) {
  CircuitConst cc;
  memcpy(&cc, cc_words, sizeof cc);
  wide_build(table);
  std::vector<uint16_t>& store = g_wide_store;
  const FieldTables ft = make_ft();
  ProverWideTables tb;
  tb.ft = ft;
  for (uint32_t zz = 0; zz < 17; zz++) for (uint32_t k = 0; k < 20; k++) tb.pow17[zz][k] = (uint8_t)pow17(zz, k);
  for (uint32_t i = 0; i < MOD17_RANGE; i++) tb.mod17[i] = (uint8_t)(i % 17u);
  tb.T6 = store.data() + 3u * (size_t)WIDE_T3_ENTRIES;
  tb.T3 = store.data() + 2u * (size_t)WIDE_T3_ENTRIES;
  prove_batch(cc, tb, wit, rnd, chal, proofs, status, chal_out, n);
}
void hc_fs_seed(const uint8_t* circuit, const uint8_t* g1s, uint32_t srs_len, const uint8_t* g2, uint32_t out[4]) {
  const FsState st = fs_seed_host(circuit, g1s, srs_len, g2);
  memcpy(out, st.v, 16);
}
void hc_fs_challenges(const uint32_t seed4[4], const uint8_t* proofs, uint8_t* chal6, size_t n) {
  FsState seed;
  memcpy(seed.v, seed4, 16);
  for (size_t i = 0; i < n; i++) {
    uint32_t pbv[27], op[7], ch[5], u;
    for (int j = 0; j < 27; j++) pbv[j] = proofs[34 * i + j];
    for (int j = 0; j < 7; j++) op[j] = proofs[34 * i + 27 + j];
    fs_derive(seed, pbv, op, ch, u);
    for (int j = 0; j < 5; j++) chal6[6 * i + j] = (uint8_t)ch[j];
    chal6[6 * i + 5] = (uint8_t)u;
  }
}
// fast-path verifier: tables built exactly as verify_tables_kernel does, from g1_mul rows of the nine key points
// chal NULL = Fiat-Shamir mode (challenges and u from the proof bytes, transcript seeded with fs_seed)
void hc_verify_fast(const uint8_t* key, const uint32_t seed4[4], const uint8_t* proofs, const uint8_t* chal, const uint8_t* u, uint8_t* verdict, uint8_t* gt, size_t n) {
  FieldTables ft = make_ft();
  FsState fs_seed;
  memcpy(fs_seed.v, seed4, 16);
  VerifyKey k;
  G1* dst[9] = {&k.qm, &k.ql, &k.qr, &k.qo, &k.qc, &k.s1, &k.s2, &k.s3, &k.g1_one};
  for (int j = 0; j < 9; j++) *dst[j] = ld(key + 3 * j);
  k.g2_one = G2{key[27], key[28]};
  k.g2_s = G2{key[29], key[30]};
  static uint32_t KT[9][17];
  for (int j = 0; j < 9; j++)
    for (uint32_t c = 0; c < 17; c++) { G1 r = g1_mul(ft, *dst[j], c); KT[j][c] = pack_g1(r.x, r.y, r.inf); }
  static VerifyTables vt;
  const int ia[4] = {0, 2, 4, 5}, ib[4] = {1, 3, 7, 6};
  for (uint32_t t = 0; t < 4; t++)
    for (uint32_t a = 0; a < 17; a++)
      for (uint32_t b = 0; b < 17; b++) {
        G1 p = unpack_g1(KT[ia[t]][a]), q = unpack_g1(KT[ib[t]][b]);
        if (t == 2) q = g1_neg(q);
        G1 r = g1_add(ft, p, q);
        vt.P2[t][a * 17 + b] = pack_g1(r.x, r.y, r.inf);
      }
  for (uint32_t c = 0; c < 17; c++) { G1 r = g1_neg(unpack_g1(KT[8][c])); vt.one_neg[c] = pack_g1(r.x, r.y, r.inf); }
  for (size_t i = 0; i < n; i++) {
    uint32_t pbv[27], op[7], ch[5];
    for (int j = 0; j < 27; j++) pbv[j] = proofs[34 * i + j];
    for (int j = 0; j < 7; j++) op[j] = proofs[34 * i + 27 + j];
    uint32_t uu;
    if (chal) { for (int j = 0; j < 5; j++) ch[j] = chal[5 * i + j]; uu = u[i]; }
    else fs_derive(fs_seed, pbv, op, ch, uu);
    VerifyOut o;
    uint32_t pick_buf[16];                                          // the kernel's per-thread sub-table column (stride 1 here)
    if (gt) verify_one_fast<true>(k, vt, ft, pbv, op, ch, uu, o, PickSmem<1>{pick_buf});
    else if (i & 1) verify_one_fast<false>(k, vt, ft, pbv, op, ch, uu, o, PickSmem<1>{pick_buf});   // verdict only: one shared final exponentiation
    else verify_one_fast<false>(k, vt, ft, pbv, op, ch, uu, o);     // ... and the select-based picks on every other item
    verdict[i] = (uint8_t)o.verdict;
    if (gt) { gt[4 * i] = (uint8_t)o.lhs.a; gt[4 * i + 1] = (uint8_t)o.lhs.b; gt[4 * i + 2] = (uint8_t)o.rhs.a; gt[4 * i + 3] = (uint8_t)o.rhs.b; }
  }
}
// table-path verifier (verifier.cuh: verify_one_log): the tables are built by the same two functions the context-creation
// kernel calls (vlt_group, vlt_entry), from a VerifyTables built as above.  Returns 0, or 1 if no generator was found.
static void build_vt(const FieldTables& ft, G1* const* dst, VerifyTables& vt) {
  static uint32_t KT[9][17];
  for (int j = 0; j < 9; j++)
    for (uint32_t c = 0; c < 17; c++) { G1 r = g1_mul(ft, *dst[j], c); KT[j][c] = pack_g1(r.x, r.y, r.inf); }
  const int ia[4] = {0, 2, 4, 5}, ib[4] = {1, 3, 7, 6};
  for (uint32_t t = 0; t < 4; t++)
    for (uint32_t a = 0; a < 17; a++)
      for (uint32_t b = 0; b < 17; b++) {
        G1 p = unpack_g1(KT[ia[t]][a]), q = unpack_g1(KT[ib[t]][b]);
        if (t == 2) q = g1_neg(q);
        G1 r = g1_add(ft, p, q);
        vt.P2[t][a * 17 + b] = pack_g1(r.x, r.y, r.inf);
      }
  for (uint32_t c = 0; c < 17; c++) { G1 r = g1_neg(unpack_g1(KT[8][c])); vt.one_neg[c] = pack_g1(r.x, r.y, r.inf); }
}
int hc_verify_log(const uint8_t* key, const uint32_t seed4[4], const uint8_t* proofs, const uint8_t* chal, const uint8_t* u, uint8_t* verdict, uint8_t* gt, size_t n) {
  FieldTables ft = make_ft();
  FsState fs_seed;
  memcpy(fs_seed.v, seed4, 16);
  VerifyKey k;
  G1* dst[9] = {&k.qm, &k.ql, &k.qr, &k.qo, &k.qc, &k.s1, &k.s2, &k.s3, &k.g1_one};
  for (int j = 0; j < 9; j++) *dst[j] = ld(key + 3 * j);
  k.g2_one = G2{key[27], key[28]};
  k.g2_s = G2{key[29], key[30]};
  static VerifyTables vt;
  build_vt(ft, dst, vt);
  static VerifyLogTables lt;
  uint8_t alog[104];
  memset(&lt, 0, sizeof lt);
  if (!vlt_group(ft, lt, alog)) return 1;
  for (uint32_t t = 0; t < VLT_ENTRIES; t++) vlt_entry(ft, k, vt, alog, lt, t);
  for (size_t i = 0; i < n; i++) {
    uint32_t pbv[27], op[7], ch[5];
    for (int j = 0; j < 27; j++) pbv[j] = proofs[34 * i + j];
    for (int j = 0; j < 7; j++) op[j] = proofs[34 * i + 27 + j];
    uint32_t uu;
    if (chal) { for (int j = 0; j < 5; j++) ch[j] = chal[5 * i + j]; uu = u[i]; }
    else fs_derive(fs_seed, pbv, op, ch, uu);
    VerifyOut o;
    if (gt) verify_one_log<true>(lt, pbv, op, ch, uu, o);
    else verify_one_log<false>(lt, pbv, op, ch, uu, o);
    verdict[i] = (uint8_t)o.verdict;
    if (gt) { gt[4 * i] = (uint8_t)o.lhs.a; gt[4 * i + 1] = (uint8_t)o.lhs.b; gt[4 * i + 2] = (uint8_t)o.rhs.a; gt[4 * i + 3] = (uint8_t)o.rhs.b; }
  }
  return 0;
}
// The premise of the table path, exhaustively: for ALL 102 x 102 pairs of curve points the reference's g1_add (and its
// canonical-point variant g1_add_c) is addition of discrete logarithms mod 102; g1_double, g1_neg and g1_mul by every
// scalar < 17 agree with it; the membership look-up equals g1_is_on_curve on all 101 x 101 x 2 encodings.  Returns the
// number of disagreements (0), or ~0 if no generator exists.
uint64_t hc_check_discrete_logs() {
  FieldTables ft = make_ft();
  static VerifyLogTables lt;
  uint8_t alog[104];
  memset(&lt, 0, sizeof lt);
  if (!vlt_group(ft, lt, alog)) return ~0ull;
  uint64_t bad = 0;
  auto pt = [&](uint32_t i) { return G1{vlt_px(lt, i), vlt_py(lt, i), i == 0u ? 1u : 0u}; };
  auto same = [&](G1 r, uint32_t k) { const uint32_t i = alog[k % GROUP_ORDER]; const G1 w = pt(i); return r.x == w.x && r.y == w.y && r.inf == w.inf; };
  for (uint32_t i = 0; i < GROUP_ORDER; i++) {
    bad += alog[lt.dlog[i]] != i;
    const G1 p = pt(i);
    const uint32_t ki = lt.dlog[i];
    bad += !same(g1_double(ft, p), 2u * ki) + !same(g1_double_c(ft, p), 2u * ki) + !same(g1_neg(p), GROUP_ORDER - ki);
    for (uint32_t sc = 0; sc < 17; sc++) bad += !same(g1_mul(ft, p, sc), sc * ki);
    for (uint32_t j = 0; j < GROUP_ORDER; j++) {
      const G1 q = pt(j);
      bad += !same(g1_add(ft, p, q), ki + lt.dlog[j]) + !same(g1_add_c(ft, p, q), ki + lt.dlog[j]);
    }
  }
  for (uint32_t x = 0; x < 101; x++)
    for (uint32_t y = 0; y < 101; y++) {
      const bool on = g1_is_on_curve(G1{x, y, 0u});
      const uint32_t i = lt.cbase[x] + (2u * y > 101u ? 1u : 0u);
      bad += on != (lt.pw[i] == (x | y << 8));
    }
  return bad;
}
// the shared-final-exponentiation identity behind pairings_equal17_c, checked on GT values directly:
// (f1^600 == f2^600)  ==  (either zero ? both zero : (f1 conj(f2))^600 == 1), f2 over ALL of GT, f1 over every `step`-th value
uint64_t hc_check_shared_final_exp(uint32_t step) {
  uint64_t bad = 0;
  for (uint32_t i = 0; i < 101u * 101u; i += step) {
    const GT f1{i / 101u, i % 101u};
    const GT e1 = final_exp600(f1);
    const bool z1 = (f1.a | f1.b) == 0u;
    for (uint32_t j = 0; j < 101u * 101u; j++) {
      const GT f2{j / 101u, j % 101u};
      const GT e2 = final_exp600(f2);
      const bool want = e1.a == e2.a && e1.b == e2.b;
      const bool z2 = (f2.a | f2.b) == 0u;
      const GT e = final_exp600(gt_mul(f1, gt_conj(f2)));
      const bool got = (z1 || z2) ? (z1 && z2) : (e.a == 1u && e.b == 0u);
      bad += want != got;
    }
  }
  return bad;
}
// The canonical-domain variants of the fast paths against the any-input functions, exhaustively on their domain:
// every pair of canonically encoded points of E(F_101) (101 finite points + the identity) for g1_add_c / g1_double_c,
// and every such point against every G2 byte pair (x, y < 101) for miller17_c.  Returns the number of mismatches.
uint64_t hc_check_canonical_variants(const uint8_t* pts /*[np][3]*/, uint32_t np) {
  const FieldTables ft = make_ft();
  uint64_t bad = 0;
  for (uint32_t i = 0; i < np; i++) {
    const G1 a = ld(pts + 3 * i);
    const G1 d0 = g1_double(ft, a), d1 = g1_double_c(ft, a);
    bad += d0.x != d1.x || d0.y != d1.y || d0.inf != d1.inf;
    for (uint32_t j = 0; j < np; j++) {
      const G1 b = ld(pts + 3 * j);
      const G1 r0 = g1_add(ft, a, b), r1 = g1_add_c(ft, a, b);
      bad += r0.x != r1.x || r0.y != r1.y || r0.inf != r1.inf;
    }
    for (uint32_t q = 0; q < 101u * 101u; q++) {
      const G2 Q{q / 101u, q % 101u};
      const GT f0 = miller17(ft, a, Q), f1 = miller17_c(ft, a, Q);
      bad += f0.a != f1.a || f0.b != f1.b;
    }
  }
  return bad;
}
// The premise of every table path: on canonically encoded curve points the reference's g1_add IS an abelian group law
// with one encoding per element -- commutative and associative as TRIPLES, identity {0,0,1} neutral.  All pairs, all triples.
uint64_t hc_check_group_law(const uint8_t* pts /*[np][3]*/, uint32_t np) {
  const FieldTables ft = make_ft();
  auto eq = [](const G1& p, const G1& q) { return p.x == q.x && p.y == q.y && p.inf == q.inf; };
  uint64_t bad = 0;
  const G1 id = g1_identity();
  for (uint32_t i = 0; i < np; i++) {
    const G1 a = ld(pts + 3 * i);
    bad += !eq(g1_add(ft, a, id), a) || !eq(g1_add(ft, id, a), a);
    bad += !eq(g1_add(ft, a, g1_neg(a)), id);
    bad += !eq(g1_add(ft, a, a), g1_double(ft, a));
    for (uint32_t j = 0; j < np; j++) {
      const G1 b = ld(pts + 3 * j);
      const G1 ab = g1_add(ft, a, b);
      bad += !eq(ab, g1_add(ft, b, a));
      bad += !g1_is_on_curve(ab) || ab.x > 100u || ab.y > 100u || (ab.inf && (ab.x | ab.y));   // closed, canonical
      for (uint32_t k = 0; k < np; k++) {
        const G1 c = ld(pts + 3 * k);
        bad += !eq(g1_add(ft, ab, c), g1_add(ft, a, g1_add(ft, b, c)));
      }
    }
  }
  return bad;
}
// Barrett reductions over their whole claimed range (field.cuh: red17 exact below 2^28, red101 below 2^26): the number of
// x with red(x) != x % p.  Also reports, through *first_bad, the smallest x >= the claimed bound at which each formula
// first fails (so the stated bounds can be seen to be safe, not tight guesses).
uint64_t hc_check_barrett(uint32_t* first_bad17, uint32_t* first_bad101) {
  uint64_t bad = 0;
  for (uint32_t x = 0; x < (1u << 28); x++) bad += red17(x) != x % 17u;
  for (uint32_t x = 0; x < (1u << 26); x++) bad += red101(x) != x % 101u;
  *first_bad17 = 0; *first_bad101 = 0;
  // beyond the claimed ranges the formulas are restated here (red17 / red101 themselves refuse such inputs in this build)
  auto f17 = [](uint32_t x) { return x - P17 * mulhi_u32(x, M17); };
  auto f101 = [](uint32_t x) { return x - P101 * mulhi_u32(x, M101); };
  for (uint64_t x = 1ull << 28; x < (1ull << 32); x++) if (f17((uint32_t)x) != (uint32_t)x % 17u) { *first_bad17 = (uint32_t)x; break; }
  for (uint64_t x = 1ull << 26; x < (1ull << 32); x++) if (f101((uint32_t)x) != (uint32_t)x % 101u) { *first_bad101 = (uint32_t)x; break; }
  return bad;
}
// packed wire v2 and the synthetic stream (wire.cuh): the device functions the kernels call, run on the host
void hc_synth(uint64_t seed, uint64_t start, int variant, const uint8_t* wtab, uint8_t* wit, uint8_t* rnd, uint8_t* chal, uint8_t* u,
              uint8_t* packed, size_t n) {
  for (size_t i = 0; i < n; i++) {
    uint32_t v[PACKED_VALUES], w[4];
    synth_item(seed, start + i, variant, wtab, v);
    for (int k = 0; k < 12; k++) wit[12 * i + k] = (uint8_t)v[k];
    for (int k = 0; k < 9; k++) rnd[9 * i + k] = (uint8_t)v[12 + k];
    for (int k = 0; k < 5; k++) chal[5 * i + k] = (uint8_t)v[21 + k];
    u[i] = (uint8_t)v[26];
    pack_input16(v, w);
    memcpy(packed + 16 * i, w, 16);
  }
}
// every 32-bit word through unpack7: returns the number of words whose digits / validity flag differ from plain division
uint64_t hc_check_unpack7(uint32_t lo, uint32_t hi, uint32_t step) {
  uint64_t bad = 0;
  for (uint64_t x = lo; x < hi; x += step) {
    uint32_t d[7];
    const bool ok = unpack7((uint32_t)x, d);
    uint32_t q = (uint32_t)x;
    bool same = true;
    for (int k = 0; k < 6; k++) { same = same && d[k] == q % 17u; q /= 17u; }
    same = same && d[6] == q && ok == (x < P17_7);
    if (ok) same = same && pack7(d) == (uint32_t)x;
    bad += !same;
  }
  return bad;
}
// prove_kernel<.., PACKED> front end + verifier word 3: packed record -> the values the kernels use, 0xFF where not an encoding
void hc_unpack_input16(const uint8_t* packed, uint8_t* v27, uint8_t* ok, size_t n) {
  for (size_t i = 0; i < n; i++) {
    uint32_t w[4], v[PACKED_VALUES];
    memcpy(w, packed + 16 * i, 16);
    ok[i] = unpack_input16(w[0], w[1], w[2], w[3], v) ? 1 : 0;
    for (int k = 0; k < PACKED_VALUES; k++) v27[PACKED_VALUES * i + k] = (uint8_t)v[k];
  }
}
int hc_bounds_checked() {
#ifdef PB_CHECK_BOUNDS
  return 1;
#else
  return 0;
#endif
}
int hc_sizeof_cc() { return (int)sizeof(CircuitConst); }
// key: 9 G1 as bytes [27] + g2[4]
void hc_verify(const uint8_t* key, const uint32_t seed4[4], const uint8_t* proofs, const uint8_t* chal, const uint8_t* u, uint8_t* verdict, uint8_t* gt, size_t n) {
  FieldTables ft = make_ft();
  FsState fs_seed;
  memcpy(fs_seed.v, seed4, 16);
  VerifyKey k;
  G1* dst[9] = {&k.qm, &k.ql, &k.qr, &k.qo, &k.qc, &k.s1, &k.s2, &k.s3, &k.g1_one};
  for (int j = 0; j < 9; j++) *dst[j] = ld(key + 3 * j);
  k.g2_one = G2{key[27], key[28]};
  k.g2_s = G2{key[29], key[30]};
  for (size_t i = 0; i < n; i++) {
    uint32_t pbv[27], op[7], ch[5];
    for (int j = 0; j < 27; j++) pbv[j] = proofs[34 * i + j];
    for (int j = 0; j < 7; j++) op[j] = proofs[34 * i + 27 + j];
    uint32_t uu;
    if (chal) { for (int j = 0; j < 5; j++) ch[j] = chal[5 * i + j]; uu = u[i]; }
    else fs_derive(fs_seed, pbv, op, ch, uu);
    VerifyOut o;
    verify_one(k, ft, pbv, op, ch, uu, o);
    verdict[i] = (uint8_t)o.verdict;
    if (gt) { gt[4 * i] = (uint8_t)o.lhs.a; gt[4 * i + 1] = (uint8_t)o.lhs.b; gt[4 * i + 2] = (uint8_t)o.rhs.a; gt[4 * i + 3] = (uint8_t)o.rhs.b; }
  }
}
}
