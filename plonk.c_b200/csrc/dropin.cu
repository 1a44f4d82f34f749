// dropin.cu -- the reference's header-level API (include/hf.h ... include/plonk.h) on top of the batch C ABI.
//
// Each function below replaces the same-named definition in plonk.c's headers (file:line cited per function).
// Arithmetic goes to the GPU through the pb_* entry points at batch size 1; memory ownership (libc malloc,
// released by the reference's *_free), in-place aliasing (poly_add_hf) and the error convention
// (message on stderr + exit(EXIT_FAILURE), or abort() where the reference asserts) are the reference's.
// Without a CUDA device every arithmetic call stops the process with the library's error text: there is no
// CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/plonk_b200.h"
extern "C" {
#include "../../include/plonk.h"
#include "../../include/pairing.h"
}

namespace {

[[noreturn]] void die(const char* msg) {           // the reference's convention: stderr + exit(EXIT_FAILURE)
  fprintf(stderr, "%s\n", msg);
  exit(EXIT_FAILURE);
}
void gpu(int rc) {                                  // a failed library call is fatal, loudly
  if (rc != PB_OK) {
    fprintf(stderr, "plonk_b200 drop-in: %s\n", pb_last_error());
    abort();
  }
}
POLY poly_own(const uint8_t* c, size_t len) {       // poly_new over raw bytes (already trimmed by the kernels)
  POLY p;
  p.len = len;
  p.coeffs = (HF*)malloc((len ? len : 1) * sizeof(HF));
  if (!p.coeffs) die("Memory allocation failed in poly_new");
  for (size_t i = 0; i < len; i++) p.coeffs[i].value = c[i];
  return p;
}
void need_len(const POLY* p, const char* who) {
  if (p->len == 0 || p->len > PB_POLY_MAX) {
    fprintf(stderr, "plonk_b200 drop-in: %s: polynomial length %zu outside [1, %d]\n", who, p->len, PB_POLY_MAX);
    abort();
  }
}
const uint8_t* bytes(const HF* p) { return reinterpret_cast<const uint8_t*>(p); }
static_assert(sizeof(HF) == 1 && sizeof(GF) == 1 && sizeof(G1) == 3 && sizeof(G2) == 2 && sizeof(GTP) == 2 && sizeof(PROOF) == 34 &&
              sizeof(CHALLENGE) == 5 && sizeof(LINE_EQ) == 3, "struct layouts must match the reference's");

POLY poly_binop(int op, const POLY* a, const POLY* b, const char* who) {
  need_len(a, who); need_len(b, who);
  uint8_t out[2 * PB_POLY_MAX], olen = 0;
  uint8_t la = (uint8_t)a->len, lb = (uint8_t)b->len;
  size_t so = op == PB_POLY_MUL ? a->len + b->len - 1 : (a->len > b->len ? a->len : b->len);
  gpu(pb_poly_binop(op, bytes(a->coeffs), &la, a->len, bytes(b->coeffs), &lb, b->len, out, &olen, so, 1));
  return poly_own(out, olen);
}
POLY poly_unop(int op, const POLY* p, uint8_t k, size_t so, const char* who) {
  need_len(p, who);
  uint8_t out[2 * PB_POLY_MAX], olen = 0, lp = (uint8_t)p->len;
  gpu(pb_poly_unop(op, bytes(p->coeffs), &lp, p->len, &k, out, &olen, so, 1));
  return poly_own(out, olen);
}
void g1_bytes(uint8_t* b, const G1* p) { b[0] = p->x.value; b[1] = p->y.value; b[2] = p->infinite ? 1 : 0; }
G1 g1_from(const uint8_t* b) { G1 p; p.x.value = b[0]; p.y.value = b[1]; p.infinite = b[2] != 0; return p; }

// contexts for plonk_prove are cached by (circuit, SRS) content
std::mutex g_ctx_mu;
std::map<std::string, pb_ctx*> g_ctx;
pb_ctx* ctx_for(const uint8_t circuit[PB_CIRCUIT_BYTES], const SRS* srs) {
  std::string key(reinterpret_cast<const char*>(circuit), PB_CIRCUIT_BYTES);
  std::vector<uint8_t> g1s(3 * srs->len);
  for (size_t i = 0; i < srs->len; i++) g1_bytes(&g1s[3 * i], &srs->g1s[i]);
  uint8_t g2[4] = {srs->g2_1.x.value, srs->g2_1.y.value, srs->g2_s.x.value, srs->g2_s.y.value};
  key.append(reinterpret_cast<const char*>(g1s.data()), g1s.size());
  key.append(reinterpret_cast<const char*>(g2), 4);
  // the caller holds g_ctx_mu (for the whole plonk_prove call: the context must not be evicted under it)
  auto it = g_ctx.find(key);
  if (it != g_ctx.end()) return it->second;
  // bounded: a program that proves against many (circuit, SRS) pairs does not accumulate contexts (each holds ~50 MB of
  // tables); drop-in proofs are serialised by g_ctx_mu, so no context is in use when the cache is emptied
  if (g_ctx.size() >= 8) {
    for (auto& kv : g_ctx) pb_ctx_destroy(kv.second);
    g_ctx.clear();
  }
  pb_ctx* c = nullptr;
  int dev = 0;
  gpu(pb_ctx_create(&c, dev, circuit, g1s.data(), (uint32_t)srs->len, g2));
  g_ctx[key] = c;
  return c;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ poly.h
POLY poly_add_hf(POLY* a, const HF b) {             // src/poly.h:67-70: in place, untrimmed, returns an alias
  uint8_t out[2 * PB_POLY_MAX], olen = 0, la = (uint8_t)a->len, k = b.value;
  need_len(a, "poly_add_hf");
  gpu(pb_poly_unop(PB_POLY_ADD_HF, bytes(a->coeffs), &la, a->len, &k, out, &olen, a->len, 1));
  a->coeffs[0].value = out[0];
  return *a;
}
POLY poly_add(const POLY* a, const POLY* b) { return poly_binop(PB_POLY_ADD, a, b, "poly_add"); }   // src/poly.h:72-87
POLY poly_sub(const POLY* a, const POLY* b) { return poly_binop(PB_POLY_SUB, a, b, "poly_sub"); }   // src/poly.h:89-104
POLY poly_mul(const POLY* a, const POLY* b) { return poly_binop(PB_POLY_MUL, a, b, "poly_mul"); }   // src/poly.h:106-122
void poly_divide(const POLY* num, const POLY* den, POLY* quot, POLY* rem) {                          // src/poly.h:124-177
  need_len(num, "poly_divide");
  if (den->len > PB_POLY_MAX) need_len(den, "poly_divide");
  if (poly_is_zero(den)) die("Division by zero polynomial in poly_divide");
  uint8_t q[PB_POLY_MAX], r[PB_POLY_MAX], ql = 0, rl = 0, st = 0, ln = (uint8_t)num->len, ld = (uint8_t)den->len;
  gpu(pb_poly_divide(bytes(num->coeffs), &ln, num->len, bytes(den->coeffs), &ld, den->len, q, &ql, PB_POLY_MAX, r, &rl, PB_POLY_MAX, &st, 1));
  if (st) die("Division by zero polynomial in poly_divide");
  *quot = poly_own(q, ql);
  *rem = poly_own(r, rl);                                                                             // len 0 for a constant divisor (hazard C-4)
}
POLY poly_scale(const POLY* p, HF scalar) { return poly_unop(PB_POLY_SCALE, p, scalar.value, p->len, "poly_scale"); }      // src/poly.h:179-197
POLY poly_shift(const POLY* p, size_t shift) {                                                                            // src/poly.h:199-216
  if (p->len + shift > 2 * PB_POLY_MAX) { fprintf(stderr, "plonk_b200 drop-in: poly_shift result too long\n"); abort(); }
  return poly_unop(PB_POLY_SHIFT, p, (uint8_t)shift, p->len + shift, "poly_shift");
}
POLY poly_slice(const POLY* p, size_t start, size_t end) {                                                               // src/poly.h:218-238
  if (start >= end || end > p->len) die("Invalid slice indices in poly_slice");
  need_len(p, "poly_slice");
  uint8_t out[PB_POLY_MAX], olen = 0, st = 0, lp = (uint8_t)p->len, s = (uint8_t)start, e = (uint8_t)end;
  gpu(pb_poly_slice(bytes(p->coeffs), &lp, p->len, &s, &e, out, &olen, p->len, &st, 1));
  if (st) die("Invalid slice indices in poly_slice");
  return poly_own(out, olen);
}
POLY poly_negate(const POLY* p) { return poly_unop(PB_POLY_NEGATE, p, 0, p->len, "poly_negate"); }                        // src/poly.h:240-254
HF poly_eval(const POLY* p, HF x) {                                                                                       // src/poly.h:265-272
  HF y = hf_zero();
  if (p->len == 0) return y;
  need_len(p, "poly_eval");
  uint8_t lp = (uint8_t)p->len;
  gpu(pb_poly_eval(bytes(p->coeffs), &lp, p->len, &x.value, &y.value, 1));
  return y;
}
POLY poly_z(const HF* points, size_t len) {                                                                               // src/poly.h:274-286
  POLY acc = poly_one();
  for (size_t i = 0; i < len; i++) {
    HF lin[2] = {hf_neg(points[i]), hf_one()};
    POLY term = poly_new(lin, 2), next = poly_mul(&acc, &term);
    poly_free(&acc); poly_free(&term);
    acc = next;
  }
  return acc;
}
POLY poly_lagrange(const HF* x_points, const HF* y_points, size_t len) {                                                  // src/poly.h:288-321
  if (len == 0) return poly_zero();
  if (len > 16) { fprintf(stderr, "plonk_b200 drop-in: poly_lagrange supports at most 16 points\n"); abort(); }
  uint8_t out[16], olen = 0, st = 0;
  gpu(pb_poly_lagrange(bytes(x_points), bytes(y_points), len, out, &olen, len, &st, 1));
  if (st) die("Error: Lagrange polynomial x points must be unique");
  return poly_own(out, olen);
}

// ------------------------------------------------------------------ matrix.h
MATRIX matrix_zero(size_t m, size_t n) {                                                                                  // src/matrix.h:15-25
  MATRIX r = {m, n, (HF*)calloc(m * n ? m * n : 1, sizeof(HF))};
  if (!r.v) die("Memory allocation failed in matrix_zero");
  return r;
}
MATRIX matrix_new(HF* v, size_t m, size_t n) {                                                                            // src/matrix.h:27-40
  MATRIX r = matrix_zero(m, n);
  memcpy(r.v, v, m * n * sizeof(HF));
  return r;
}
HF matrix_get(const MATRIX* a, size_t row, size_t col) {                                                                  // src/matrix.h:43-49
  if (row >= a->m || col >= a->n) die("Index out of bounds in matrix_get");
  return a->v[col + row * a->n];
}
void matrix_set(MATRIX* a, size_t row, size_t col, HF value) {                                                            // src/matrix.h:52-58
  if (row >= a->m || col >= a->n) die("Index out of bounds in matrix_set");
  a->v[col + row * a->n] = value;
}
void matrix_free(MATRIX* a) { free(a->v); a->v = NULL; a->m = 0; a->n = 0; }                                              // src/matrix.h:60-67
MATRIX matrix_add(const MATRIX* a, const MATRIX* b) {                                                                     // src/matrix.h:69-79
  if (a->m != b->m || a->n != b->n) die("Matrix dimensions must match for additoin");
  MATRIX r = matrix_zero(a->m, a->n);
  gpu(pb_field_op(17, PB_OP_ADD, bytes(a->v), bytes(b->v), reinterpret_cast<uint8_t*>(r.v), a->m * a->n));
  return r;
}
MATRIX matrix_mul(const MATRIX* a, const MATRIX* b) {                                                                     // src/matrix.h:81-98
  if (a->n != b->m) {
    fprintf(stderr, "Matrix multiplication error: Dimensions (%zu x %zu) and (%zu x %zu) incompatible.\n", a->m, a->n, b->m, b->n);
    exit(EXIT_FAILURE);
  }
  MATRIX r = matrix_zero(a->m, b->n);
  gpu(pb_matrix_mul(bytes(a->v), bytes(b->v), reinterpret_cast<uint8_t*>(r.v), (uint32_t)a->m, (uint32_t)a->n, (uint32_t)b->n, 1));
  return r;
}
void matrix_gauss_jordan(MATRIX* a) {                                                                                    // src/matrix.h:100-149
  gpu(pb_matrix_gauss_jordan(reinterpret_cast<uint8_t*>(a->v), (uint32_t)a->m, (uint32_t)a->n, 1));
}
MATRIX matrix_inv(const MATRIX* a) {                                                                                      // src/matrix.h:151-176
  if (a->m != a->n) die("Only square matrices can be inverted");
  MATRIX r = matrix_zero(a->n, a->n);
  gpu(pb_matrix_inv(bytes(a->v), reinterpret_cast<uint8_t*>(r.v), (uint32_t)a->n, 1));
  return r;
}

// ------------------------------------------------------------------ g1.h
G1 g1_new(uint64_t x, uint64_t y) { G1 p; p.x = f101((int64_t)x); p.y = f101((int64_t)y); p.infinite = false; return p; }     // src/g1.h:13-20
G1 g1_generator(void) { return g1_new(1, 2); }                                                                            // src/g1.h:22-24
G1 g1_identity(void) { G1 p; p.x.value = 0; p.y.value = 0; p.infinite = true; return p; }                                 // src/g1.h:33-35
GF g1_generator_subgroup_size(void) { return f101(17); }                                                                  // src/g1.h:105-107
bool g1_is_on_curve(const G1* p) {                                                                                        // src/g1.h:26-31
  uint8_t b[3], out = 0;
  g1_bytes(b, p);
  gpu(pb_g1_is_on_curve(b, &out, 1));
  return out != 0;
}
static G1 g1_call(int op, const G1* a, const G1* b) {
  uint8_t x[3], y[3] = {0, 0, 0}, out[3];
  g1_bytes(x, a);
  if (b) g1_bytes(y, b);
  gpu(pb_g1_op(op, x, y, out, 1));
  return g1_from(out);
}
G1 g1_double(const G1* a) { return g1_call(PB_G_DOUBLE, a, nullptr); }                                                    // src/g1.h:37-56
G1 g1_add(const G1* a, const G1* b) { return g1_call(PB_G_ADD, a, b); }                                                   // src/g1.h:59-83
G1 g1_neg(G1* a) { return g1_call(PB_G_NEG, a, nullptr); }                                                                // src/g1.h:85-89
G1 g1_mul(const G1* p, uint64_t scalar) {                                                                                 // src/g1.h:91-103
  uint8_t x[3], out[3];
  g1_bytes(x, p);
  gpu(pb_g1_mul(x, &scalar, out, 1));
  return g1_from(out);
}

// ------------------------------------------------------------------ g2.h
G2 g2_new(uint64_t x, uint64_t y) { G2 p; p.x = f101((int64_t)x); p.y = f101((int64_t)y); return p; }                         // src/g2.h:11-17
G2 g2_generator(void) { return g2_new(36, 31); }                                                                          // src/g2.h:19-21
uint64_t g2_embedding_degree(void) { return 2; }                                                                          // src/g2.h:23-25
static G2 g2_call(int op, const G2* a, const G2* b) {
  uint8_t x[2] = {a->x.value, a->y.value}, y[2] = {0, 0}, out[2];
  if (b) { y[0] = b->x.value; y[1] = b->y.value; }
  gpu(pb_g2_op(op, x, y, out, 1));
  G2 r; r.x.value = out[0]; r.y.value = out[1];
  return r;
}
G2 g2_neg(G2* p) { return g2_call(PB_G_NEG, p, nullptr); }                                                                // src/g2.h:27-30
G2 g2_add(const G2* p, const G2* q) { return g2_call(PB_G_ADD, p, q); }                                                   // src/g2.h:32-66
G2 g2_mul(G2 base, uint64_t scalar) {                                                                                     // src/g2.h:68-84 (scalar 0: undefined there)
  uint8_t x[2] = {base.x.value, base.y.value}, out[2];
  gpu(pb_g2_mul(x, &scalar, out, 1));
  G2 r; r.x.value = out[0]; r.y.value = out[1];
  return r;
}

// ------------------------------------------------------------------ gt.h / pairing.h
GTP gtp_new(GF a, GF b) { GTP p; p.a = a; p.b = b; return p; }                                                            // src/gt.h:11-17
GTP gtp_neg(GTP* p) { return gtp_new(p->a, gf_neg(p->b)); }                                                               // src/gt.h:19-21
GTP gtp_mul(GTP* base, GTP* rhs) {                                                                                        // src/gt.h:23-28
  uint8_t x[2] = {base->a.value, base->b.value}, y[2] = {rhs->a.value, rhs->b.value}, out[2];
  gpu(pb_gtp_mul(x, y, out, 1));
  GTP r; r.a.value = out[0]; r.b.value = out[1];
  return r;
}
GTP gtp_pow(GTP* base, uint64_t exp) {                                                                                    // src/gt.h:30-51
  uint8_t x[2] = {base->a.value, base->b.value}, out[2];
  gpu(pb_gtp_pow(x, &exp, out, 1));
  GTP r; r.a.value = out[0]; r.b.value = out[1];
  return r;
}
int gtp_equal(const GTP* p, const GTP* q) { return gf_equal(p->a, q->a) && gf_equal(p->b, q->b); }                        // src/pairing.h:9-11
LINE_EQ line(const G1* a, const G1* b) {                                                                                  // src/pairing.h:19-29
  uint8_t x[3], y[3], out[3];
  g1_bytes(x, a); g1_bytes(y, b);
  gpu(pb_line(x, y, out, 1));
  LINE_EQ l; l.x.value = out[0]; l.y.value = out[1]; l.c.value = out[2];
  return l;
}
GTP pairing_f(uint64_t r, const G1* p, const G2* q) {                                                                     // src/pairing.h:31-64
  uint8_t x[3], y[2] = {q->x.value, q->y.value}, out[2];
  g1_bytes(x, p);
  gpu(pb_pairing_f(r, x, y, out, 1));
  GTP g; g.a.value = out[0]; g.b.value = out[1];
  return g;
}
GTP pairing(const G1* p, const G2* q) {                                                                                   // src/pairing.h:66-83
  uint8_t x[3], y[2] = {q->x.value, q->y.value}, out[2];
  g1_bytes(x, p);
  gpu(pb_pairing(x, y, out, 1));
  GTP g; g.a.value = out[0]; g.b.value = out[1];
  return g;
}

// ------------------------------------------------------------------ srs.h
SRS srs_create(GF secret, size_t n) {                                                                                     // src/srs.h:18-43
  SRS srs;
  srs.len = n + 1;
  srs.g1s = (G1*)malloc(srs.len * sizeof(G1));
  if (!srs.g1s) die("Mamory allocation failed in srs_create");
  // g1s[i] = g1_mul(identity, secret^(i+1)): multiples of the IDENTITY, as the reference computes (src/srs.h:27-35)
  std::vector<uint8_t> pts(3 * srs.len), out(3 * srs.len);
  std::vector<uint64_t> sc(srs.len);
  GF s_pow = secret;
  for (size_t i = 0; i < srs.len; i++) {
    pts[3 * i] = 0; pts[3 * i + 1] = 0; pts[3 * i + 2] = 1;
    sc[i] = s_pow.value;
    s_pow = gf_mul(s_pow, secret);
  }
  gpu(pb_g1_mul(pts.data(), sc.data(), out.data(), srs.len));
  for (size_t i = 0; i < srs.len; i++) srs.g1s[i] = g1_from(&out[3 * i]);
  srs.g2_1 = g2_generator();
  srs.g2_s = g2_mul(srs.g2_1, secret.value);
  return srs;
}
void srs_free(SRS* srs) { free(srs->g1s); srs->g1s = NULL; srs->len = 0; }                                                // src/srs.h:45-51
G1 srs_eval_at_s(const SRS* srs, const POLY* vs) {                                                                        // src/srs.h:53-68
  if (vs->len > srs->len) {
    fprintf(stderr, "Poynomial degree exceeds SRS size: POLY degree: %zu, SRS supports up to degree: %zu \n", vs->len, vs->len);
    exit(EXIT_FAILURE);
  }
  // one launch: g1_mul(g1s[i], coeff_i) per term and the reference's left-to-right sum, on the device
  if (vs->len == 0) return g1_identity();
  if (vs->len > PB_POLY_MAX) need_len(vs, "srs_eval_at_s");
  std::vector<uint8_t> pts(3 * vs->len);
  for (size_t i = 0; i < vs->len; i++) g1_bytes(&pts[3 * i], &srs->g1s[i]);
  uint8_t out[3], st = 0, len = (uint8_t)vs->len;
  gpu(pb_srs_eval_at_s_raw(pts.data(), (uint32_t)vs->len, bytes(vs->coeffs), &len, vs->len, /*trim=*/0, out, &st, 1));
  return g1_from(out);
}

// ------------------------------------------------------------------ constraints.h (host-side circuit authoring)
GATE gate_new(HF q_l, HF q_r, HF q_o, HF q_m, HF q_c) { GATE g = {q_l, q_r, q_o, q_m, q_c}; return g; }                    // src/constraints.h:84-87
GATE gate_sum_a_b(void) { return gate_new(hf_one(), hf_one(), f17(-1), hf_zero(), hf_zero()); }                           // a + b - c
GATE gate_sub_a_b(void) { return gate_new(hf_one(), f17(-1), f17(-1), hf_zero(), hf_zero()); }                            // a - b - c
GATE gate_mul_a_b(void) { return gate_new(hf_zero(), hf_zero(), f17(-1), hf_one(), hf_zero()); }                          // a b - c
GATE gate_bind_a(HF value) { return gate_new(hf_one(), hf_zero(), hf_zero(), hf_zero(), value); }                         // a + q_c
GATE gate_bind_to_zero(void) { return gate_new(hf_zero(), hf_zero(), hf_one(), hf_zero(), hf_zero()); }                   // c
CONSTRAINTS constraints_new(GATE* gates, size_t num_gates, COPY_OF* c_a, COPY_OF* c_b, COPY_OF* c_c, size_t num_constraints) {   // src/constraints.h:114-143
  CONSTRAINTS k;
  HF** sel[5] = {&k.q_l, &k.q_r, &k.q_o, &k.q_m, &k.q_c};
  for (auto s : sel) *s = (HF*)malloc((num_gates ? num_gates : 1) * sizeof(HF));
  for (size_t i = 0; i < num_gates; i++) {
    k.q_l[i] = gates[i].q_l; k.q_r[i] = gates[i].q_r; k.q_o[i] = gates[i].q_o; k.q_m[i] = gates[i].q_m; k.q_c[i] = gates[i].q_c;
  }
  k.num_gates = num_gates;
  COPY_OF* src[3] = {c_a, c_b, c_c};
  COPY_OF** dst[3] = {&k.c_a, &k.c_b, &k.c_c};
  for (int s = 0; s < 3; s++) {
    *dst[s] = (COPY_OF*)malloc((num_constraints ? num_constraints : 1) * sizeof(COPY_OF));
    memcpy(*dst[s], src[s], num_constraints * sizeof(COPY_OF));
  }
  k.num_constraints = num_constraints;
  return k;
}
bool constraints_satisfy(const CONSTRAINTS* c, const ASSIGNMENTS* a) {                                                    // src/constraints.h:145-171
  // the gate equations are evaluated on the GPU (pb_constraints_satisfy_rows); the first failing row is reported on stdout, like the reference
  const size_t rows = c->num_constraints;
  if (rows == 0) return true;
  std::vector<uint8_t> q(5 * rows);
  const HF* sel[5] = {c->q_l, c->q_r, c->q_o, c->q_m, c->q_c};
  for (int s = 0; s < 5; s++) memcpy(&q[s * rows], sel[s], rows);
  int32_t bad = -1;
  gpu(pb_constraints_satisfy_rows(q.data(), (uint32_t)rows, bytes(a->a), bytes(a->b), bytes(a->c), &bad, 1));
  if (bad >= 0) { printf("Constraint %zu not satisfied.\n", (size_t)bad); return false; }
  return true;
}
void constraints_free(CONSTRAINTS* k) {                                                                                   // src/constraints.h:173-183
  free(k->q_l); free(k->q_r); free(k->q_o); free(k->q_m); free(k->q_c);
  free(k->c_a); free(k->c_b); free(k->c_c);
}
void var_map_init(VAR_MAP* vm) { vm->count = 0; }                                                                         // src/constraints.h:193-195
size_t var_map_get_or_add(VAR_MAP* vm, const char* name) {                                                                // src/constraints.h:197-213
  for (size_t i = 0; i < vm->count; i++)
    if (!strcmp(vm->names[i], name)) return vm->indices[i];
  char* copy = strdup(name);
  if (!copy) die("Memory allocation failed in var_map_get_or_add");
  vm->names[vm->count] = copy;
  vm->indices[vm->count] = vm->count;
  return vm->indices[vm->count++];
}
const char* var_map_get_name(VAR_MAP* vm, size_t index) { return index < vm->count ? vm->names[index] : NULL; }           // src/constraints.h:215-219
void var_map_free(VAR_MAP* vm) { for (size_t i = 0; i < vm->count; i++) free(vm->names[i]); vm->count = 0; }              // src/constraints.h:221-225
void gate_list_init(GATE_LIST* gl) {                                                                                      // src/constraints.h:236-243
  gl->capacity = 10;
  gl->num_gates = 0;
  gl->gates = (GATE*)malloc(gl->capacity * sizeof(GATE));
  gl->a_indices = (size_t*)malloc(gl->capacity * sizeof(size_t));
  gl->b_indices = (size_t*)malloc(gl->capacity * sizeof(size_t));
  gl->c_indices = (size_t*)malloc(gl->capacity * sizeof(size_t));
}
void gate_list_append(GATE_LIST* gl, GATE g, size_t ai, size_t bi, size_t ci) {                                           // src/constraints.h:245-264
  if (gl->num_gates >= gl->capacity) {
    gl->capacity *= 2;
    gl->gates = (GATE*)realloc(gl->gates, gl->capacity * sizeof(GATE));
    gl->a_indices = (size_t*)realloc(gl->a_indices, gl->capacity * sizeof(size_t));
    gl->b_indices = (size_t*)realloc(gl->b_indices, gl->capacity * sizeof(size_t));
    gl->c_indices = (size_t*)realloc(gl->c_indices, gl->capacity * sizeof(size_t));
    if (!gl->gates || !gl->a_indices || !gl->b_indices || !gl->c_indices) die("Memory reallocation failed in gate_list_append");
  }
  size_t k = gl->num_gates++;
  gl->gates[k] = g; gl->a_indices[k] = ai; gl->b_indices[k] = bi; gl->c_indices[k] = ci;
}
void gate_list_free(GATE_LIST* gl) { free(gl->gates); free(gl->a_indices); free(gl->b_indices); free(gl->c_indices); }    // src/constraints.h:266-271
size_t eval_expr(EXPRESSION* e, VAR_MAP* vars, GATE_LIST* gates) {                                                        // src/constraints.h:273-309
  char name[24];
  switch (e->type) {
    case EXPR_VAR:
      return var_map_get_or_add(vars, e->data.var_name);
    case EXPR_CONST:
      snprintf(name, sizeof name, "const_%u", e->data.const_value.value);
      return var_map_get_or_add(vars, name);
    case EXPR_SUM: case EXPR_SUB: case EXPR_MUL: {
      size_t l = eval_expr(e->data.binary.left, vars, gates), r = eval_expr(e->data.binary.right, vars, gates);
      size_t out = vars->count;
      snprintf(name, sizeof name, "v%zu", out);
      var_map_get_or_add(vars, name);
      gate_list_append(gates, e->type == EXPR_SUM ? gate_sum_a_b() : e->type == EXPR_SUB ? gate_sub_a_b() : gate_mul_a_b(), l, r, out);
      return out;
    }
  }
  die("Unknown expression type");
}

// ------------------------------------------------------------------ plonk.h
PLONK plonk_new(SRS srs, size_t n) {                                                                                      // src/plonk.h:53-119
  PLONK pk;
  pk.srs = srs;
  pk.h_len = n;
  const HF omega = hf_new(OMEGA_VALUE), k1 = hf_new(K1_VALUE), k2 = hf_new(K2_VALUE);
  pk.h = (HF*)malloc((n ? n : 1) * sizeof(HF));
  pk.k1_h = (HF*)malloc((n ? n : 1) * sizeof(HF));
  pk.k2_h = (HF*)malloc((n ? n : 1) * sizeof(HF));
  if (!pk.h || !pk.k1_h || !pk.k2_h) die("Memory allocation failed in plonk_new");
  for (size_t i = 0; i < n; i++) pk.h[i] = hf_pow(omega, (uint8_t)i);          // the reference's loop counter is a uint8_t (src/plonk.h:69)
  for (size_t i = 0; i < n; i++)
    if (hf_equal(pk.h[i], k1) || hf_equal(pk.h[i], k2)) die("K1 or K2 is in H, which is not allowed");
  for (size_t i = 0; i < n; i++) pk.k1_h[i] = hf_mul(pk.h[i], k1);
  for (size_t i = 0; i < n; i++)
    if (hf_equal(pk.k1_h[i], k2)) die("K1 or K2 is in H, which is not allowed");
  for (size_t i = 0; i < n; i++) pk.k2_h[i] = hf_mul(pk.h[i], k2);
  MATRIX vdm = matrix_zero(n, n);                                              // h[r]^c, inverted on the GPU
  for (size_t c = 0; c < n; c++)
    for (size_t r = 0; r < n; r++) matrix_set(&vdm, r, c, hf_pow(pk.h[r], c));
  pk.h_pows_inv = matrix_inv(&vdm);
  matrix_free(&vdm);
  pk.z_h_x = poly_z(pk.h, n);
  return pk;
}
void plonk_free(PLONK* pk) {                                                                                              // src/plonk.h:121-140
  srs_free(&pk->srs);
  matrix_free(&pk->h_pows_inv);
  poly_free(&pk->z_h_x);
  free(pk->h); pk->h = NULL;
  free(pk->k1_h); pk->k1_h = NULL;
  free(pk->k2_h); pk->k2_h = NULL;
}
void copy_constraints_to_roots(const PLONK* pk, const COPY_OF* copy_of, size_t len, HF* sigma) {                          // src/plonk.h:142-160
  for (size_t i = 0; i < len; i++) {
    const size_t at = copy_of[i].index - 1;
    const HF* col = copy_of[i].type == COPYOF_A ? pk->h : copy_of[i].type == COPYOF_B ? pk->k1_h : copy_of[i].type == COPYOF_C ? pk->k2_h : NULL;
    if (!col) die("Invalid copy_of type");
    sigma[i] = col[at];
  }
}
POLY interpolate_at_h(const PLONK* pk, const HF* values, size_t len) {                                                    // src/plonk.h:162-195
  if (len != pk->h_len) {
    fprintf(stderr, "Length mismatch in interpolate_at_h: len= %zu, plonk->h_len = %zu\n", len, pk->h_len);
    exit(EXIT_FAILURE);
  }
  MATRIX col = matrix_zero(len, 1);
  memcpy(col.v, values, len * sizeof(HF));
  MATRIX prod = matrix_mul(&pk->h_pows_inv, &col);                             // h_pows_inv * values on the GPU
  POLY out = poly_new(prod.v, prod.m);
  matrix_free(&col);
  matrix_free(&prod);
  return out;
}
void poly_print(const POLY* p) {                                                                                          // src/plonk.h:197-220
  bool any = false;
  for (size_t i = 0; i < p->len; i++) {
    unsigned c = p->coeffs[i].value;
    if (!c) continue;
    if (any) printf("+");
    if (i == 0) printf("%u", c);
    else if (i == 1) printf("%ux", c);
    else { if (c != 1) printf("%u", c); printf("x^%zu", i); }
    any = true;
  }
  printf(any ? "\n" : "0\n");
}
PROOF plonk_prove(PLONK* pk, CONSTRAINTS* cs, ASSIGNMENTS* as, CHALLENGE* ch, HF rnd[9]) {                                // src/plonk.h:223-656
  if (cs->num_constraints != 4 || pk->h_len != 4) {
    fprintf(stderr, "plonk_b200 drop-in: plonk_prove supports the reference's domain only (h_len = num_constraints = 4; omega = 4 has order 4)\n");
    abort();
  }
  uint8_t circuit[PB_CIRCUIT_BYTES];
  const HF* sel[5] = {cs->q_l, cs->q_r, cs->q_o, cs->q_m, cs->q_c};
  for (int s = 0; s < 5; s++) for (int i = 0; i < 4; i++) circuit[4 * s + i] = sel[s][i].value;
  const COPY_OF* cp[3] = {cs->c_a, cs->c_b, cs->c_c};
  for (int s = 0; s < 3; s++) for (int i = 0; i < 4; i++) { circuit[20 + 8 * s + i] = (uint8_t)cp[s][i].type; circuit[24 + 8 * s + i] = (uint8_t)cp[s][i].index; }
  std::lock_guard<std::mutex> lock(g_ctx_mu);
  pb_ctx* ctx = ctx_for(circuit, &pk->srs);
  uint8_t wit[12], proof[PB_PROOF_BYTES], status = 0;
  for (int i = 0; i < 4; i++) { wit[i] = as->a[i].value; wit[4 + i] = as->b[i].value; wit[8 + i] = as->c[i].value; }
  gpu(pb_plonk_prove(ctx, wit, bytes(rnd), reinterpret_cast<const uint8_t*>(ch), proof, &status, 1));
  switch (status) {                               // the reference's exits, in its order
    case PB_PROVE_OK: break;
    case PB_PROVE_UNSATISFIED: constraints_satisfy(cs, as); /* prints the failing row */ /* fallthrough */
    case PB_PROVE_ACC_ASSERT: case PB_PROVE_OPENING_ASSERT:
      fprintf(stderr, "plonk_prove: assertion failed (reference exit path %u)\n", status);
      abort();
    case PB_PROVE_BAD_COPY_TYPE: die("Invalid copy_of type");
    case PB_PROVE_REMAINDER: die("Non-zero remainder in t(x) division");
    case PB_PROVE_SLICE: die("Invalid slice indices in poly_slice");
    case PB_PROVE_SRS_ABC: case PB_PROVE_SRS_Z: case PB_PROVE_SRS_T: case PB_PROVE_SRS_W: die("Poynomial degree exceeds SRS size");
    default:
      fprintf(stderr, "plonk_b200 drop-in: plonk_prove: input bytes are not F17 residues (status %u)\n", status);
      abort();
  }
  PROOF out;
  memcpy(&out, proof, sizeof out);
  return out;
}

}  // extern "C"
