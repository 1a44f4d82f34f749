/*
 * plonk_b200.h -- C ABI of the B200-native batched prove/verify path of plonk.c.
 *
 * The reference (kazuakiishiguro/plonk.c) has no FFI: its API is ten headers that define their
 * functions in place (src/hf.h ... src/plonk.h) and are #included into one translation unit.
 * This library is the batched equivalent of those functions.  Every entry point below names the
 * reference function whose per-item result it reproduces byte for byte; the drop-in scalar headers
 * in this directory (hf.h gf.h poly.h matrix.h g1.h g2.h gt.h pairing.h srs.h constraints.h plonk.h)
 * route the reference's heavyweight functions through these entry points at batch size 1.
 *
 * Conventions
 *  - plain pointers and sizes only; all data are bytes (one field element per byte, canonical
 *    residues: F17 values in [0,17), F101 values in [0,101)) unless stated otherwise.
 *  - G1 is the reference's struct {GF x; GF y; bool infinite} = 3 bytes (g1.h:8-11), G2 is
 *    {GF x; GF y} = 2 bytes (g2.h:7-9), GTP is {GF a; GF b} = 2 bytes (gt.h:7-9), PROOF is 34 bytes
 *    (plonk.h:24-41), CHALLENGE is 5 bytes (plonk.h:16-22).  Batches are arrays of these structs.
 *  - `*_dev` entry points take DEVICE pointers and enqueue on `stream` (a cudaStream_t passed as
 *    void*, NULL = default stream) without synchronising.  The entry points without the suffix take
 *    HOST pointers, copy in, run, copy out and synchronise.
 *  - return value: 0 on success, negative pb_status on failure (pb_last_error() has the text).
 *    There is no CPU fallback: without a CUDA device every compute entry point fails with
 *    PB_ERR_NO_DEVICE.
 *  - where the reference would exit() or abort() on an item, the batch entry point writes a per-item
 *    status byte instead (SURVEY.md Appendix B numbering for plonk_prove) and zero-fills the item.
 *  - per-item bytes are validated on the device.  A polynomial row whose length byte exceeds its stride, a
 *    coefficient / value / scalar byte above 16 where a field element of F17 is expected, a shift that does not
 *    fit the output row: none of these is a value any reference path produces (HF is "always kept in range",
 *    hf.h:11-14; poly_new exits on a bad length, poly.h:29-32).  Such an item is REPORTED and never computed on:
 *    status[i] = 2 where the entry point has a status array, olen[i] = 0 with a zero row where the result has a
 *    length (every valid result has len >= 1 there), 0xFF from poly_eval, PB_PROVE_BAD_INPUT from the prover.
 *    The other items of the batch are unaffected.  (pb_config2_items, the fused fixed-shape entry point, and the
 *    group / field element-wise kernels do not validate: any byte is memory-safe there and yields a
 *    deterministic, unspecified value.)
 *  - `*_dev` entry points that take a context must be called with the context's device current
 *    (cudaSetDevice); otherwise PB_ERR_ARG.
 */
#ifndef PLONK_B200_H
#define PLONK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PB_OK = 0,
  PB_ERR_NO_DEVICE = -1, /* no CUDA device / driver: the product path refuses to run */
  PB_ERR_CUDA = -2,      /* a CUDA runtime call failed */
  PB_ERR_ARG = -3        /* invalid argument (null pointer, unsupported size, malformed circuit) */
} pb_status;

/* per-item status of pb_plonk_prove*: the SURVEY.md Appendix-B row of the first reference exit that
 * fires, in the reference's execution order.  0 = the reference returns a PROOF. */
enum {
  PB_PROVE_OK = 0,
  PB_PROVE_UNSATISFIED = 1,     /* assert(constraints_satisfy) plonk.h:231                        */
  PB_PROVE_BAD_COPY_TYPE = 3,   /* "Invalid copy_of type" plonk.h:155-157                         */
  PB_PROVE_SRS_ABC = 5,         /* a/b/c longer than the SRS: srs.h:54-57 via plonk.h:299-301     */
  PB_PROVE_ACC_ASSERT = 6,      /* assert(acc_x(omega^n) == 1) plonk.h:368                        */
  PB_PROVE_SRS_Z = 7,           /* z longer than the SRS: plonk.h:379                             */
  PB_PROVE_REMAINDER = 8,       /* "Non-zero remainder in t(x) division" plonk.h:507-510          */
  PB_PROVE_SLICE = 9,           /* "Invalid slice indices" poly.h:219-222 via plonk.h:517-519     */
  PB_PROVE_SRS_T = 10,          /* t_lo/t_mid/t_hi longer than the SRS: plonk.h:522-524           */
  PB_PROVE_OPENING_ASSERT = 11, /* assert(poly_is_zero(rem)) plonk.h:610,617                      */
  PB_PROVE_SRS_W = 12,          /* W_z / W_zw longer than the SRS: plonk.h:620-621                */
  PB_PROVE_BAD_INPUT = 254      /* an input byte is not a canonical F17 residue (not a reference path) */
};

/* verdicts of pb_plonk_verify* (the verifier is new: the reference has none, plonk.h:656-659) */
enum { PB_VERIFY_REJECT = 0, PB_VERIFY_ACCEPT = 1, PB_VERIFY_BAD_POINT = 2, PB_VERIFY_BAD_SCALAR = 3 };

/* field / group operation selectors */
enum { PB_OP_ADD = 0, PB_OP_SUB = 1, PB_OP_MUL = 2, PB_OP_DIV = 3, PB_OP_NEG = 4, PB_OP_INV = 5, PB_OP_POW = 6 };
enum { PB_POLY_ADD = 0, PB_POLY_SUB = 1, PB_POLY_MUL = 2 };
enum { PB_POLY_SCALE = 0, PB_POLY_NEGATE = 1, PB_POLY_SHIFT = 2, PB_POLY_ADD_HF = 3 };
enum { PB_G_ADD = 0, PB_G_DOUBLE = 1, PB_G_NEG = 2 };

#define PB_CIRCUIT_BYTES 44 /* q_l[4] q_r[4] q_o[4] q_m[4] q_c[4] | c_a.type[4] c_a.index[4] | c_b.. | c_c.. */
#define PB_WITNESS_BYTES 12 /* ASSIGNMENTS a[4] b[4] c[4] (constraints.h:57-62) */
#define PB_RAND_BYTES 9     /* HF rand[9] (plonk.h:228) */
#define PB_CHALLENGE_BYTES 5
#define PB_PROOF_BYTES 34
#define PB_POLY_MAX 64      /* longest polynomial the generic poly entry points accept */
#define PB_SRS_MAX 64       /* longest SRS a context accepts */

const char *pb_last_error(void);
int pb_abi_version(void);
int pb_device_count(void); /* >= 0, or PB_ERR_NO_DEVICE */

/* pinned host memory for the host-pointer entry points (optional; any host memory works) */
int pb_host_alloc(void **out, size_t bytes);
int pb_host_free(void *p);

/* ---- context: one circuit + one SRS on one device.  Replaces plonk_new (plonk.h:53-119) plus the
 * circuit-constant part of plonk_prove (plonk.h:254-275: sigma mapping and the eleven interpolations
 * that do not depend on the witness).  h_len is fixed at 4 (omega = 4 has order 4 in F17, plonk.h:12). */
typedef struct pb_ctx pb_ctx;
int pb_ctx_create(pb_ctx **out, int device, const uint8_t circuit[PB_CIRCUIT_BYTES],
                  const uint8_t *srs_g1s /* [srs_len][3] */, uint32_t srs_len, const uint8_t srs_g2[4]);
int pb_ctx_destroy(pb_ctx *ctx);
/* what plonk_new computes: h[4] k1_h[4] k2_h[4] h_pows_inv[16] z_h_x[8] z_h_x.len -> out[37] (plonk.h:43-51) */
int pb_ctx_setup_dump(const pb_ctx *ctx, uint8_t out[37]);
/* sigma_1..3[4] then S_sigma1..3 / QL QR QO QM QC / L1 coefficient rows [4] -> out[12 + 9*4] */
int pb_ctx_circuit_dump(const pb_ctx *ctx, uint8_t out[48]);
/* the eight preprocessed commitments [q_M][q_L][q_R][q_O][q_C][S1][S2][S3] and srs.g1s[0]: out[9][3] */
int pb_ctx_verifier_key(const pb_ctx *ctx, uint8_t out[27]);
/* the fixed-base table T[i][c] = g1_mul(srs.g1s[i], c), c in [0,17): out[srs_len][17][3] */
int pb_ctx_srs_table(const pb_ctx *ctx, uint8_t *out);

/* ---- kernel family (1): hf.h / gf.h element-wise.  field = 17 or 101.  out[i] = a[i] op b[i];
 * NEG and INV ignore b (may be NULL); POW uses b[i] as the exponent.  hf.h:79-203, gf.h:87-172. */
int pb_field_op_dev(int field, int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n, void *stream);
int pb_field_op(int field, int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n);

/* hf_new / gf_new (hf.h:25-35, gf.h:24-34): out[i] = v[i] mod p as the C remainder with negatives folded into [0, p) */
int pb_field_new_dev(int field, const int64_t *v, uint8_t *out, size_t n, void *stream);
int pb_field_new(int field, const int64_t *v, uint8_t *out, size_t n);

/* ---- kernel family (2): poly.h, matrix.h interpolation.  A polynomial batch is rows of `stride`
 * coefficient bytes (low degree first) plus one length byte per row; rows are passed through
 * poly_new first (trailing zeros trimmed, poly.h:20-38); 1 <= len <= stride <= PB_POLY_MAX. */
/* poly_add / poly_sub / poly_mul (poly.h:72-122); so >= the untrimmed result length */
int pb_poly_binop_dev(int op, const uint8_t *a, const uint8_t *alen, size_t sa, const uint8_t *b, const uint8_t *blen,
                      size_t sb, uint8_t *out, uint8_t *olen, size_t so, size_t n, void *stream);
int pb_poly_binop(int op, const uint8_t *a, const uint8_t *alen, size_t sa, const uint8_t *b, const uint8_t *blen,
                  size_t sb, uint8_t *out, uint8_t *olen, size_t so, size_t n);
/* poly_divide (poly.h:124-177); status[i] = 1 where the reference exits on a zero divisor */
int pb_poly_divide_dev(const uint8_t *num, const uint8_t *nlen, size_t sn, const uint8_t *den, const uint8_t *dlen, size_t sd,
                       uint8_t *quot, uint8_t *qlen, size_t sq, uint8_t *rem, uint8_t *rlen, size_t sr,
                       uint8_t *status, size_t n, void *stream);
int pb_poly_divide(const uint8_t *num, const uint8_t *nlen, size_t sn, const uint8_t *den, const uint8_t *dlen, size_t sd,
                   uint8_t *quot, uint8_t *qlen, size_t sq, uint8_t *rem, uint8_t *rlen, size_t sr,
                   uint8_t *status, size_t n);
/* poly_divide by the context's Z_H = x^4 - 1 (the prover's own call, plonk.h:505: poly_divide(&t_numer, &plonk->z_h_x, ...)):
 * numerator stride sn = 11 or 22; quot[n][sn - 4], rem[n][4]; byte-identical to pb_poly_divide with that divisor */
int pb_poly_divide_zh_dev(const pb_ctx *ctx, const uint8_t *num, const uint8_t *nlen, size_t sn, uint8_t *quot, uint8_t *qlen,
                          uint8_t *rem, uint8_t *rlen, uint8_t *status, size_t n, void *stream);
int pb_poly_divide_zh(const pb_ctx *ctx, const uint8_t *num, const uint8_t *nlen, size_t sn, uint8_t *quot, uint8_t *qlen,
                      uint8_t *rem, uint8_t *rlen, uint8_t *status, size_t n);
/* poly_eval (poly.h:265-272) */
int pb_poly_eval_dev(const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *x, uint8_t *out, size_t n, void *stream);
int pb_poly_eval(const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *x, uint8_t *out, size_t n);
/* poly_scale / poly_negate / poly_shift / poly_add_hf (poly.h:179-216, 240-254, 67-70); k[i] is the scalar or shift */
int pb_poly_unop_dev(int op, const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *k, uint8_t *out,
                     uint8_t *olen, size_t so, size_t n, void *stream);
int pb_poly_unop(int op, const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *k, uint8_t *out,
                 uint8_t *olen, size_t so, size_t n);
/* poly_slice (poly.h:218-238); status[i] = 1 where the reference exits on bad indices */
int pb_poly_slice_dev(const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *start, const uint8_t *end,
                      uint8_t *out, uint8_t *olen, size_t so, uint8_t *status, size_t n, void *stream);
int pb_poly_slice(const uint8_t *p, const uint8_t *plen, size_t sp, const uint8_t *start, const uint8_t *end,
                  uint8_t *out, uint8_t *olen, size_t so, uint8_t *status, size_t n);
/* poly_lagrange (poly.h:288-321) over `len` points per item; status[i] = 1 on duplicate x */
int pb_poly_lagrange_dev(const uint8_t *xs, const uint8_t *ys, size_t len, uint8_t *out, uint8_t *olen, size_t so,
                         uint8_t *status, size_t n, void *stream);
int pb_poly_lagrange(const uint8_t *xs, const uint8_t *ys, size_t len, uint8_t *out, uint8_t *olen, size_t so,
                     uint8_t *status, size_t n);
/* interpolate_at_h (plonk.h:162-195): h_pows_inv (4x4) times vals[i][4]; out[i][4] + trimmed length */
int pb_interpolate_at_h_dev(const pb_ctx *ctx, const uint8_t *vals, uint8_t *out, uint8_t *olen, size_t n, void *stream);
int pb_interpolate_at_h(const pb_ctx *ctx, const uint8_t *vals, uint8_t *out, uint8_t *olen, size_t n);
/* BASELINE config 2 in one launch (SURVEY.md section 8(d) "config 2 unit"): per item a[6], b[6], x, vals[4] ->
 * prod = poly_mul(a, b) [11]+len, (quot [7]+len, rem [4]+len) = poly_divide(prod, Z_H of the context), evals = poly_eval(a, x),
 * interp [4]+len = interpolate_at_h(vals).  Byte-identical to the four separate entry points. */
int pb_config2_items_dev(const pb_ctx *ctx, const uint8_t *a, const uint8_t *b, const uint8_t *x, const uint8_t *vals,
                         uint8_t *prod, uint8_t *prod_len, uint8_t *quot, uint8_t *quot_len, uint8_t *rem, uint8_t *rem_len,
                         uint8_t *evals, uint8_t *interp, uint8_t *interp_len, size_t n, void *stream);
int pb_config2_items(const pb_ctx *ctx, const uint8_t *a, const uint8_t *b, const uint8_t *x, const uint8_t *vals,
                     uint8_t *prod, uint8_t *prod_len, uint8_t *quot, uint8_t *quot_len, uint8_t *rem, uint8_t *rem_len,
                     uint8_t *evals, uint8_t *interp, uint8_t *interp_len, size_t n);
/* matrix_mul (matrix.h:81-98): out[i] = a[i] (m x k) times b[i] (k x c), row-major, m,k,c <= 8 */
int pb_matrix_mul_dev(const uint8_t *a, const uint8_t *b, uint8_t *out, uint32_t m, uint32_t k, uint32_t c, size_t n, void *stream);
int pb_matrix_mul(const uint8_t *a, const uint8_t *b, uint8_t *out, uint32_t m, uint32_t k, uint32_t c, size_t n);
/* matrix_inv (matrix.h:151-176, Gauss-Jordan on (M | I), no singularity detection), dim <= 8 */
int pb_matrix_inv_dev(const uint8_t *a, uint8_t *out, uint32_t dim, size_t n, void *stream);
int pb_matrix_inv(const uint8_t *a, uint8_t *out, uint32_t dim, size_t n);

/* matrix_gauss_jordan (matrix.h:100-149): in-place reduced row echelon form, rows <= 8, cols <= 16 */
int pb_matrix_gauss_jordan_dev(uint8_t *a, uint32_t rows, uint32_t cols, size_t n, void *stream);
int pb_matrix_gauss_jordan(uint8_t *a, uint32_t rows, uint32_t cols, size_t n);

/* ---- kernel family (3): g1.h, g2.h, srs.h */
/* g1_add / g1_double / g1_neg (g1.h:37-89); b is ignored for DOUBLE and NEG */
int pb_g1_op_dev(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n, void *stream);
int pb_g1_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n);
/* g1_mul (g1.h:91-103): raw 64-bit scalars, LSB-first double-and-add */
int pb_g1_mul_dev(const uint8_t *points, const uint64_t *scalars, uint8_t *out, size_t n, void *stream);
int pb_g1_mul(const uint8_t *points, const uint64_t *scalars, uint8_t *out, size_t n);
/* same with one-byte scalars (the KZG case: scalars are F17 coefficients) */
int pb_g1_mul_u8_dev(const uint8_t *points, const uint8_t *scalars, uint8_t *out, size_t n, void *stream);
int pb_g1_mul_u8(const uint8_t *points, const uint8_t *scalars, uint8_t *out, size_t n);
/* g1_is_on_curve (g1.h:26-31): out[i] in {0,1} */
int pb_g1_is_on_curve_dev(const uint8_t *points, uint8_t *out, size_t n, void *stream);
int pb_g1_is_on_curve(const uint8_t *points, uint8_t *out, size_t n);
/* g2_add / g2_neg (g2.h:27-66) */
int pb_g2_op_dev(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n, void *stream);
int pb_g2_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n);
/* g2_mul (g2.h:68-84); scalar 0 is undefined in the reference and yields (0xFF,0xFF) here */
int pb_g2_mul_dev(const uint8_t *points, const uint64_t *scalars, uint8_t *out, size_t n, void *stream);
int pb_g2_mul(const uint8_t *points, const uint64_t *scalars, uint8_t *out, size_t n);
/* srs_eval_at_s (srs.h:53-68), the KZG commitment, against the context's SRS;
 * status[i] = 1 where the reference exits because the polynomial is longer than the SRS */
int pb_srs_eval_at_s_dev(const pb_ctx *ctx, const uint8_t *polys, const uint8_t *plen, size_t sp, uint8_t *out,
                         uint8_t *status, size_t n, void *stream);
int pb_srs_eval_at_s(const pb_ctx *ctx, const uint8_t *polys, const uint8_t *plen, size_t sp, uint8_t *out,
                     uint8_t *status, size_t n);

/* the same against an SRS handed over per call (no context): srs_g1s[srs_len][3]; the reference's loop as written, g1_mul
 * per term by double-and-add.  trim = 1: rows pass through poly_new first (as everywhere in this ABI); trim = 0: the loop runs
 * over plen[i] terms as given, like the reference's over POLY.len (a trailing zero term is not always a no-op, g1.h:60).
 * What the drop-in srs_eval_at_s calls: one launch per commitment instead of one per term. */
int pb_srs_eval_at_s_raw_dev(const uint8_t *srs_g1s, uint32_t srs_len, const uint8_t *polys, const uint8_t *plen, size_t sp,
                             int trim, uint8_t *out, uint8_t *status, size_t n, void *stream);
int pb_srs_eval_at_s_raw(const uint8_t *srs_g1s, uint32_t srs_len, const uint8_t *polys, const uint8_t *plen, size_t sp,
                         int trim, uint8_t *out, uint8_t *status, size_t n);

/* ---- kernel family (4): gt.h, pairing.h */
int pb_gtp_mul_dev(const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n, void *stream); /* gt.h:23-28 */
int pb_gtp_mul(const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n);
int pb_gtp_pow_dev(const uint8_t *a, const uint64_t *e, uint8_t *out, size_t n, void *stream); /* gt.h:30-51 */
int pb_gtp_pow(const uint8_t *a, const uint64_t *e, uint8_t *out, size_t n);
int pb_line_dev(const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n, void *stream); /* pairing.h:19-29, out[i][3] */
int pb_line(const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n);
/* pairing (pairing.h:66-83): out[i] = e(p[i], q[i]) in GT */
int pb_pairing_dev(const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n, void *stream);
int pb_pairing(const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n);
/* pairing_f (pairing.h:31-64), the Miller function for r >= 1 */
int pb_pairing_f_dev(uint64_t r, const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n, void *stream);
int pb_pairing_f(uint64_t r, const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n);

/* ---- protocol: plonk.h */
/* constraints_satisfy (constraints.h:145-171): out[i] in {0,1} */
int pb_constraints_satisfy_dev(const pb_ctx *ctx, const uint8_t *witness, uint8_t *out, size_t n, void *stream);
int pb_constraints_satisfy(const pb_ctx *ctx, const uint8_t *witness, uint8_t *out, size_t n);
/* constraints_satisfy for a gate list of any length (no context): selectors[5][rows] = q_l q_r q_o q_m q_c rows, a / b / c
 * [n][rows]; first_bad[i] = index of the first row whose gate equation fails (the row the reference prints), -1 if none */
int pb_constraints_satisfy_rows_dev(const uint8_t *selectors, uint32_t rows, const uint8_t *a, const uint8_t *b,
                                    const uint8_t *c, int32_t *first_bad, size_t n, void *stream);
int pb_constraints_satisfy_rows(const uint8_t *selectors, uint32_t rows, const uint8_t *a, const uint8_t *b, const uint8_t *c,
                                int32_t *first_bad, size_t n);
/* plonk_prove (plonk.h:223-656) over a batch: witness[n][12], rnd[n][9], chal[n][5] -> proofs[n][34], status[n] */
int pb_plonk_prove_dev(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, const uint8_t *chal,
                       uint8_t *proofs, uint8_t *status, size_t n, void *stream);
int pb_plonk_prove(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, const uint8_t *chal,
                   uint8_t *proofs, uint8_t *status, size_t n);
/* plonk_verify (new; specification: oracle/verify_spec.inc): proofs[n][34], chal[n][5], u[n] ->
 * verdict[n]; gt (optional, may be NULL): [n][4] = lhs.a lhs.b rhs.a rhs.b of the final pairing check.
 * Three implementations with identical results, chosen per context: the specification's own sequence of group operations
 * (any key); for a key of canonical curve points, fixed-base tables + one joint double-and-add + two Miller loops; and, by
 * default for such a key, discrete logarithms mod 102 with two 102-entry pairing tables built at context creation with the
 * first implementation's pairing (environment PB_VERIFY_TABLES=0 at pb_ctx_create keeps the second). */
int pb_plonk_verify_dev(const pb_ctx *ctx, const uint8_t *proofs, const uint8_t *chal, const uint8_t *u,
                        uint8_t *verdict, uint8_t *gt, size_t n, void *stream);
int pb_plonk_verify(const pb_ctx *ctx, const uint8_t *proofs, const uint8_t *chal, const uint8_t *u,
                    uint8_t *verdict, uint8_t *gt, size_t n);
/* verify only the items whose prove status is 0; verdict[i] = 0xFF elsewhere (device pointers only) */
int pb_plonk_verify_completed_dev(const pb_ctx *ctx, const uint8_t *proofs, const uint8_t *chal, const uint8_t *u,
                                  const uint8_t *status, uint8_t *verdict, size_t n, void *stream);
/* prove then verify every completed proof (status 0); verdict[i] = 0xFF where status[i] != 0.
 * Host-pointer version: chunked, copies overlapped with compute on internal streams. */
int pb_plonk_prove_verify_dev(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, const uint8_t *chal,
                              const uint8_t *u, uint8_t *proofs, uint8_t *status, uint8_t *verdict, size_t n, void *stream);
/* same, recording the CUDA event `mid_event` (a cudaEvent_t passed as void*, may be NULL) between the prover and the
 * verifier launch, so that a caller can time the two kernels of one call separately */
int pb_plonk_prove_verify_ex_dev(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, const uint8_t *chal,
                                 const uint8_t *u, uint8_t *proofs, uint8_t *status, uint8_t *verdict, size_t n, void *stream,
                                 void *mid_event);
int pb_plonk_prove_verify(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, const uint8_t *chal,
                          const uint8_t *u, uint8_t *proofs, uint8_t *status, uint8_t *verdict, size_t n);
/* ---- outputs that cost less PCIe than the reference's structs.  End to end the path is bound by the host link, not by
 * the kernels: 27 B in + 36 B out per proof with the struct arrays above.
 *
 * Compact output: same struct inputs; D2H carries only the proofs that exist.  proofs_dense (capacity n x 34 bytes)
 * receives the PROOF structs of the completed items (status 0) in item order, *n_done their number; status[n] and
 * verdict[n] as in pb_plonk_prove_verify.  pb_wire_scatter_proofs rebuilds the [n][34] array of the struct API (the
 * records of items on which the reference exits are zero by definition, plonk_b200.h header note). */
int pb_plonk_prove_verify_compact(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, const uint8_t *chal,
                                  const uint8_t *u, uint8_t *proofs_dense, size_t *n_done, uint8_t *status,
                                  uint8_t *verdict, size_t n);
/* the dense list on the device: proofs[n][34] + status[n] -> proofs_dense (16-byte aligned, capacity n x 34),
 * *n_done_dev; offs_scratch: (n / 128 + 4) uint32 of device scratch */
int pb_gather_completed_dev(const uint8_t *proofs, const uint8_t *status, uint8_t *proofs_dense, uint32_t *n_done_dev,
                            uint32_t *offs_scratch, size_t n, void *stream);

/* Packed wire v2 (layout: csrc/wire.cuh; Python twin: plonk.c_b200/wire.py).  Input record 16 bytes = four little-endian
 * u32, each seven base-17 digits (least significant first) of the 27 values a[4] b[4] c[4] | rand[9] | alpha beta gamma z v
 * | u, one spare digit 0.  Output: 22 bytes per COMPLETED proof, dense, item order (nine u16 points x | y << 7 |
 * infinite << 14, one u32 with the seven openings as base-17 digits), and one sv byte per item (low nibble: prove status,
 * 15 = PB_PROVE_BAD_INPUT; high nibble: verdict, 15 = not verified).  16 B in, ~14 B out per item instead of 27 / 36. */
#define PB_PACKED_IN_BYTES 16
#define PB_PACKED_PROOF_BYTES 22
int pb_plonk_prove_verify_packed(const pb_ctx *ctx, const uint8_t *packed_in, uint8_t *packed_proofs /* capacity n x 22 */,
                                 size_t *n_done, uint8_t *sv, size_t n);
size_t pb_packed_workspace_bytes(size_t n);
int pb_plonk_prove_verify_packed_dev(const pb_ctx *ctx, const uint8_t *packed_in, uint8_t *packed_proofs,
                                     uint32_t *n_done_dev, uint8_t *sv, void *workspace, size_t n, void *stream);
/* format conversion on the host (plain C loops, no arithmetic of the path) */
int pb_wire_pack_inputs(const uint8_t *witness, const uint8_t *rnd, const uint8_t *chal, const uint8_t *u, uint8_t *packed, size_t n);
int pb_wire_unpack_inputs(const uint8_t *packed, uint8_t *witness, uint8_t *rnd, uint8_t *chal, uint8_t *u,
                          uint8_t *valid /* optional */, size_t n);
int pb_wire_pack_proofs(const uint8_t *proofs, uint8_t *packed, size_t n);
int pb_wire_unpack_proofs(const uint8_t *packed, uint8_t *proofs, size_t n);
int pb_wire_scatter_proofs(const uint8_t *proofs_dense, const uint8_t *status, uint8_t *proofs, size_t n);
int pb_wire_split_sv(const uint8_t *sv, uint8_t *status /* optional */, uint8_t *verdict /* optional */, size_t n);

/* Packed wire v3 (layout: csrc/wire.cuh; Python twin: plonk.c_b200/wire.py): the same information in 14 B in and 12 B per
 * COMPLETED proof out (+ the sv byte per item).  Input record = three little-endian u32 + one little-endian u16: bits 0..28
 * of word k are values 7k..7k+6 as base-17 digits; G = alpha + 17 beta + ... + 17^5 u < 2^25 has its low 16 bits in the u16
 * and bits 16+3k..18+3k in bits 29..31 of word k.  Proof record = three little-endian u32; a commitment travels as the
 * 7-bit INDEX of its point in the list of the 102 points of E(F_101) (0 = infinity, then by (x, y)): word k = three
 * indices | (opening 2k + 17 opening 2k+1) << 21 | two bits of the seventh opening << 30.  Only for contexts whose SRS
 * points are canonical points of the curve (both benchmark SRS modes; otherwise PB_ERR_ARG: use v2) -- every
 * commitment is then on the curve.  pb_wire3_pack_proofs refuses records with a point off the curve. */
#define PB_PACKED3_IN_BYTES 14
#define PB_PACKED3_PROOF_BYTES 12
int pb_plonk_prove_verify_packed3(const pb_ctx *ctx, const uint8_t *packed_in, uint8_t *packed_proofs /* capacity n x 12 */,
                                  size_t *n_done, uint8_t *sv, size_t n);
/* workspace: pb_packed_workspace_bytes(n) */
int pb_plonk_prove_verify_packed3_dev(const pb_ctx *ctx, const uint8_t *packed_in, uint8_t *packed_proofs,
                                      uint32_t *n_done_dev, uint8_t *sv, void *workspace, size_t n, void *stream);
int pb_wire3_pack_inputs(const uint8_t *witness, const uint8_t *rnd, const uint8_t *chal, const uint8_t *u, uint8_t *packed, size_t n);
int pb_wire3_unpack_inputs(const uint8_t *packed, uint8_t *witness, uint8_t *rnd, uint8_t *chal, uint8_t *u,
                           uint8_t *valid /* optional */, size_t n);
int pb_wire3_pack_proofs(const uint8_t *proofs, uint8_t *packed, size_t n);
int pb_wire3_unpack_proofs(const uint8_t *packed, uint8_t *proofs, size_t n);

/* Seeded mode (SURVEY.md sections 7.4 / 8(e)): items [start, start + count) of the synthetic stream `seed` are generated
 * on the device (draw j of item i = splitmix64(seed + 16 i + j): witness row, nine blinding scalars, five challenges, u;
 * variant 0 = "U17", 1 = "NZ"; plonk.c_b200/workload.py make_batch is the host twin), proved, verified and tallied there;
 * only counts[18] (pb_tally_dev layout) come back.  No batch data crosses PCIe. */
int pb_plonk_prove_verify_seeded(const pb_ctx *ctx, uint64_t seed, uint64_t start, size_t count, int variant, int64_t counts[18]);
size_t pb_seeded_workspace_bytes(size_t n);
int pb_plonk_prove_verify_seeded_dev(const pb_ctx *ctx, uint64_t seed, uint64_t start, size_t n, int variant, void *workspace,
                                     int64_t *counts_dev, void *stream);
/* the generator alone: struct arrays (witness rnd chal u, all or none) and / or packed records (may be NULL) */
int pb_synth_batch_dev(const pb_ctx *ctx, uint64_t seed, uint64_t start, size_t n, int variant, uint8_t *witness,
                       uint8_t *rnd, uint8_t *chal, uint8_t *u, uint8_t *packed, void *stream);

/* ---- Fiat-Shamir mode (optional).  The reference's prover takes its challenges from
 * the caller (CHALLENGE, plonk.h:16-22,227) -- the entry points above keep that interface and are bit-exact with it.
 * Here the five challenges, and the verifier's u, are drawn from a transcript hash of the circuit, the SRS and the proof
 * elements produced so far -- a duplex sponge over a 128-bit ARX state, HalfSipHash's round function (specification:
 * oracle/fs_spec.inc; kernel side: csrc/transcript.cuh).  Given the challenges
 * the transcript yields, proofs and statuses are exactly what plonk_prove returns for them. */
/* the transcript state after absorbing this context's circuit + SRS: four 32-bit words */
int pb_ctx_fs_seed(const pb_ctx *ctx, uint32_t out[4]);
/* witness[n][12], rnd[n][9] -> proofs[n][34], status[n]; chal_out (optional, may be NULL): [n][6] = alpha beta gamma z v u
 * as drawn, 0xFF for a challenge the reference's execution exits before drawing */
int pb_plonk_prove_fs_dev(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, uint8_t *proofs, uint8_t *status,
                          uint8_t *chal_out, size_t n, void *stream);
int pb_plonk_prove_fs(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, uint8_t *proofs, uint8_t *status,
                      uint8_t *chal_out, size_t n);
/* the verifier re-derives all six challenges from the proof bytes */
int pb_plonk_verify_fs_dev(const pb_ctx *ctx, const uint8_t *proofs, uint8_t *verdict, uint8_t *gt, size_t n, void *stream);
int pb_plonk_verify_fs(const pb_ctx *ctx, const uint8_t *proofs, uint8_t *verdict, uint8_t *gt, size_t n);
/* prove, then verify every completed proof (verdict 0xFF elsewhere); mid_event as in pb_plonk_prove_verify_ex_dev */
int pb_plonk_prove_verify_fs_dev(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, uint8_t *proofs, uint8_t *status,
                                 uint8_t *verdict, size_t n, void *stream, void *mid_event);
int pb_plonk_prove_verify_fs(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, uint8_t *proofs, uint8_t *status,
                             uint8_t *verdict, size_t n);
/* the six challenges of each PROOF record as a verifier derives them: chal6[n][6] */
int pb_fs_challenges_dev(const pb_ctx *ctx, const uint8_t *proofs, uint8_t *chal6, size_t n, void *stream);
int pb_fs_challenges(const pb_ctx *ctx, const uint8_t *proofs, uint8_t *chal6, size_t n);
/* ---- circuit front-end (host side).  Lowers what the reference's expression compiler produces --
 * a GATE_LIST: gates[n][5] = GATE{q_l q_r q_o q_m q_c} and the variable index on each gate's a / b / c wire
 * (eval_expr, gate_list_append: constraints.h:227-309) -- to the 44-byte circuit of pb_ctx_create: selectors by gate,
 * copy constraints by wiring every position to the next one (A1..A4 B1..B4 C1..C4, cyclically) that carries the same
 * variable.  equal_pairs[n_equal][2] (optional) lists variables asserted equal, e.g. the outputs of two expressions.
 * num_gates <= 4: the domain is fixed by omega = 4 of order 4 (plonk.h:12); unused rows become all-zero gates.  For
 * the gate list mul(x,x) mul(y,y) mul(z,z) sum(xx,yy)=zz it yields exactly plonk-test.c's hand-written circuit. */
int pb_circuit_from_gates(const uint8_t *gates, const size_t *a_idx, const size_t *b_idx, const size_t *c_idx,
                          size_t num_gates, const size_t *equal_pairs, size_t n_equal, uint8_t circuit[PB_CIRCUIT_BYTES]);
/* witness[n][12] = a[4] b[4] c[4] from per-item variable values var_values[n][n_vars] and the same wire indices */
int pb_witness_from_values(const size_t *a_idx, const size_t *b_idx, const size_t *c_idx, size_t num_gates,
                           const uint8_t *var_values, size_t n_vars, uint8_t *witness, size_t n);

/* on-device tally: counts[0..15] += number of items per status byte (0..14, 15 = anything else),
 * counts[16] += verdict==1, counts[17] += a 64-bit sum of all proof bytes (checksum). counts: int64[18] device ptr */
int pb_tally_dev(const uint8_t *proofs, const uint8_t *status, const uint8_t *verdict, size_t n, int64_t *counts, void *stream);
/* prove, verify and count in one call: counts[18] += what pb_tally_dev would add for this batch's proofs / status / verdict.
 * With the table-path verifier the counters come out of the verifier's epilogue (no third pass over the batch); otherwise
 * this is pb_plonk_prove_verify_ex_dev followed by pb_tally_dev.  mid_event as in pb_plonk_prove_verify_ex_dev. */
int pb_plonk_prove_verify_tally_dev(const pb_ctx *ctx, const uint8_t *witness, const uint8_t *rnd, const uint8_t *chal,
                                    const uint8_t *u, uint8_t *proofs, uint8_t *status, uint8_t *verdict, int64_t *counts,
                                    size_t n, void *stream, void *mid_event);


/* ---- roofline denominators measured on the device the caller is on (MEASURED_PEAKS.json
 * has no integer peak).  Each runs a dependency-free instruction stream on every SM and reports how many
 * thread-level operations it issued; the caller times it.  kind 0: 32-bit IMAD (fma pipe); kind 1: LOP3
 * (alu pipe); kind 2: IMAD and LOP3 interleaved 1:1 (both pipes); kind 3: shared-memory byte look-ups (LDS.U8);
 * kind 4: FFMA; kind 5: HFMA2; kind 6: IDP4A; kind 7: IMAD.WIDE. */
int pb_peak_probe_dev(int kind, uint32_t iters, uint64_t *ops_out_host, uint32_t *sink_dev, void *stream);
#ifdef __cplusplus
}
#endif
#endif /* PLONK_B200_H */
