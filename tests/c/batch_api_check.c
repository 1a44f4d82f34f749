/* tests/c/batch_api_check.c -- the batch C ABI (include/plonk_b200.h) used from plain C, as INTEGRATION.md section 2
 * shows: context creation, a small batch through pb_plonk_prove_verify with host pointers, and the per-family entry
 * points, checked against the reference's known answers.  Needs a GPU. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "plonk_b200.h"

#define CHECK(cond) do { if (!(cond)) { printf("FAIL %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, pb_last_error()); return 1; } } while (0)

int main(void) {
  /* the plonk-test circuit (plonk-test.c:157-213) */
  const uint8_t circuit[PB_CIRCUIT_BYTES] = {0, 0, 0, 1,  0, 0, 0, 1,  16, 16, 16, 16,  1, 1, 1, 0,  0, 0, 0, 0,
                                             1, 1, 1, 2, 1, 2, 3, 1,   0, 0, 0, 2, 1, 2, 3, 2,   0, 1, 2, 2, 4, 4, 4, 3};
  /* generator SRS: g1s[i] = g1_mul(G, 2^i), built with the library's own batch g1_mul */
  enum { SRS_LEN = 7, N = 1000 };
  uint8_t gen[SRS_LEN * 3], g1s[SRS_LEN * 3];
  uint64_t sc[SRS_LEN];
  for (int i = 0; i < SRS_LEN; i++) { gen[3 * i] = 1; gen[3 * i + 1] = 2; gen[3 * i + 2] = 0; sc[i] = 1ull << i; }
  CHECK(pb_device_count() >= 1);
  CHECK(pb_g1_mul(gen, sc, g1s, SRS_LEN) == PB_OK);
  CHECK(g1s[3] == 68 && g1s[4] == 74 && g1s[18] == 65 && g1s[19] == 3);          /* 2G and 64G */
  const uint8_t g2[4] = {36, 31, 90, 82};
  pb_ctx *ctx = NULL;
  CHECK(pb_ctx_create(&ctx, 0, circuit, g1s, SRS_LEN, g2) == PB_OK);

  /* N copies of the shipped test vector (plonk-test.c:229-267), one of them with a broken witness */
  uint8_t *wit = malloc(N * 12), *rnd = malloc(N * 9), *chal = malloc(N * 5), *u = malloc(N);
  uint8_t *proofs = malloc(N * PB_PROOF_BYTES), *status = malloc(N), *verdict = malloc(N);
  const uint8_t w0[12] = {3, 4, 5, 9, 3, 4, 5, 16, 9, 16, 8, 8}, r0[9] = {7, 4, 11, 12, 16, 2, 14, 11, 7}, c0[5] = {15, 12, 13, 5, 12};
  for (int i = 0; i < N; i++) { memcpy(wit + 12 * i, w0, 12); memcpy(rnd + 9 * i, r0, 9); memcpy(chal + 5 * i, c0, 5); u[i] = 4; }
  wit[12 * 500 + 8] = 10;                                                          /* 3*3 != 10 */
  CHECK(pb_plonk_prove_verify(ctx, wit, rnd, chal, u, proofs, status, verdict, N) == PB_OK);
  const uint8_t want[PB_PROOF_BYTES] = {91, 66, 0, 26, 45, 0, 91, 35, 0, 32, 59, 0, 12, 32, 0, 26, 45, 0, 91, 66, 0, 91, 35, 0, 65, 98, 0,
                                        15, 13, 5, 1, 12, 15, 15};                 /* SURVEY.md Appendix A.2 */
  for (int i = 0; i < N; i++) {
    if (i == 500) { CHECK(status[i] == PB_PROVE_UNSATISFIED && verdict[i] == 0xFF && proofs[34 * i] == 0); continue; }
    CHECK(status[i] == PB_PROVE_OK && verdict[i] == PB_VERIFY_ACCEPT && memcmp(proofs + 34 * i, want, 34) == 0);
  }
  /* tamper with a_z: the verifier must reject, both GT values as in SURVEY.md Appendix A.3 */
  uint8_t gt[4], v1;
  proofs[27] = 16;
  CHECK(pb_plonk_verify(ctx, proofs, chal, u, &v1, gt, 1) == PB_OK);
  CHECK(v1 == PB_VERIFY_REJECT && gt[0] == 93 && gt[1] == 76 && gt[2] == 59 && gt[3] == 52);
  /* families: hf_div with 1/0 = 0, poly_mul (poly-test.c:101-115), pairing e(G, H) = (7, 28) */
  uint8_t a[16] = {5, 3}, b[16] = {0, 6}, q[16];
  CHECK(pb_field_op(17, PB_OP_DIV, a, b, q, 16) == PB_OK && q[0] == 0 && q[1] == 9);
  uint8_t pa[4] = {5, 0, 10, 6}, pb_[3] = {1, 2, 4}, la = 4, lb = 3, out[6], lo;
  CHECK(pb_poly_binop(PB_POLY_MUL, pa, &la, 4, pb_, &lb, 3, out, &lo, 6, 1) == PB_OK);
  CHECK(lo == 6 && out[0] == 5 && out[2] == 13 && out[5] == 7);
  uint8_t G[3] = {1, 2, 0}, H[2] = {36, 31}, e[2];
  CHECK(pb_pairing(G, H, e, 1) == PB_OK && e[0] == 7 && e[1] == 28);
  CHECK(pb_ctx_destroy(ctx) == PB_OK);
  printf("batch_api_check: all checks passed\n");
  return 0;
}
