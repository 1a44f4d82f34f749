/* plonk.h -- drop-in for plonk.c's src/plonk.h.  Same structs (PROOF is 34 bytes, src/plonk.h:24-41), same
 * functions; plonk_prove runs the fused CUDA prover (pb_plonk_prove at batch size 1) and terminates the process
 * exactly where the reference does (exit(EXIT_FAILURE) or abort()).  As in the reference
 * this header does NOT pull in gt.h / pairing.h (src/plonk.h:4-10).  The verifier, which the reference lacks, is
 * only in the batch ABI (plonk_b200.h: pb_plonk_verify). */
#ifndef PLONK_H
#define PLONK_H

#include <assert.h>
#include <stdlib.h>
#include "constraints.h"
#include "matrix.h"
#include "poly.h"
#include "srs.h"
#include "hf.h"

#define OMEGA_VALUE 4
#define K1_VALUE 2
#define K2_VALUE 3

typedef struct {
  HF alpha;
  HF beta;
  HF gamma;
  HF z;
  HF v;
} CHALLENGE;

typedef struct {
  G1 a_s;
  G1 b_s;
  G1 c_s;
  G1 z_s;
  G1 t_lo_s;
  G1 t_mid_s;
  G1 t_hi_s;
  G1 w_z_s;
  G1 w_z_omega_s;
  HF a_z;
  HF b_z;
  HF c_z;
  HF s_sigma_1_z;
  HF s_sigma_2_z;
  HF r_z;
  HF z_omega_z;
} PROOF;

typedef struct {
  SRS srs; /* owned: plonk_free releases it */
  HF *h;
  MATRIX h_pows_inv;
  size_t h_len;
  HF *k1_h;
  HF *k2_h;
  POLY z_h_x;
} PLONK;

#ifdef __cplusplus
extern "C" {
#endif
PLONK plonk_new(SRS srs, size_t n);
void plonk_free(PLONK *plonk);
void copy_constraints_to_roots(const PLONK *plonk, const COPY_OF *copy_of, size_t len, HF *sigma);
POLY interpolate_at_h(const PLONK *plonk, const HF *values, size_t len);
void poly_print(const POLY *p);
PROOF plonk_prove(PLONK *plonk, CONSTRAINTS *constraints, ASSIGNMENTS *assignments, CHALLENGE *challenge, HF rand[9]);
#ifdef __cplusplus
}
#endif

#endif /* PLONK_H */
