/* srs.h -- drop-in for plonk.c's src/srs.h (KZG structured reference string, src/srs.h:11-16). */
#ifndef SRS_H
#define SRS_H

#include <stdio.h>
#include <stdlib.h>
#include "g1.h"
#include "g2.h"
#include "poly.h"

typedef struct {
  G1 *g1s;    /* libc malloc, released by srs_free */
  size_t len; /* n + 1 */
  G2 g2_1;
  G2 g2_s;
} SRS;

#ifdef __cplusplus
extern "C" {
#endif
SRS srs_create(GF secret, size_t n);   /* reproduces the reference's degenerate SRS: multiples of the identity (src/srs.h:27-35) */
void srs_free(SRS *srs);
G1 srs_eval_at_s(const SRS *srs, const POLY *vs);
#ifdef __cplusplus
}
#endif

#endif /* SRS_H */
