/* poly.h -- drop-in for plonk.c's src/poly.h.  POLY is the reference's struct (src/poly.h:10-13): coefficients
 * low degree first in a libc-malloc'd block the caller releases with poly_free.  Constructors (trim trailing
 * zeros, src/poly.h:20-38) are inline; the arithmetic is executed by libplonk_b200.so on the GPU (batch size 1
 * of the pb_poly_* entry points of plonk_b200.h) and exits like the reference where the reference exits. */
#ifndef POLY_H
#define POLY_H

#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <stdbool.h>
#include "hf.h"

typedef struct {
  HF *coeffs;
  size_t len;
} POLY;

static inline POLY poly_new(const HF *coeffs, size_t len) {
  POLY p;
  while (len > 1 && coeffs[len - 1].value == 0) --len;
  p.len = len;
  p.coeffs = (HF *)malloc(len * sizeof(HF));
  if (!p.coeffs) { fputs("Memory allocation failed in poly_new\n", stderr); exit(EXIT_FAILURE); }
  for (size_t i = 0; i < len; ++i) p.coeffs[i] = coeffs[i];
  return p;
}
static inline POLY poly_zero(void) { HF z = hf_zero(); return poly_new(&z, 1); }
static inline POLY poly_one(void) { HF o = hf_one(); return poly_new(&o, 1); }
static inline bool poly_is_zero(const POLY *p) {
  for (size_t i = 0; i < p->len; ++i)
    if (p->coeffs[i].value) return false;
  return true;
}
static inline void poly_free(POLY *p) {
  free(p->coeffs);
  p->coeffs = NULL;
  p->len = 0;
}

#ifdef __cplusplus
extern "C" {
#endif
POLY poly_add_hf(POLY *a, const HF b);                 /* in place, returns an alias (src/poly.h:67-70) */
POLY poly_add(const POLY *a, const POLY *b);
POLY poly_sub(const POLY *a, const POLY *b);
POLY poly_mul(const POLY *a, const POLY *b);
void poly_divide(const POLY *num, const POLY *den, POLY *quot, POLY *rem);
POLY poly_scale(const POLY *p, HF scalar);
POLY poly_shift(const POLY *p, size_t shift);
POLY poly_slice(const POLY *p, size_t start, size_t end);
POLY poly_negate(const POLY *p);
HF poly_eval(const POLY *p, HF x);
POLY poly_z(const HF *points, size_t len);
POLY poly_lagrange(const HF *x_points, const HF *y_points, size_t len);
#ifdef __cplusplus
}
#endif

#endif /* POLY_H */
