/* matrix.h -- drop-in for plonk.c's src/matrix.h (row-major MATRIX over F17, src/matrix.h:9-13,48). */
#ifndef MATRIX_H
#define MATRIX_H

#include <stdio.h>
#include <stdlib.h>
#include "hf.h"

typedef struct {
  size_t m; /* rows */
  size_t n; /* columns */
  HF *v;    /* v[col + row * n], libc malloc, released by matrix_free */
} MATRIX;

#ifdef __cplusplus
extern "C" {
#endif
MATRIX matrix_zero(size_t m, size_t n);
MATRIX matrix_new(HF *v, size_t m, size_t n);
HF matrix_get(const MATRIX *matrix, size_t row, size_t col);   /* exits when out of bounds */
void matrix_set(MATRIX *matrix, size_t row, size_t col, HF value);
void matrix_free(MATRIX *matrix);
MATRIX matrix_add(const MATRIX *a, const MATRIX *b);
MATRIX matrix_mul(const MATRIX *a, const MATRIX *b);
void matrix_gauss_jordan(MATRIX *matrix);                        /* in-place RREF, src/matrix.h:100-149 */
MATRIX matrix_inv(const MATRIX *matrix);
#ifdef __cplusplus
}
#endif

#endif /* MATRIX_H */
