/* g1.h -- drop-in for plonk.c's src/g1.h: affine points of y^2 = x^3 + 3 over F101 with an explicit
 * infinity flag (3 bytes, src/g1.h:8-11). */
#ifndef G1_H
#define G1_H

#include <stdbool.h>
#include <stdint.h>
#include "gf.h"

typedef struct {
  GF x, y;
  bool infinite;
} G1;

#ifdef __cplusplus
extern "C" {
#endif
G1 g1_new(uint64_t x_val, uint64_t y_val);
G1 g1_generator(void);                 /* (1, 2) */
G1 g1_identity(void);                  /* {0, 0, infinite} */
bool g1_is_on_curve(const G1 *point);
G1 g1_double(const G1 *a);
G1 g1_add(const G1 *a, const G1 *b);
G1 g1_neg(G1 *a);
G1 g1_mul(const G1 *point, uint64_t scalar);
GF g1_generator_subgroup_size(void);   /* 17, as a GF */
#ifdef __cplusplus
}
#endif

#endif /* G1_H */
