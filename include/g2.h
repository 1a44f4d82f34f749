/* g2.h -- drop-in for plonk.c's src/g2.h: points (x, y u) with u^2 = -2, two bytes, no identity (src/g2.h:7-9). */
#ifndef G2_H
#define G2_H

#include <stdint.h>
#include "gf.h"

typedef struct {
  GF x, y;
} G2;

#ifdef __cplusplus
extern "C" {
#endif
G2 g2_new(uint64_t x, uint64_t y);
G2 g2_generator(void);                 /* (36, 31) */
uint64_t g2_embedding_degree(void);    /* 2 */
G2 g2_neg(G2 *p);
G2 g2_add(const G2 *p, const G2 *q);
G2 g2_mul(G2 base, uint64_t scalar);   /* by value, as in the reference; scalar 0 is undefined there */
#ifdef __cplusplus
}
#endif

#endif /* G2_H */
