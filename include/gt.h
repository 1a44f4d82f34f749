/* gt.h -- drop-in for plonk.c's src/gt.h: F101[u]/(u^2 + 2), element a + b u (src/gt.h:7-9). */
#ifndef GT_H
#define GT_H

#include <stdint.h>
#include "gf.h"

typedef struct {
  GF a, b;
} GTP;

#ifdef __cplusplus
extern "C" {
#endif
GTP gtp_new(GF a, GF b);
GTP gtp_neg(GTP *p);                   /* conjugation (a, -b) */
GTP gtp_mul(GTP *base, GTP *rhs);
GTP gtp_pow(GTP *base, uint64_t exp);
#ifdef __cplusplus
}
#endif

#endif /* GT_H */
