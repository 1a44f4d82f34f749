/* pairing.h -- drop-in for plonk.c's src/pairing.h. */
#ifndef PAIRING_H
#define PAIRING_H

#include <stdint.h>
#include "g1.h"
#include "g2.h"
#include "gt.h"

typedef struct {
  GF x;
  GF y;
  GF c;
} LINE_EQ;

#ifdef __cplusplus
extern "C" {
#endif
int gtp_equal(const GTP *p, const GTP *q);
LINE_EQ line(const G1 *a, const G1 *b);
GTP pairing_f(uint64_t r, const G1 *p, const G2 *q);
GTP pairing(const G1 *g1, const G2 *g2);
#ifdef __cplusplus
}
#endif

#endif /* PAIRING_H */
