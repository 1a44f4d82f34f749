/* gf.h -- drop-in for plonk.c's src/gf.h: the base field F101 ("GF").  Include guard FE_H and the hf.h
 * include are part of the contract (src/gf.h:1-6; poly-test.c relies on both). */
#ifndef FE_H
#define FE_H

#include <stdbool.h>
#include <stdint.h>
#include "hf.h"

#define MODULO_GF 101

PB_FIELD_DEFINE(GF, gf, MODULO_GF)

static inline GF f101(int64_t n) { return gf_new(n); }
static inline bool is_odd(uint64_t n) { return (n & 1u) != 0; }

/* square-and-multiply (src/gf.h:140-151) */
static inline GF gf_pow(GF field, uint64_t exp) {
  GF acc = gf_one();
  for (; exp; exp >>= 1, field = gf_mul(field, field))
    if (is_odd(exp)) acc = gf_mul(acc, field);
  return acc;
}
/* Fermat inverse x^(p-2), hence 1/0 = 0 (src/gf.h:159-162, pinned by gf-test.c:14) */
static inline GF gf_inv(GF x) { return gf_pow(x, MODULO_GF - 2); }
static inline GF gf_div(GF a, GF b) { return gf_mul(a, gf_inv(b)); }
static inline GF gf_from_hf(HF e) { return gf_new(e.value); }

#endif /* FE_H */
