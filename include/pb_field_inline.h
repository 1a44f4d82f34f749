/*
 * pb_field_inline.h -- generator for the scalar field types of the drop-in headers (hf.h, gf.h).
 *
 * One macro stamps out the element type and its inline operations for a prime p < 128.  Results are the
 * canonical residues the reference computes with `%` (src/hf.h:79-137, src/gf.h:87-151).  These one-byte
 * operations are the only arithmetic of the drop-in surface that runs on the host: they are the element
 * accessors of the data model (the reference declares them `static inline` for the same reason).  Everything
 * from polynomials upwards is executed by the CUDA library.
 */
#ifndef PB_FIELD_INLINE_H
#define PB_FIELD_INLINE_H

#include <stdbool.h>
#include <stdint.h>

#define PB_FIELD_DEFINE(T, pfx, P)                                                                      \
  typedef struct { uint8_t value; } T;                                                                   \
  static inline T pfx##_new(int64_t v) { int64_t r = v % (P); T e = { (uint8_t)(r < 0 ? r + (P) : r) }; return e; } \
  static inline T pfx##_zero(void) { T e = { 0 }; return e; }                                            \
  static inline T pfx##_one(void) { T e = { 1 }; return e; }                                             \
  static inline bool pfx##_equal(T a, T b) { return a.value == b.value; }                                \
  static inline T pfx##_add(T a, T b) { unsigned s = (unsigned)a.value + b.value; T e = { (uint8_t)(s >= (P) ? s - (P) : s) }; return e; } \
  static inline T pfx##_sub(T a, T b) { int d = (int)a.value - (int)b.value; T e = { (uint8_t)(d < 0 ? d + (P) : d) }; return e; } \
  static inline T pfx##_mul(T a, T b) { T e = { (uint8_t)(((unsigned)a.value * b.value) % (P)) }; return e; } \
  static inline T pfx##_neg(T a) { T e = { (uint8_t)(a.value ? (P) - a.value : 0) }; return e; }

#endif
