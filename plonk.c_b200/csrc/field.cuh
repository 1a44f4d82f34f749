// field.cuh -- F17 ("HF", src/hf.h) and F101 ("GF", src/gf.h) arithmetic for sm_100a.
//
// Elements live in 32-bit registers.  Sums and products are kept UNREDUCED ("raw") as long as the
// bound fits, and reduced once with a multiply-high Barrett step (IMAD.HI + IMAD): for p = 17 the
// magic is ceil(2^32/17) and the quotient is exact for x < 2^28; for p = 101 it is ceil(2^32/101),
// exact for x < 2^26.  Results are always the canonical residue, i.e. exactly what the reference's
// `%` produces (hf.h:105-109, gf.h:115-120).  Inverses are table look-ups in shared memory: the
// F17 table is the reference's own (hf.h:145-180, inv(0)=0); the F101 table holds x^99, which is
// what gf_inv computes (gf.h:159-162), so inv(0)=0 there too (256 entries: the fast paths index it with
// unreduced differences, curve.cuh).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#ifdef PB_CHECK_BOUNDS
#include <cstdio>
#include <cstdlib>
#endif

namespace pb {

#ifdef __CUDACC__
#define PB_HD __host__ __device__ __forceinline__
#define PB_D __device__ __forceinline__
#else
#define PB_HD inline
#define PB_D inline
#endif

constexpr uint32_t P17 = 17u;
constexpr uint32_t P101 = 101u;
constexpr uint32_t M17 = 252645136u;  // ceil(2^32 / 17)
constexpr uint32_t M101 = 42524429u;  // ceil(2^32 / 101)

// Range checks of the reductions' inputs, compiled into the HOST build of the test library only (tests/hostcheck is built
// with -DPB_CHECK_BOUNDS): every differential test on the CPU is then also a test of the bounds stated in the comments.
#if defined(PB_CHECK_BOUNDS) && !defined(__CUDA_ARCH__)
#define PB_BOUND(x, lim, what) \
  do { if ((uint64_t)(x) >= (uint64_t)(lim)) { fprintf(stderr, "plonk_b200: bound violated: %s(%llu)\n", what, (unsigned long long)(x)); abort(); } } while (0)
#else
#define PB_BOUND(x, lim, what) do { } while (0)
#endif

PB_HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// canonical residue of a raw value x < 2^28
PB_HD uint32_t red17(uint32_t x) { PB_BOUND(x, 1u << 28, "red17"); return x - P17 * mulhi_u32(x, M17); }
// canonical residue of a raw value x < 2^26
PB_HD uint32_t red101(uint32_t x) { PB_BOUND(x, 1u << 26, "red101"); return x - P101 * mulhi_u32(x, M101); }

// inputs canonical; outputs canonical
PB_HD uint32_t add17(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= P17 ? s - P17 : s; }
PB_HD uint32_t sub17(uint32_t a, uint32_t b) { uint32_t s = a + P17 - b; return s >= P17 ? s - P17 : s; }
PB_HD uint32_t neg17(uint32_t a) { return a ? P17 - a : 0u; }
PB_HD uint32_t mul17(uint32_t a, uint32_t b) { return red17(a * b); }
PB_HD uint32_t add101(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= P101 ? s - P101 : s; }
PB_HD uint32_t sub101(uint32_t a, uint32_t b) { uint32_t s = a + P101 - b; return s >= P101 ? s - P101 : s; }
PB_HD uint32_t neg101(uint32_t a) { return a ? P101 - a : 0u; }
PB_HD uint32_t mul101(uint32_t a, uint32_t b) { return red101(a * b); }

// square-and-multiply, LSB first, as hf_pow / gf_pow do (hf.h:127-137, gf.h:140-151)
PB_HD uint32_t pow17(uint32_t base, uint64_t e) {
  uint32_t r = 1u;
  while (e) { if (e & 1u) r = mul17(r, base); base = mul17(base, base); e >>= 1; }
  return r;
}
PB_HD uint32_t pow101(uint32_t base, uint64_t e) {
  uint32_t r = 1u;
  while (e) { if (e & 1u) r = mul101(r, base); e >>= 1; base = mul101(base, base); }
  return r;
}

// hf_new / gf_new: C remainder of a signed 64-bit value, negatives folded up (hf.h:25-35, gf.h:24-34)
PB_HD uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
PB_HD uint32_t new17(int64_t v) { int64_t t = v % 17; if (t < 0) t += 17; return (uint32_t)t; }
PB_HD uint32_t new101(int64_t v) { int64_t t = v % 101; if (t < 0) t += 101; return (uint32_t)t; }

// Shared-memory look-up tables, copied once per block from a compile-time image (kernels.cuh: build_field_tables).
struct alignas(16) FieldTables {
  uint8_t inv17[32];    // hf_inverses, hf.h:145-180
  uint8_t inv101[256];  // (x mod 101)^99 mod 101, gf.h:159-162; entries 101..255 serve unreduced indices (curve.cuh: g1_add_c)
};

PB_HD uint32_t inv17(const FieldTables& t, uint32_t a) { PB_BOUND(a, 32u, "inv17 index"); return t.inv17[a]; }
PB_HD uint32_t inv101(const FieldTables& t, uint32_t a) { PB_BOUND(a, 256u, "inv101 index"); return t.inv101[a]; }

}  // namespace pb
