// poly_fast.cuh -- shape-specialised polynomial kernels (BASELINE config 2 and the prover's shapes).
//
// The generic kernels in kernels.cuh keep polynomials in per-thread local arrays with dynamic lengths; they are
// correct for every shape but LSU-bound.  For the fixed strides that matter, the kernels below hold every coefficient
// in a register (all loops unrolled over the STRIDE, lengths applied as masks).  These kernels move ~26-45 bytes per
// item, so they are only HBM-bound if the whole item costs ~150-250 issue slots; hence
//   - INPUT records are read straight from global memory with aligned 32-bit loads + funnel shifts (a warp's records
//     are one contiguous span, so every sector is fetched once; no shared memory, no barrier on the way in);
//   - OUTPUT records (odd strides: 11, 7, 4 bytes) are assembled in shared memory and leave the block as 128-bit
//     stores of its contiguous slice, so HBM sees full lines instead of one partial sector per byte;
//   - the one-byte-per-item arrays (lengths, x, status, results) are accessed directly (already coalesced).
// Semantics are the generic kernels' (= the reference's poly_new trimming + operation + trimming), bit for bit.
#pragma once
#include "kernels.cuh"

namespace pb {

constexpr int PF_BLOCK = 256;

// The ITEM bytes of record t of a byte array (16-byte aligned base, n records), as little-endian words w[0..].
// Words past the end of the array are never dereferenced (index clamped), they only feed bytes that are masked off.
template <int ITEM>
PB_D void load_record(uint32_t (&w)[(ITEM + 3) / 4], const uint8_t* __restrict__ base, size_t t, size_t n) {
  constexpr int NW = (ITEM + 3) / 4;
  const size_t off = t * ITEM;
  const size_t last = (n * ITEM - 1) >> 2;
  const uint32_t* g = reinterpret_cast<const uint32_t*>(base);
  const size_t i0 = off >> 2;
  const uint32_t sh = (uint32_t)(off & 3u) * 8u;
  uint32_t raw[NW + 1];
#pragma unroll
  for (int k = 0; k <= NW; k++) { const size_t i = i0 + k; raw[k] = g[i < last ? i : last]; }
#pragma unroll
  for (int k = 0; k < NW; k++) w[k] = __funnelshift_r(raw[k], raw[k + 1], sh);
}
// The first `len` coefficient bytes of an N-byte record: the words are masked by length (bytes at positions >= len become 0,
// as zero padding), tested for bytes above 16 -- b > 16  <=>  bit 7 of ((b & 0x7F) + 0x6F) | b -- and unpacked.  Returns
// false when the row is not a polynomial over F17 of at most N coefficients (same rule as the generic kernels' lp_load).
template <int N>
PB_D bool unpack_masked(uint32_t (&r)[N], const uint32_t (&w)[(N + 3) / 4], uint32_t len) {
  constexpr int NW = (N + 3) / 4;
  uint32_t bad = 0u, m[NW];
#pragma unroll
  for (int k = 0; k < NW; k++) {
    const int left = (int)len - 4 * k;                                      // valid bytes in this word
    const uint32_t mask = left >= 4 ? 0xFFFFFFFFu : left <= 0 ? 0u : (1u << (8 * left)) - 1u;
    m[k] = w[k] & mask;
    bad |= (((m[k] & 0x7F7F7F7Fu) + 0x6F6F6F6Fu) | m[k]) & 0x80808080u;
  }
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = (m[i >> 2] >> (8 * (i & 3))) & 0xFFu;
  return bad == 0u && len <= (uint32_t)N;
}
// the block's slice of an output array with ITEM bytes per record: shared memory -> global, 128-bit stores
template <int ITEM>
PB_D void store_slice(uint8_t* __restrict__ g, const uint8_t* smem, size_t first, size_t n) {
  constexpr int NV = PF_BLOCK * ITEM / 16;
  uint8_t* dst = g + first * ITEM;
  if (first + PF_BLOCK <= n) {
#pragma unroll
    for (int k = threadIdx.x; k < NV; k += PF_BLOCK) reinterpret_cast<uint4*>(dst)[k] = reinterpret_cast<const uint4*>(smem)[k];
  } else {
    const int cnt = (int)(n - first) * ITEM;
    for (int k = threadIdx.x; k < cnt; k += PF_BLOCK) dst[k] = smem[k];
  }
}

// a 6-byte record (2-byte aligned: the base is 16-byte aligned) as three 16-bit loads; a warp's 32 records are one
// 192-byte span, so the three requests hit the same six sectors in L1
PB_D void load_record6(uint32_t (&r)[6], const uint8_t* __restrict__ base, size_t t) {
  const uint16_t* g = reinterpret_cast<const uint16_t*>(base) + 3 * t;
#pragma unroll
  for (int k = 0; k < 3; k++) { const uint32_t w = g[k]; r[2 * k] = w & 0xFFu; r[2 * k + 1] = w >> 8; }
}

// poly_mul (poly.h:106-122) for fixed strides SA x SB -> SA+SB-1
template <int SA, int SB>
__global__ void __launch_bounds__(PF_BLOCK) poly_mul_fast_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ alen,
                                                                 const uint8_t* __restrict__ b, const uint8_t* __restrict__ blen,
                                                                 uint8_t* __restrict__ out, uint8_t* __restrict__ olen, size_t n) {
  constexpr int SO = SA + SB - 1;
  __shared__ __align__(16) uint8_t so[PF_BLOCK * SO];
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * PF_BLOCK, t = first + tid;
  if (t < n) {
    uint32_t wa[(SA + 3) / 4], wb[(SB + 3) / 4], ra[SA], rb[SB], ro[SO];
    load_record<SA>(wa, a, t, n);
    load_record<SB>(wb, b, t, n);
    const uint32_t la = alen[t], lb = blen[t];
    bool ok = unpack_masked(ra, wa, la);
    ok &= unpack_masked(rb, wb, lb);
    ok &= la != 0u && lb != 0u;
#pragma unroll
    for (int k = 0; k < SO; k++) ro[k] = 0u;
    mul_acc<SA, SB>(ro, ra, rb);            // raw < min(SA,SB) * 2^8
#pragma unroll
    for (int k = 0; k < SO; k++) { ro[k] = red17(ro[k]); so[tid * SO + k] = ok ? (uint8_t)ro[k] : 0; }
    // F17[x] has no zero divisors: trimming the product gives its canonical length (1 for a zero product);
    // an invalid row (kernels.cuh: ITEM_INVALID rule) is reported as length 0
    olen[t] = ok ? (uint8_t)canon_len(ro) : 0;
  }
  __syncthreads();
  store_slice<SO>(out, so, first, n);
}

// poly_divide (poly.h:124-177) for fixed strides: numerator SN, divisor SD, quotient SN-SD+1 columns, remainder SD-1
template <int SN, int SD>
__global__ void __launch_bounds__(PF_BLOCK) poly_divide_fast_kernel(const uint8_t* __restrict__ num, const uint8_t* __restrict__ nlen,
                                                                    const uint8_t* __restrict__ den, const uint8_t* __restrict__ dlen,
                                                                    uint8_t* __restrict__ quot, uint8_t* __restrict__ qlen,
                                                                    uint8_t* __restrict__ rem, uint8_t* __restrict__ rlen,
                                                                    uint8_t* __restrict__ status, size_t n) {
  constexpr int SQ = SN - SD + 1, SR = SD - 1;
  __shared__ FieldTables ft;
  __shared__ __align__(16) uint8_t sq[PF_BLOCK * SQ];
  __shared__ __align__(16) uint8_t sr[PF_BLOCK * SR];
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * PF_BLOCK, t = first + tid;
  build_field_tables(ft);
  __syncthreads();
  if (t < n) {
    uint32_t wn[(SN + 3) / 4], wd[(SD + 3) / 4], r[SN], d[SD];
    load_record<SN>(wn, num, t, n);
    load_record<SD>(wd, den, t, n);
    const uint32_t nlen0 = nlen[t], dlen0 = dlen[t];
    bool ok = unpack_masked(r, wn, nlen0);
    ok &= unpack_masked(d, wd, dlen0);
    const uint32_t nl = nlen0 == 0 ? 0u : canon_len(r);      // poly_new trims the inputs first
    const uint32_t dl = dlen0 == 0 ? 0u : canon_len(d);
    bool zero_den = true;
#pragma unroll
    for (int j = 0; j < SD; j++) zero_den &= d[j] == 0u;
    zero_den |= !ok;                                             // an invalid row takes the zero-output path; its status is ITEM_INVALID
    if (!ok) {
#pragma unroll
      for (int j = 0; j < SD; j++) d[j] = 0u;                    // keep the inverse-table index in range
    }
    // top-aligned divisor: dt[j] = d[dl-1-j] (dt[0] is the leading coefficient)
    uint32_t dt[SD];
#pragma unroll
    for (int j = 0; j < SD; j++) {
      uint32_t v = 0u;
#pragma unroll
      for (int s = j; s < SD; s++) v = (dl == (uint32_t)(s + 1)) ? d[s - j] : v;
      dt[j] = v;
    }
    const uint32_t lead_inv = inv17(ft, dt[0]);
    uint32_t qt[SN];                                             // qt[k] = quotient coefficient k - (dl-1)
#pragma unroll
    for (int k = SN - 1; k >= 0; k--) {
      const bool active = (uint32_t)k < nl && (uint32_t)(k + 1) >= dl && !zero_den;
      const uint32_t rk = red17(r[k]);
      const uint32_t f = active ? red17(rk * lead_inv) : 0u;
      qt[k] = f;
      const uint32_t nf = f ? P17 - f : 0u;                      // r[k-j] -= f * dt[j], kept raw (< SD * 17 * 16 + 17)
      r[k] = rk;
#pragma unroll
      for (int j = 0; j < SD; j++)
        if (k - j >= 0) r[k - j] += nf * dt[j];
    }
#pragma unroll
    for (int k = 0; k < SR; k++) r[k] = red17(r[k]);
    // quotient columns: q[i] = qt[i + dl - 1]
    uint32_t top = 0u;
    bool any = false;
#pragma unroll
    for (int k = 0; k < SN; k++) { if (qt[k]) { top = (uint32_t)k; any = true; } }
    const uint32_t ql = any ? top + 2u - dl : 1u;
#pragma unroll
    for (int i = 0; i < SQ; i++) {
      uint32_t v = 0u;
#pragma unroll
      for (int s = 0; s < SD; s++)
        if (i + s < SN) v = (dl == (uint32_t)(s + 1)) ? qt[i + s] : v;
      sq[tid * SQ + i] = zero_den ? 0 : (uint8_t)v;
    }
    // remainder: min(dl-1, nl) columns, trimmed while > 1; length 0 for a constant divisor (hazard C-4)
    uint32_t rl = dl - 1u < nl ? dl - 1u : nl;
    if (dl == 0u) rl = 0u;
#pragma unroll
    for (int k = SR - 1; k >= 1; k--) rl = (rl == (uint32_t)(k + 1) && r[k] == 0u) ? (uint32_t)k : rl;
#pragma unroll
    for (int k = 0; k < SR; k++) sr[tid * SR + k] = (!zero_den && (uint32_t)k < rl) ? (uint8_t)r[k] : 0;
    qlen[t] = zero_den ? 0 : (uint8_t)ql;
    rlen[t] = zero_den ? 0 : (uint8_t)rl;
    status[t] = !ok ? ITEM_INVALID : zero_den ? 1 : 0;            // "Division by zero polynomial", poly.h:125-128
  }
  __syncthreads();
  store_slice<SQ>(quot, sq, first, n);
  store_slice<SR>(rem, sr, first, n);
}

// Horner from the top coefficient (poly.h:265-272) with a reduction every third step only: for canonical y, x, c the running
// value stays below 17 * 16^3 + 16^3 < 2^17 between reductions, far inside red17's range; the result is the same residue
template <int N>
PB_D uint32_t horner17(const uint32_t (&c)[N], uint32_t x) {
  uint32_t y = 0u;
#pragma unroll
  for (int k = N - 1; k >= 0; k--) {
    y = y * x + c[k];
    if ((N - 1 - k) % 3 == 2 || k == 0) y = red17(y);
  }
  return y;
}

// poly_eval (poly.h:265-272) for a fixed stride
template <int SP>
__global__ void __launch_bounds__(PF_BLOCK) poly_eval_fast_kernel(const uint8_t* __restrict__ p, const uint8_t* __restrict__ plen,
                                                                  const uint8_t* __restrict__ x, uint8_t* __restrict__ out, size_t n) {
  const size_t t = (size_t)blockIdx.x * PF_BLOCK + threadIdx.x;
  if (t >= n) return;
  uint32_t w[(SP + 3) / 4], c[SP];
  load_record<SP>(w, p, t, n);
  const uint32_t len = plen[t], xv = x[t];
  const bool ok = unpack_masked(c, w, len) && xv <= 16u;
  const uint32_t y = horner17(c, xv & 31u);                     // masked coefficients are 0, so y stays 0 above len
  out[t] = ok ? (uint8_t)y : 0xFF;
}

// BASELINE config 2 as ONE launch.  SURVEY.md section 8(d) defines the unit as {A[6], B[6], x, vals[4]} (17 bytes in) ->
// poly_mul(A, B), poly_divide(A*B, Z_H), poly_eval(A, x), interpolate_at_h(vals) (31 bytes out).  The four separate
// entry points move 26 + 36 + 9 + 9 bytes per item in four launches of 20-120 us each; fused, every input byte is read
// once, A*B never leaves registers, and the division by the context's Z_H = x^4 - 1 is seven additions.  Results are
// byte-identical to the four separate calls (tests/test_gpu_parity.py::test_config2_fused).
__global__ void __launch_bounds__(PF_BLOCK) config2_kernel(const __grid_constant__ CircuitConst cc, const uint8_t* __restrict__ a,
                                                           const uint8_t* __restrict__ b, const uint8_t* __restrict__ x,
                                                           const uint8_t* __restrict__ vals, uint8_t* __restrict__ prod,
                                                           uint8_t* __restrict__ prod_len, uint8_t* __restrict__ quot,
                                                           uint8_t* __restrict__ quot_len, uint8_t* __restrict__ rem,
                                                           uint8_t* __restrict__ rem_len, uint8_t* __restrict__ evals,
                                                           uint8_t* __restrict__ interp, uint8_t* __restrict__ interp_len, size_t n) {
  __shared__ __align__(16) uint8_t sp[PF_BLOCK * 11];
  __shared__ __align__(16) uint8_t sq[PF_BLOCK * 7];
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * PF_BLOCK, t = first + tid;
  if (t < n) {
    uint32_t ra[6], rb[6], p[11];
    load_record6(ra, a, t);
    load_record6(rb, b, t);
#pragma unroll
    for (int k = 0; k < 11; k++) p[k] = 0u;
    mul_acc<6, 6>(p, ra, rb);                                     // poly_mul, poly.h:106-122
#pragma unroll
    for (int k = 0; k < 11; k++) { p[k] = red17(p[k]); sp[tid * 11 + k] = (uint8_t)p[k]; }
    prod_len[t] = (uint8_t)canon_len(p);
    // poly_divide by Z_H = x^4 - 1 (poly.h:124-177): q[j] = p[j+4] + q[j+4], remainder p[k] + q[k]
    uint32_t q[7];
    q[6] = p[10]; q[5] = p[9]; q[4] = p[8]; q[3] = p[7];
    q[2] = add17(p[6], q[6]); q[1] = add17(p[5], q[5]); q[0] = add17(p[4], q[4]);
#pragma unroll
    for (int k = 0; k < 7; k++) sq[tid * 7 + k] = (uint8_t)q[k];
    quot_len[t] = (uint8_t)canon_len(q);
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) r[k] = add17(p[k], q[k]);
    reinterpret_cast<uint32_t*>(rem)[t] = r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24);
    rem_len[t] = (uint8_t)canon_len(r);
    // poly_eval(A, x), Horner from the top (poly.h:265-272)
    const uint32_t xv = x[t];
    evals[t] = (uint8_t)horner17(ra, xv & 31u);
    // interpolate_at_h(vals) = h_pows_inv * vals (plonk.h:162-195)
    const uint32_t wv = reinterpret_cast<const uint32_t*>(vals)[t];
    uint32_t v[4] = {wv & 0xFFu, (wv >> 8) & 0xFFu, (wv >> 16) & 0xFFu, wv >> 24}, f[4];
    interpolate(cc, v, f);
    reinterpret_cast<uint32_t*>(interp)[t] = f[0] | (f[1] << 8) | (f[2] << 16) | (f[3] << 24);
    interp_len[t] = (uint8_t)canon_len(f);
  }
  __syncthreads();
  store_slice<11>(prod, sp, first, n);
  store_slice<7>(quot, sq, first, n);
}

// ---- four consecutive items per thread -------------------------------------------------------------------------------
// One item per thread leaves these kernels issue-bound on bookkeeping: per-item address arithmetic, one narrow load per
// record, one byte store per output coefficient (config2_kernel: 329 warp-instructions per item-warp for ~110 of
// arithmetic, ncu profiles/r2).  Four consecutive items make every record group word-aligned -- 4 x 6 = 24 bytes in,
// 4 x 11 = 44 and 4 x 7 = 28 bytes out -- so a thread moves whole 32/64/128-bit words: 8 loads and 11 + 7 shared-memory
// word stores per FOUR items, output bytes inserted into their words with one PRMT each, lengths and one-byte results as
// one word per array.  Same arithmetic, same results.  Full blocks only (PF4_ITEMS items); the caller runs the ragged
// tail through the one-item-per-thread kernel.
#ifndef PB_PF4_BLOCK
#define PB_PF4_BLOCK 128
#endif
#ifndef PB_PF4_MINBLOCKS
#define PB_PF4_MINBLOCKS 1
#endif
constexpr int PF4_BLOCK = PB_PF4_BLOCK;
constexpr int PF4_ITEMS = 4 * PF4_BLOCK;
// byte I of the word array w (compile-time I): one PRMT / shift
template <int I, int NW>
PB_D uint32_t byte_of(const uint32_t (&w)[NW]) { return __byte_perm(w[I >> 2], 0u, 0x4440 | (I & 3)); }
// w's byte I := v (v < 256), compile-time I
template <int I, int NW>
PB_D void put_byte(uint32_t (&w)[NW], uint32_t v) {
  constexpr uint32_t sel = (I & 3) == 0 ? 0x3214u : (I & 3) == 1 ? 0x3240u : (I & 3) == 2 ? 0x3410u : 0x4210u;
  w[I >> 2] = __byte_perm(w[I >> 2], v, sel);
}
template <int N, int BASE, int NW>
PB_D void get_bytes(uint32_t (&r)[N], const uint32_t (&w)[NW]) {
  if constexpr (N > 0) {
    uint32_t head[N > 1 ? N - 1 : 1];
    if constexpr (N > 1) { get_bytes<N - 1, BASE>(head, w);
#pragma unroll
      for (int i = 0; i < N - 1; i++) r[i] = head[i]; }
    r[N - 1] = byte_of<BASE + N - 1>(w);
  }
}
template <int N, int BASE, int NW>
PB_D void put_bytes(uint32_t (&w)[NW], const uint32_t (&r)[N]) {
  if constexpr (N > 0) {
    if constexpr (N > 1) { uint32_t head[N - 1];
#pragma unroll
      for (int i = 0; i < N - 1; i++) head[i] = r[i];
      put_bytes<N - 1, BASE>(w, head); }
    put_byte<BASE + N - 1>(w, r[N - 1]);
  }
}
// the block's contiguous slice of an array with WORDS words per thread: shared memory (word stores at stride WORDS, an odd
// number, so conflict-free) -> global as 128-bit stores
template <int WORDS>
PB_D void store_slice4(uint8_t* __restrict__ g, const uint32_t* smem, size_t block) {
  constexpr int NV = PF4_BLOCK * WORDS / 4;
  uint4* dst = reinterpret_cast<uint4*>(g + block * (size_t)(PF4_BLOCK * WORDS * 4));
#pragma unroll
  for (int k0 = 0; k0 < NV; k0 += PF4_BLOCK) {
    const int k = k0 + (int)threadIdx.x;
    if (k0 + PF4_BLOCK <= NV || k < NV) dst[k] = reinterpret_cast<const uint4*>(smem)[k];
  }
}

// NW consecutive words of a thread (thread tg owns words [tg * NW, tg * NW + NW)) with the widest aligned vector access
template <int NW>
PB_D void load_words(uint32_t (&w)[NW], const uint8_t* __restrict__ base, size_t tg) {
  if constexpr (NW % 4 == 0) {
    const uint4* g = reinterpret_cast<const uint4*>(base) + tg * (NW / 4);
#pragma unroll
    for (int k = 0; k < NW / 4; k++) { const uint4 v = g[k]; w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
  } else if constexpr (NW % 2 == 0) {
    const uint2* g = reinterpret_cast<const uint2*>(base) + tg * (NW / 2);
#pragma unroll
    for (int k = 0; k < NW / 2; k++) { const uint2 v = g[k]; w[2 * k] = v.x; w[2 * k + 1] = v.y; }
  } else {
    const uint32_t* g = reinterpret_cast<const uint32_t*>(base) + tg * NW;
#pragma unroll
    for (int k = 0; k < NW; k++) w[k] = g[k];
  }
}
template <int NW>
PB_D void store_words(uint8_t* __restrict__ base, size_t tg, const uint32_t (&w)[NW]) {
  if constexpr (NW % 4 == 0) {
    uint4* g = reinterpret_cast<uint4*>(base) + tg * (NW / 4);
#pragma unroll
    for (int k = 0; k < NW / 4; k++) g[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
  } else if constexpr (NW % 2 == 0) {
    uint2* g = reinterpret_cast<uint2*>(base) + tg * (NW / 2);
#pragma unroll
    for (int k = 0; k < NW / 2; k++) g[k] = make_uint2(w[2 * k], w[2 * k + 1]);
  } else {
    uint32_t* g = reinterpret_cast<uint32_t*>(base) + tg * NW;
#pragma unroll
    for (int k = 0; k < NW; k++) g[k] = w[k];
  }
}
// item IT (of four) of ITEM-byte records held in `w` (ITEM words), shifted down to a word boundary: out[0..] little-endian
template <int ITEM, int IT>
PB_D void item_words(uint32_t (&out)[(ITEM + 3) / 4], const uint32_t (&w)[ITEM]) {
  constexpr int B0 = ITEM * IT, W0 = B0 / 4, SH = 8 * (B0 % 4);
#pragma unroll
  for (int k = 0; k < (ITEM + 3) / 4; k++) {
    const uint32_t lo = W0 + k < ITEM ? w[W0 + k] : 0u, hi = W0 + k + 1 < ITEM ? w[W0 + k + 1] : 0u;
    out[k] = SH == 0 ? lo : __funnelshift_r(lo, hi, SH);
  }
}

template <int IT>
PB_D void config2_item(const CircuitConst& cc, const uint32_t (&wa)[6], const uint32_t (&wb)[6], uint32_t xw, const uint32_t (&vw)[4],
                       uint32_t (&pw)[11], uint32_t (&qw)[7], uint32_t (&rw)[4], uint32_t (&fw)[4], uint32_t (&lens)[4], uint32_t& yw) {
  uint32_t ra[6], rb[6], p[11];
  get_bytes<6, 6 * IT>(ra, wa);
  get_bytes<6, 6 * IT>(rb, wb);
#pragma unroll
  for (int k = 0; k < 11; k++) p[k] = 0u;
  mul_acc<6, 6>(p, ra, rb);                                     // poly_mul, poly.h:106-122
#pragma unroll
  for (int k = 0; k < 11; k++) p[k] = red17(p[k]);
  put_bytes<11, 11 * IT>(pw, p);
  // poly_divide by Z_H = x^4 - 1 (poly.h:124-177): q[j] = p[j+4] + q[j+4], remainder p[k] + q[k]
  uint32_t q[7], r[4], f[4];
  q[6] = p[10]; q[5] = p[9]; q[4] = p[8]; q[3] = p[7];
  q[2] = add17(p[6], q[6]); q[1] = add17(p[5], q[5]); q[0] = add17(p[4], q[4]);
  put_bytes<7, 7 * IT>(qw, q);
#pragma unroll
  for (int k = 0; k < 4; k++) r[k] = add17(p[k], q[k]);
  rw[IT] = r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24);
  // poly_eval(A, x), Horner from the top (poly.h:265-272)
  const uint32_t xv = __byte_perm(xw, 0u, 0x4440 | IT);
  yw |= horner17(ra, xv & 31u) << (8 * IT);
  // interpolate_at_h(vals) = h_pows_inv * vals (plonk.h:162-195)
  const uint32_t v[4] = {vw[IT] & 0xFFu, __byte_perm(vw[IT], 0u, 0x4441), __byte_perm(vw[IT], 0u, 0x4442), vw[IT] >> 24};
  interpolate(cc, v, f);
  fw[IT] = f[0] | (f[1] << 8) | (f[2] << 16) | (f[3] << 24);
  lens[0] |= canon_len(p) << (8 * IT);
  lens[1] |= canon_len(q) << (8 * IT);
  lens[2] |= canon_len(r) << (8 * IT);
  lens[3] |= canon_len(f) << (8 * IT);
}

// BASELINE config 2, one launch, four items per thread; n_blocks * PF4_ITEMS items
__global__ void __launch_bounds__(PF4_BLOCK, PB_PF4_MINBLOCKS) config2_kernel4(const __grid_constant__ CircuitConst cc, const uint8_t* __restrict__ a,
                                                             const uint8_t* __restrict__ b, const uint8_t* __restrict__ x,
                                                             const uint8_t* __restrict__ vals, uint8_t* __restrict__ prod,
                                                             uint8_t* __restrict__ prod_len, uint8_t* __restrict__ quot,
                                                             uint8_t* __restrict__ quot_len, uint8_t* __restrict__ rem,
                                                             uint8_t* __restrict__ rem_len, uint8_t* __restrict__ evals,
                                                             uint8_t* __restrict__ interp, uint8_t* __restrict__ interp_len) {
  const size_t tg = (size_t)blockIdx.x * PF4_BLOCK + threadIdx.x;      // this thread's items: 4 tg .. 4 tg + 3
  uint32_t wa[6], wb[6], vw[4];
  {
    const uint2* A = reinterpret_cast<const uint2*>(a) + 3 * tg;
    const uint2* B = reinterpret_cast<const uint2*>(b) + 3 * tg;
#pragma unroll
    for (int k = 0; k < 3; k++) { const uint2 u = A[k], v = B[k]; wa[2 * k] = u.x; wa[2 * k + 1] = u.y; wb[2 * k] = v.x; wb[2 * k + 1] = v.y; }
    const uint4 v4 = reinterpret_cast<const uint4*>(vals)[tg];
    vw[0] = v4.x; vw[1] = v4.y; vw[2] = v4.z; vw[3] = v4.w;
  }
  const uint32_t xw = reinterpret_cast<const uint32_t*>(x)[tg];
  uint32_t pw[11], qw[7], rw[4], fw[4], lens[4] = {0u, 0u, 0u, 0u}, yw = 0u;
#pragma unroll
  for (int k = 0; k < 11; k++) pw[k] = 0u;
#pragma unroll
  for (int k = 0; k < 7; k++) qw[k] = 0u;
  config2_item<0>(cc, wa, wb, xw, vw, pw, qw, rw, fw, lens, yw);
  config2_item<1>(cc, wa, wb, xw, vw, pw, qw, rw, fw, lens, yw);
  config2_item<2>(cc, wa, wb, xw, vw, pw, qw, rw, fw, lens, yw);
  config2_item<3>(cc, wa, wb, xw, vw, pw, qw, rw, fw, lens, yw);
  reinterpret_cast<uint4*>(rem)[tg] = make_uint4(rw[0], rw[1], rw[2], rw[3]);
  reinterpret_cast<uint4*>(interp)[tg] = make_uint4(fw[0], fw[1], fw[2], fw[3]);
  reinterpret_cast<uint32_t*>(evals)[tg] = yw;
  reinterpret_cast<uint32_t*>(prod_len)[tg] = lens[0];
  reinterpret_cast<uint32_t*>(quot_len)[tg] = lens[1];
  reinterpret_cast<uint32_t*>(rem_len)[tg] = lens[2];
  reinterpret_cast<uint32_t*>(interp_len)[tg] = lens[3];
  // word stores straight to global memory: a warp's 11 (7) store instructions together cover its contiguous 1408 (896) bytes
  // (measured against staging the block's slice through shared memory for 128-bit stores: 51.4 vs 53.1 us per 2^22 items)
  store_words<11>(prod, tg, pw);
  store_words<7>(quot, tg, qw);
}

// ---- the separate entry points for the config-2 shapes, four items per thread (full PF4_ITEMS groups; same rules and
// results as the one-item kernels above, which run the ragged tail)
template <int SA, int SB, int IT>
PB_D void mul_item4(const uint32_t (&wa)[SA], const uint32_t (&wb)[SB], uint32_t la4, uint32_t lb4, uint32_t (&pw)[SA + SB - 1], uint32_t& lens) {
  constexpr int SO = SA + SB - 1;
  uint32_t ia[(SA + 3) / 4], ib[(SB + 3) / 4], ra[SA], rb[SB], ro[SO];
  item_words<SA, IT>(ia, wa);
  item_words<SB, IT>(ib, wb);
  const uint32_t la = __byte_perm(la4, 0u, 0x4440 | IT), lb = __byte_perm(lb4, 0u, 0x4440 | IT);
  bool ok = unpack_masked(ra, ia, la);
  ok &= unpack_masked(rb, ib, lb);
  ok &= la != 0u && lb != 0u;
#pragma unroll
  for (int k = 0; k < SO; k++) ro[k] = 0u;
  mul_acc<SA, SB>(ro, ra, rb);
#pragma unroll
  for (int k = 0; k < SO; k++) ro[k] = ok ? red17(ro[k]) : 0u;
  put_bytes<SO, SO * IT>(pw, ro);
  lens |= (ok ? canon_len(ro) : 0u) << (8 * IT);
}
template <int SA, int SB>
__global__ void __launch_bounds__(PF4_BLOCK) poly_mul_fast4_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ alen,
                                                                   const uint8_t* __restrict__ b, const uint8_t* __restrict__ blen,
                                                                   uint8_t* __restrict__ out, uint8_t* __restrict__ olen) {
  constexpr int SO = SA + SB - 1;
  const size_t tg = (size_t)blockIdx.x * PF4_BLOCK + threadIdx.x;
  uint32_t wa[SA], wb[SB], pw[SO], lens = 0u;
  load_words<SA>(wa, a, tg);
  load_words<SB>(wb, b, tg);
  const uint32_t la4 = reinterpret_cast<const uint32_t*>(alen)[tg], lb4 = reinterpret_cast<const uint32_t*>(blen)[tg];
#pragma unroll
  for (int k = 0; k < SO; k++) pw[k] = 0u;
  mul_item4<SA, SB, 0>(wa, wb, la4, lb4, pw, lens);
  mul_item4<SA, SB, 1>(wa, wb, la4, lb4, pw, lens);
  mul_item4<SA, SB, 2>(wa, wb, la4, lb4, pw, lens);
  mul_item4<SA, SB, 3>(wa, wb, la4, lb4, pw, lens);
  store_words<SO>(out, tg, pw);
  reinterpret_cast<uint32_t*>(olen)[tg] = lens;
}

template <int SP, int IT>
PB_D void eval_item4(const uint32_t (&w)[SP], uint32_t l4, uint32_t x4, uint32_t& yw) {
  uint32_t iw[(SP + 3) / 4], c[SP];
  item_words<SP, IT>(iw, w);
  const uint32_t len = __byte_perm(l4, 0u, 0x4440 | IT), xv = __byte_perm(x4, 0u, 0x4440 | IT);
  const bool ok = unpack_masked(c, iw, len) && xv <= 16u;
  const uint32_t y = horner17(c, xv & 31u);
  yw |= (ok ? y : 0xFFu) << (8 * IT);
}
template <int SP>
__global__ void __launch_bounds__(PF4_BLOCK) poly_eval_fast4_kernel(const uint8_t* __restrict__ p, const uint8_t* __restrict__ plen,
                                                                    const uint8_t* __restrict__ x, uint8_t* __restrict__ out) {
  const size_t tg = (size_t)blockIdx.x * PF4_BLOCK + threadIdx.x;
  uint32_t w[SP], yw = 0u;
  load_words<SP>(w, p, tg);
  const uint32_t l4 = reinterpret_cast<const uint32_t*>(plen)[tg], x4 = reinterpret_cast<const uint32_t*>(x)[tg];
  eval_item4<SP, 0>(w, l4, x4, yw);
  eval_item4<SP, 1>(w, l4, x4, yw);
  eval_item4<SP, 2>(w, l4, x4, yw);
  eval_item4<SP, 3>(w, l4, x4, yw);
  reinterpret_cast<uint32_t*>(out)[tg] = yw;
}

__global__ void __launch_bounds__(PF4_BLOCK) interpolate4_kernel(const __grid_constant__ CircuitConst cc, const uint8_t* __restrict__ vals,
                                                                 uint8_t* __restrict__ out, uint8_t* __restrict__ olen) {
  const size_t tg = (size_t)blockIdx.x * PF4_BLOCK + threadIdx.x;
  const uint4 v4 = reinterpret_cast<const uint4*>(vals)[tg];
  const uint32_t vw[4] = {v4.x, v4.y, v4.z, v4.w};
  uint32_t fw[4], lens = 0u;
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const uint32_t v[4] = {vw[it] & 0xFFu, __byte_perm(vw[it], 0u, 0x4441), __byte_perm(vw[it], 0u, 0x4442), vw[it] >> 24};
    uint32_t f[4];
    const bool ok = ((((vw[it] & 0x7F7F7F7Fu) + 0x6F6F6F6Fu) | vw[it]) & 0x80808080u) == 0u;    // every byte <= 16
    interpolate(cc, v, f);
    fw[it] = ok ? f[0] | (f[1] << 8) | (f[2] << 16) | (f[3] << 24) : 0u;
    lens |= (ok ? canon_len(f) : 0u) << (8 * it);
  }
  reinterpret_cast<uint4*>(out)[tg] = make_uint4(fw[0], fw[1], fw[2], fw[3]);
  reinterpret_cast<uint32_t*>(olen)[tg] = lens;
}

// poly_divide(p, Z_H) with the context's Z_H = x^4 - 1 (the prover's own call: plonk.h:505 divides t_numer by pk->z_h_x).  The
// division by this monic sparse divisor is q[j] = p[j+4] + q[j+4], remainder p[k] + q[k]: SN - 4 additions per item instead of
// a general long division with a per-item divisor.  Quotient SN - 4 columns, remainder 4 columns, status 0 (or ITEM_INVALID).
template <int SN, int IT>
PB_D void divzh_item4(const uint32_t (&w)[SN], uint32_t l4, uint32_t (&qw)[SN - 4], uint32_t& rw, uint32_t& ql4, uint32_t& rl4, uint32_t& st4) {
  constexpr int SQ = SN - 4;
  uint32_t iw[(SN + 3) / 4], p[SN], q[SQ], r[4];
  item_words<SN, IT>(iw, w);
  const uint32_t len = __byte_perm(l4, 0u, 0x4440 | IT);
  const bool ok = unpack_masked(p, iw, len);
#pragma unroll
  for (int j = SQ - 1; j >= 0; j--) q[j] = j + 4 < SQ ? add17(p[j + 4], q[j + 4]) : p[j + 4];
#pragma unroll
  for (int k = 0; k < 4; k++) r[k] = add17(p[k], q[k]);
  // lengths as poly_divide reports them (poly.h:157-170): the numerator is trimmed first (nl = 0 for a length-0 row); the
  // remainder has min(4, nl) columns, trimmed while > 1 -- for nl < 4 the quotient is zero and r = p
  const uint32_t nl = len == 0u ? 0u : canon_len(p);
  uint32_t rl = nl == 0u ? 0u : canon_len(r);
#pragma unroll
  for (int k = 0; k < SQ; k++) q[k] = ok ? q[k] : 0u;
  put_bytes<SQ, SQ * IT>(qw, q);
  rw = ok ? r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24) : 0u;
  ql4 |= (ok ? canon_len(q) : 0u) << (8 * IT);
  rl4 |= (ok ? rl : 0u) << (8 * IT);
  st4 |= (ok ? 0u : (uint32_t)ITEM_INVALID) << (8 * IT);
}
template <int SN>
__global__ void __launch_bounds__(PF4_BLOCK) poly_divide_zh4_kernel(const uint8_t* __restrict__ num, const uint8_t* __restrict__ nlen,
                                                                    uint8_t* __restrict__ quot, uint8_t* __restrict__ qlen,
                                                                    uint8_t* __restrict__ rem, uint8_t* __restrict__ rlen,
                                                                    uint8_t* __restrict__ status) {
  constexpr int SQ = SN - 4;
  const size_t tg = (size_t)blockIdx.x * PF4_BLOCK + threadIdx.x;
  uint32_t w[SN], qw[SQ], rw[4], ql4 = 0u, rl4 = 0u, st4 = 0u;
  load_words<SN>(w, num, tg);
  const uint32_t l4 = reinterpret_cast<const uint32_t*>(nlen)[tg];
#pragma unroll
  for (int k = 0; k < SQ; k++) qw[k] = 0u;
  divzh_item4<SN, 0>(w, l4, qw, rw[0], ql4, rl4, st4);
  divzh_item4<SN, 1>(w, l4, qw, rw[1], ql4, rl4, st4);
  divzh_item4<SN, 2>(w, l4, qw, rw[2], ql4, rl4, st4);
  divzh_item4<SN, 3>(w, l4, qw, rw[3], ql4, rl4, st4);
  store_words<SQ>(quot, tg, qw);
  reinterpret_cast<uint4*>(rem)[tg] = make_uint4(rw[0], rw[1], rw[2], rw[3]);
  reinterpret_cast<uint32_t*>(qlen)[tg] = ql4;
  reinterpret_cast<uint32_t*>(rlen)[tg] = rl4;
  reinterpret_cast<uint32_t*>(status)[tg] = st4;
}
// the same, one item per thread (ragged tails and unaligned batches)
template <int SN>
__global__ void __launch_bounds__(PF_BLOCK) poly_divide_zh_kernel(const uint8_t* __restrict__ num, const uint8_t* __restrict__ nlen,
                                                                  uint8_t* __restrict__ quot, uint8_t* __restrict__ qlen, uint8_t* __restrict__ rem,
                                                                  uint8_t* __restrict__ rlen, uint8_t* __restrict__ status, size_t n) {
  constexpr int SQ = SN - 4;
  const size_t t = (size_t)blockIdx.x * PF_BLOCK + threadIdx.x;
  if (t >= n) return;
  uint32_t p[SN], q[SQ], r[4];
  const uint32_t len = nlen[t];
  bool ok = len <= (uint32_t)SN;
#pragma unroll
  for (int k = 0; k < SN; k++) { const uint32_t c = num[t * SN + k]; p[k] = (uint32_t)k < len ? c : 0u; ok &= p[k] <= 16u; }
#pragma unroll
  for (int j = SQ - 1; j >= 0; j--) q[j] = j + 4 < SQ ? add17(p[j + 4] & 31u, q[j + 4]) : (p[j + 4] & 31u);
#pragma unroll
  for (int k = 0; k < 4; k++) r[k] = add17(p[k] & 31u, q[k]);
  const uint32_t nl = len == 0u ? 0u : canon_len(p);
  const uint32_t rl = nl == 0u ? 0u : canon_len(r);
#pragma unroll
  for (int k = 0; k < SQ; k++) quot[t * SQ + k] = ok ? (uint8_t)q[k] : 0;
#pragma unroll
  for (int k = 0; k < 4; k++) rem[t * 4 + k] = ok ? (uint8_t)r[k] : 0;
  qlen[t] = ok ? (uint8_t)canon_len(q) : 0;
  rlen[t] = ok ? (uint8_t)rl : 0;
  status[t] = ok ? 0 : ITEM_INVALID;
}

}  // namespace pb
