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
template <int N>
PB_D void unpack_masked(uint32_t (&r)[N], const uint32_t (&w)[(N + 3) / 4], uint32_t len) {
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = (uint32_t)i < len ? ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu) : 0u;
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
    unpack_masked(ra, wa, alen[t]);
    unpack_masked(rb, wb, blen[t]);
#pragma unroll
    for (int k = 0; k < SO; k++) ro[k] = 0u;
    mul_acc<SA, SB>(ro, ra, rb);            // raw < min(SA,SB) * 2^8
#pragma unroll
    for (int k = 0; k < SO; k++) { ro[k] = red17(ro[k]); so[tid * SO + k] = (uint8_t)ro[k]; }
    // F17[x] has no zero divisors: trimming the product gives its canonical length (1 for a zero product)
    olen[t] = (uint8_t)canon_len(ro);
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
    unpack_masked(r, wn, nlen0);
    unpack_masked(d, wd, dlen0);
    const uint32_t nl = nlen0 == 0 ? 0u : canon_len(r);      // poly_new trims the inputs first
    const uint32_t dl = dlen0 == 0 ? 0u : canon_len(d);
    bool zero_den = true;
#pragma unroll
    for (int j = 0; j < SD; j++) zero_den &= d[j] == 0u;
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
    status[t] = zero_den ? 1 : 0;                                // "Division by zero polynomial", poly.h:125-128
  }
  __syncthreads();
  store_slice<SQ>(quot, sq, first, n);
  store_slice<SR>(rem, sr, first, n);
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
  unpack_masked(c, w, len);
  uint32_t y = 0u;
#pragma unroll
  for (int k = SP - 1; k >= 0; k--) y = red17(y * xv + c[k]);    // Horner from the top; masked coefficients are 0 and y stays 0 above len
  out[t] = (uint8_t)y;
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
    uint32_t y = 0u;
#pragma unroll
    for (int k = 5; k >= 0; k--) y = red17(y * xv + ra[k]);
    evals[t] = (uint8_t)y;
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

}  // namespace pb
