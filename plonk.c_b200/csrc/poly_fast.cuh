// poly_fast.cuh -- shape-specialised polynomial kernels (BASELINE config 2 and the prover's shapes).
//
// The generic kernels in kernels.cuh keep polynomials in per-thread local arrays with dynamic lengths; they are
// correct for every shape but LSU-bound.  For the fixed strides that matter, the kernels below hold every coefficient
// in a register (all loops unrolled over the STRIDE, lengths applied as masks), and move each block's contiguous
// slice of every byte array through shared memory with 128-bit accesses, so that HBM sees only full-line traffic.
// Semantics are the generic kernels' (= the reference's poly_new trimming + operation + trimming), bit for bit.
#pragma once
#include "kernels.cuh"

namespace pb {

constexpr int PF_BLOCK = 256;

template <int N>
PB_D void load_masked(uint32_t (&r)[N], const uint8_t* row, uint32_t len) {
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = (uint32_t)i < len ? row[i] : 0u;
}

// poly_mul (poly.h:106-122) for fixed strides SA x SB -> SA+SB-1
template <int SA, int SB>
__global__ void __launch_bounds__(PF_BLOCK) poly_mul_fast_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ alen,
                                                                 const uint8_t* __restrict__ b, const uint8_t* __restrict__ blen,
                                                                 uint8_t* __restrict__ out, uint8_t* __restrict__ olen, size_t n) {
  constexpr int SO = SA + SB - 1;
  __shared__ __align__(16) uint8_t sa[PF_BLOCK * SA];
  __shared__ __align__(16) uint8_t sb[PF_BLOCK * SB];
  __shared__ __align__(16) uint8_t so[PF_BLOCK * SO];
  __shared__ __align__(16) uint8_t sla[PF_BLOCK], slb[PF_BLOCK], slo[PF_BLOCK];
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * PF_BLOCK;
  stage_in<SA, PF_BLOCK>(sa, a, first, n);
  stage_in<SB, PF_BLOCK>(sb, b, first, n);
  stage_in<1, PF_BLOCK>(sla, alen, first, n);
  stage_in<1, PF_BLOCK>(slb, blen, first, n);
  __syncthreads();
  if (first + tid < n) {
    uint32_t ra[SA], rb[SB], ro[SO];
    load_masked(ra, sa + tid * SA, sla[tid]);
    load_masked(rb, sb + tid * SB, slb[tid]);
#pragma unroll
    for (int k = 0; k < SO; k++) ro[k] = 0u;
    mul_acc<SA, SB>(ro, ra, rb);            // raw < min(SA,SB) * 2^8
#pragma unroll
    for (int k = 0; k < SO; k++) { ro[k] = red17(ro[k]); so[tid * SO + k] = (uint8_t)ro[k]; }
    // untrimmed length is la' + lb' - 1 with la', lb' the trimmed input lengths; trimming the product gives its
    // canonical length (F17[x] has no zero divisors), and 1 for a zero product
    slo[tid] = (uint8_t)canon_len(ro);
  }
  __syncthreads();
  stage_out<SO, PF_BLOCK>(out, so, first, n);
  stage_out<1, PF_BLOCK>(olen, slo, first, n);
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
  __shared__ __align__(16) uint8_t sn[PF_BLOCK * SN];
  __shared__ __align__(16) uint8_t sd[PF_BLOCK * SD];
  __shared__ __align__(16) uint8_t sq[PF_BLOCK * SQ];
  __shared__ __align__(16) uint8_t sr[PF_BLOCK * SR];
  __shared__ __align__(16) uint8_t sln[PF_BLOCK], sld[PF_BLOCK], slq[PF_BLOCK], slr[PF_BLOCK], sst[PF_BLOCK];
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * PF_BLOCK;
  build_field_tables(ft);
  stage_in<SN, PF_BLOCK>(sn, num, first, n);
  stage_in<SD, PF_BLOCK>(sd, den, first, n);
  stage_in<1, PF_BLOCK>(sln, nlen, first, n);
  stage_in<1, PF_BLOCK>(sld, dlen, first, n);
  __syncthreads();
  if (first + tid < n) {
    uint32_t r[SN], d[SD];
    load_masked(r, sn + tid * SN, sln[tid]);
    load_masked(d, sd + tid * SD, sld[tid]);
    const uint32_t nl = sln[tid] == 0 ? 0u : canon_len(r);      // poly_new trims the inputs first
    const uint32_t dl = sld[tid] == 0 ? 0u : canon_len(d);
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
    for (int k = 0; k < SN; k++) r[k] = red17(r[k]);
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
    slq[tid] = zero_den ? 0 : (uint8_t)ql;
    slr[tid] = zero_den ? 0 : (uint8_t)rl;
    sst[tid] = zero_den ? 1 : 0;                                 // "Division by zero polynomial", poly.h:125-128
  }
  __syncthreads();
  stage_out<SQ, PF_BLOCK>(quot, sq, first, n);
  stage_out<SR, PF_BLOCK>(rem, sr, first, n);
  stage_out<1, PF_BLOCK>(qlen, slq, first, n);
  stage_out<1, PF_BLOCK>(rlen, slr, first, n);
  stage_out<1, PF_BLOCK>(status, sst, first, n);
}

// poly_eval (poly.h:265-272) for a fixed stride
template <int SP>
__global__ void __launch_bounds__(PF_BLOCK) poly_eval_fast_kernel(const uint8_t* __restrict__ p, const uint8_t* __restrict__ plen,
                                                                  const uint8_t* __restrict__ x, uint8_t* __restrict__ out, size_t n) {
  __shared__ __align__(16) uint8_t spv[PF_BLOCK * SP];
  __shared__ __align__(16) uint8_t sl[PF_BLOCK], sx[PF_BLOCK], sy[PF_BLOCK];
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * PF_BLOCK;
  stage_in<SP, PF_BLOCK>(spv, p, first, n);
  stage_in<1, PF_BLOCK>(sl, plen, first, n);
  stage_in<1, PF_BLOCK>(sx, x, first, n);
  __syncthreads();
  if (first + tid < n) {
    const uint32_t len = sl[tid], xv = sx[tid];
    uint32_t y = 0u;
#pragma unroll
    for (int k = SP - 1; k >= 0; k--) y = (uint32_t)k < len ? red17(y * xv + spv[tid * SP + k]) : y;   // Horner from the top
    sy[tid] = (uint8_t)y;
  }
  __syncthreads();
  stage_out<1, PF_BLOCK>(out, sy, first, n);
}

}  // namespace pb
