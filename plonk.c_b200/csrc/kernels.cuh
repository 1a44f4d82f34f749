// kernels.cuh -- __global__ entry points.  One item per thread everywhere: items are independent, a few
// dozen bytes each, so there is no inter-thread communication beyond staging tables in shared memory.
#pragma once
#include <type_traits>
#include "prover.cuh"
#include "verifier.cuh"
#include "wire.cuh"

namespace pb {

#ifndef PB_BLOCK
#define PB_BLOCK 128
#endif
constexpr int BLOCK = PB_BLOCK;   // threads per block for the heavy kernels (prove / verify / pairing)
constexpr int BLOCK_LIGHT = 256;   // for the byte-streaming kernels
constexpr int POLY_MAX = 64;

// The inverse tables: x^15 mod 17 is the reference's table (hf.h:145-180), x^99 mod 101 is literally what gf_inv computes
// (gf.h:159-162).  They are evaluated at COMPILE time (constexpr, same square-and-multiply) into a 288-byte device constant
// that every block copies into shared memory with 72 four-byte loads -- the light kernels (one pairing or one scalar
// multiplication per thread) would otherwise spend ~10 % of their instructions re-deriving 288 powers per block.
struct FieldTablesImage {
  uint32_t w[sizeof(FieldTables) / 4];
  constexpr FieldTablesImage() : w{} {
    static_assert(sizeof(FieldTables) == 32 + 256, "FieldTables layout: inv17[32] then inv101[256]");
    auto cpow = [](uint32_t base, uint32_t e, uint32_t p) {
      uint32_t r = 1u;
      while (e) { if (e & 1u) r = r * base % p; base = base * base % p; e >>= 1; }
      return r;
    };
    for (uint32_t i = 0; i < 32; i++) w[i / 4] |= (i < 17u ? cpow(i, 15u, 17u) : 0u) << (8u * (i % 4u));
    for (uint32_t i = 0; i < 256; i++) w[8 + i / 4] |= cpow(i % 101u, 99u, 101u) << (8u * (i % 4u));
  }
};
__device__ __align__(16) const FieldTablesImage g_field_tables_image = FieldTablesImage();
PB_D void build_field_tables(FieldTables& ft) {
  for (int i = threadIdx.x; i < (int)(sizeof(FieldTables) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&ft)[i] = g_field_tables_image.w[i];
}

// ---- 1-D bulk asynchronous copies (cp.async.bulk, the TMA engine without a tensor map; UBLKCP in SASS) with an mbarrier.
// One elected thread issues the copies of a block's tables and input slices; they cost no LSU issue slots and ONE memory
// round trip, where per-thread LDG -> STS loops cost one round trip per iteration (ncu r1: the table loop and the three
// staging loops were 13 % of the prover's stall samples, profiles/r2/NOTES.md).  Sizes and addresses are multiples of 16.
#ifndef PB_BULK
#define PB_BULK 1
#endif
PB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
PB_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PB_D void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
PB_D void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
PB_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// shared -> global: the writers fence their generic-proxy stores, the block synchronises, one thread issues and waits until the
// engine has READ shared memory (the block may then retire; the global writes complete on their own)
PB_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
PB_D void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
PB_D void bulk_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

PB_D G1 load_g1(const uint8_t* p) { return G1{p[0], p[1], p[2] != 0 ? 1u : 0u}; }
PB_D void store_g1(uint8_t* p, G1 g) { p[0] = (uint8_t)g.x; p[1] = (uint8_t)g.y; p[2] = (uint8_t)g.inf; }

// ------------------------------------------------------------------ family (1)
template <int FIELD>
PB_D uint32_t field_apply(const FieldTables& ft, int op, uint32_t a, uint32_t b) {
  if (FIELD == 17) {
    switch (op) {
      case 0: return add17(a, b);
      case 1: return sub17(a, b);
      case 2: return mul17(a, b);
      case 3: return mul17(a, inv17(ft, b & 31u));   // bytes outside [0,17) are not field elements: the index is kept inside the table
      case 4: return neg17(a);
      case 5: return inv17(ft, a & 31u);
      default: return pow17(a, b);
    }
  } else {
    switch (op) {
      case 0: return add101(a, b);
      case 1: return sub101(a, b);
      case 2: return mul101(a, b);
      case 3: return mul101(a, inv101(ft, b));
      case 4: return neg101(a);
      case 5: return inv101(ft, a);
      default: return pow101(a, b);
    }
  }
}

// 16 elements per thread through 128-bit loads/stores; the tail (n % 16) is handled byte-wise by the last threads.
// OP is a template parameter: the operation is resolved at compile time (one small kernel per field and operation) instead of
// a switch on a kernel argument inside a loop that has ~10 instructions per element to spend
template <int FIELD, int OP>
__global__ void __launch_bounds__(BLOCK_LIGHT) field_op_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                               uint8_t* __restrict__ out, size_t n) {
  constexpr int op = OP;
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  const size_t nvec = n / 16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    uint4 va = reinterpret_cast<const uint4*>(a)[i];
    uint4 vb = b ? reinterpret_cast<const uint4*>(b)[i] : make_uint4(0, 0, 0, 0);
    uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w}, wo[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
      if constexpr (OP == 4) {
        // negation of four canonical residues at once: p - a byte-wise (no borrow: a <= p - 1), zero bytes stay zero
        const uint32_t nz = (((wa[w] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | wa[w]) & 0x80808080u;     // bit 7 of every non-zero byte
        const bool canon = ((((wa[w] & 0x7F7F7F7Fu) + (0x80u - FIELD) * 0x01010101u) | wa[w]) & 0x80808080u) == 0u;   // every byte < p
        if (canon) { wo[w] = ((uint32_t)FIELD * 0x01010101u - wa[w]) & ((nz >> 7) * 0xFFu); continue; }
      }
      uint32_t r = 0;
#pragma unroll
      for (int k = 0; k < 4; k++)
        r |= field_apply<FIELD>(ft, op, (wa[w] >> (8 * k)) & 0xFFu, (wb[w] >> (8 * k)) & 0xFFu) << (8 * k);
      wo[w] = r;
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(wo[0], wo[1], wo[2], wo[3]);
  }
  for (size_t i = nvec * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (uint8_t)field_apply<FIELD>(ft, op, a[i], b ? b[i] : 0u);
}

// hf_pow / gf_pow (hf.h:127-137, gf.h:140-151) with a one-byte exponent.  The reference's square-and-multiply returns
// (a mod p)^e with 0^0 = 1; over a prime field that is alog[(log a * e) mod (p - 1)] for a != 0, with discrete logarithms
// to the primitive root 3 (F17) / 2 (F101).  Both tables are built at COMPILE time and are at most 101 bytes each, i.e.
// they sit inside one 128-byte row of shared memory: no bank conflicts for any access pattern.  (Round 1 looked a^k up in
// a 10 KB per-block table: 10.5 M bank conflicts per 2^27-element launch, 58 % of the HBM roofline for gf_pow.)
template <int FIELD>
struct PowTablesImage {
  uint32_t w[64];                           // bytes 0..127: log[a] (log[0] unused), bytes 128..255: alog[k], k < p - 1
  constexpr PowTablesImage() : w{} {
    constexpr uint32_t P = FIELD, G = FIELD == 17 ? 3u : 2u;
    uint32_t v = 1u;
    for (uint32_t k = 0; k < P - 1; k++) {
      w[(128 + k) / 4] |= v << (8u * ((128 + k) % 4u));      // alog[k] = g^k
      w[v / 4] |= k << (8u * (v % 4u));                      // log[g^k] = k
      v = v * G % P;
    }
  }
};
__device__ __align__(16) const PowTablesImage<17> g_pow_tables_17 = PowTablesImage<17>();
__device__ __align__(16) const PowTablesImage<101> g_pow_tables_101 = PowTablesImage<101>();

template <int FIELD>
__global__ void __launch_bounds__(BLOCK_LIGHT) field_pow_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ e,
                                                                uint8_t* __restrict__ out, size_t n) {
  constexpr uint32_t ORD = FIELD - 1;
  __shared__ __align__(128) uint8_t tab[256];
  if (threadIdx.x < 64) reinterpret_cast<uint32_t*>(tab)[threadIdx.x] = FIELD == 17 ? g_pow_tables_17.w[threadIdx.x] : g_pow_tables_101.w[threadIdx.x];
  __syncthreads();
  auto one = [&](uint32_t base, uint32_t ex) -> uint32_t {
    const uint32_t b = FIELD == 17 ? red17(base) : red101(base);
    const uint32_t t = tab[b] * ex;                                                    // log a * e < 100 * 256
    const uint32_t k = FIELD == 17 ? (t & 15u) : t - 100u * ((t * 5243u) >> 19);       // mod (p - 1); floor(t / 100) exact for t < 43 690
    const uint32_t r = tab[128u + k];
    return b == 0u ? (ex == 0u ? 1u : 0u) : r;                                         // 0^0 = 1, 0^e = 0
  };
  const size_t nvec = n / 16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 va = reinterpret_cast<const uint4*>(a)[i], ve = reinterpret_cast<const uint4*>(e)[i];
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, we[4] = {ve.x, ve.y, ve.z, ve.w};
    uint32_t wo[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
      uint32_t r = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) r |= one((wa[w] >> (8 * k)) & 0xFFu, (we[w] >> (8 * k)) & 0xFFu) << (8 * k);
      wo[w] = r;
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(wo[0], wo[1], wo[2], wo[3]);
  }
  for (size_t i = nvec * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint8_t)one(a[i], e[i]);
  (void)ORD;
}

// hf_new / gf_new (hf.h:25-35, gf.h:24-34): C remainder of a signed 64-bit value, negatives folded up
template <int FIELD>
__global__ void __launch_bounds__(BLOCK_LIGHT) field_new_kernel(const long long* __restrict__ v, uint8_t* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (uint8_t)(FIELD == 17 ? new17(v[i]) : new101(v[i]));
}

// ------------------------------------------------------------------ family (2), generic shapes
// Polynomials in per-thread local arrays; dynamic lengths.  (Fixed-shape fast paths: poly_fast.cuh.)
// Inputs are validated per item: a length above the row stride, a coefficient byte above 16, a shift that does not fit the
// output row are not values any reference path produces (HF is "always kept in range", hf.h:11-14; poly_new exits on a bad
// length, poly.h:29-32).  Such an item is REPORTED, never computed on: olen[i] = 0 with a zero row where the result has a
// length (every valid result has len >= 1 there), status[i] = 2 where a status array exists, 0xFF for poly_eval.
constexpr uint8_t ITEM_INVALID = 2;
struct LPoly { uint8_t c[2 * POLY_MAX]; int len; };

// poly_new: trim while len > 1 (poly.h:20-24).  false = the row is not a polynomial over F17 of at most `stride` coefficients
PB_D bool lp_load(LPoly& p, const uint8_t* row, int len, int stride) {
  if (len > stride) { p.len = 0; return false; }
  bool ok = true;
  for (int i = 0; i < len; i++) { p.c[i] = row[i]; ok &= row[i] <= 16; }
  while (len > 1 && p.c[len - 1] == 0) len--;
  p.len = len;
  return ok;
}
PB_D void lp_trim(LPoly& p) { while (p.len > 1 && p.c[p.len - 1] == 0) p.len--; }
PB_D void lp_store(const LPoly& p, uint8_t* row, int stride, uint8_t* len) {
  for (int i = 0; i < stride; i++) row[i] = i < p.len ? p.c[i] : 0;
  *len = (uint8_t)p.len;
}
PB_D void lp_invalid(uint8_t* row, int stride, uint8_t* len) {
  for (int i = 0; i < stride; i++) row[i] = 0;
  *len = 0;
}

__global__ void __launch_bounds__(BLOCK_LIGHT) poly_binop_kernel(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ alen, int sa,
                                                                 const uint8_t* __restrict__ b, const uint8_t* __restrict__ blen, int sb,
                                                                 uint8_t* __restrict__ out, uint8_t* __restrict__ olen, int so, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  LPoly pa, pb, r;
  bool ok = lp_load(pa, a + i * sa, alen[i], sa);
  ok &= lp_load(pb, b + i * sb, blen[i], sb);
  if (!ok || pa.len == 0 || pb.len == 0) { lp_invalid(out + i * so, so, olen + i); return; }
  if (op == 2) {                                                 // poly_mul, poly.h:106-122
    r.len = pa.len + pb.len - 1;
    for (int k = 0; k < r.len; k++) {
      uint32_t s = 0;
      int lo = k - (pb.len - 1) > 0 ? k - (pb.len - 1) : 0, hi = k < pa.len - 1 ? k : pa.len - 1;
      for (int x = lo; x <= hi; x++) s += (uint32_t)pa.c[x] * pb.c[k - x];
      r.c[k] = (uint8_t)red17(s);
    }
  } else {                                                       // poly_add / poly_sub, poly.h:72-104
    r.len = pa.len > pb.len ? pa.len : pb.len;
    for (int k = 0; k < r.len; k++) {
      uint32_t x = k < pa.len ? pa.c[k] : 0u, y = k < pb.len ? pb.c[k] : 0u;
      r.c[k] = (uint8_t)(op == 0 ? add17(x, y) : sub17(x, y));
    }
  }
  lp_trim(r);
  lp_store(r, out + i * so, so, olen + i);
}

__global__ void __launch_bounds__(BLOCK_LIGHT) poly_divide_kernel(const uint8_t* __restrict__ num, const uint8_t* __restrict__ nlen, int sn,
                                                                  const uint8_t* __restrict__ den, const uint8_t* __restrict__ dlen, int sd,
                                                                  uint8_t* __restrict__ quot, uint8_t* __restrict__ qlen, int sq,
                                                                  uint8_t* __restrict__ rem, uint8_t* __restrict__ rlen, int sr,
                                                                  uint8_t* __restrict__ status, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  LPoly pn, pd, q;
  bool ok = lp_load(pn, num + i * sn, nlen[i], sn);
  ok &= lp_load(pd, den + i * sd, dlen[i], sd);
  bool zero_den = true;                                          // poly_is_zero, poly.h:55-64
  for (int k = 0; k < pd.len; k++) zero_den &= pd.c[k] == 0;
  // the quotient / remainder rows must hold what the division produces (the natural strides sn - sd + 1 and sd - 1 always do)
  const int need_q = pn.len >= pd.len ? pn.len - pd.len + 1 : 1, need_r = pd.len - 1 < pn.len ? pd.len - 1 : pn.len;
  if (ok && !zero_den && (need_q > sq || need_r > sr)) ok = false;
  if (!ok || zero_den) {                                         // "Division by zero polynomial", poly.h:125-128 -> 1; not a polynomial -> 2
    status[i] = ok ? 1 : ITEM_INVALID;
    for (int k = 0; k < sq; k++) quot[i * sq + k] = 0;
    for (int k = 0; k < sr; k++) rem[i * sr + k] = 0;
    qlen[i] = 0; rlen[i] = 0;
    return;
  }
  status[i] = 0;
  const int nl = pn.len, dl = pd.len;
  for (int k = 0; k < nl; k++) q.c[k] = 0;
  const uint32_t lead_inv = inv17(ft, pd.c[dl - 1]);
  for (int k = nl - 1; k >= dl - 1; k--) {                       // long division, poly.h:147-155
    uint32_t f = mul17(pn.c[k], lead_inv);
    q.c[k - (dl - 1)] = (uint8_t)f;
    for (int j = 0; j < dl; j++) pn.c[k - j] = (uint8_t)sub17(pn.c[k - j], mul17(f, pd.c[dl - 1 - j]));
  }
  q.len = nl >= dl ? nl - dl + 1 : 1;
  if (nl < dl) q.c[0] = 0;
  lp_trim(q);
  int rl = dl - 1;                                               // poly.h:163-170; 0 when the divisor is a constant
  if (rl > nl) rl = nl;
  while (rl > 1 && pn.c[rl - 1] == 0) rl--;
  pn.len = rl;
  lp_store(q, quot + i * sq, sq, qlen + i);
  lp_store(pn, rem + i * sr, sr, rlen + i);
}

__global__ void __launch_bounds__(BLOCK_LIGHT) poly_eval_kernel(const uint8_t* __restrict__ p, const uint8_t* __restrict__ plen, int sp,
                                                                const uint8_t* __restrict__ x, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* row = p + i * sp;
  uint32_t y = 0, xv = x[i];
  const int len = plen[i];
  bool ok = len <= sp && xv <= 16u;
  for (int k = (ok ? len : 0) - 1; k >= 0; k--) { ok &= row[k] <= 16; y = red17(y * xv + (row[k] & 31u)); }   // Horner, poly.h:265-272
  out[i] = ok ? (uint8_t)y : 0xFF;
}

__global__ void __launch_bounds__(BLOCK_LIGHT) poly_unop_kernel(int op, const uint8_t* __restrict__ p, const uint8_t* __restrict__ plen, int sp,
                                                                const uint8_t* __restrict__ k, uint8_t* __restrict__ out,
                                                                uint8_t* __restrict__ olen, int so, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  LPoly a, r;
  bool ok = lp_load(a, p + i * sp, plen[i], sp) && a.len >= 1;
  uint32_t kv = k ? k[i] : 0u;
  if ((op == 0 || op == 3) && kv > 16u) ok = false;             // the scalar of poly_scale / poly_add_hf is a field element
  bool z = true;
  for (int j = 0; j < a.len; j++) z &= a.c[j] == 0;
  if (op == 2 && !z && a.len + (int)kv > so) ok = false;        // the shifted polynomial must fit the output row
  if (!ok) { lp_invalid(out + i * so, so, olen + i); return; }
  if (op == 0) {                                                 // poly_scale, poly.h:179-197: scalar 0 -> [0]
    if (kv == 0) { r.len = 1; r.c[0] = 0; }
    else { r.len = a.len; for (int j = 0; j < a.len; j++) r.c[j] = (uint8_t)mul17(a.c[j], kv); lp_trim(r); }
  } else if (op == 1) {                                          // poly_negate, poly.h:240-254
    r.len = a.len;
    for (int j = 0; j < a.len; j++) r.c[j] = (uint8_t)neg17(a.c[j]);
    lp_trim(r);
  } else if (op == 2) {                                          // poly_shift, poly.h:199-216: zero stays [0]
    if (z) { r.len = 1; r.c[0] = 0; }
    else {
      r.len = a.len + (int)kv;
      for (int j = 0; j < r.len; j++) r.c[j] = j >= (int)kv ? a.c[j - kv] : 0;
      lp_trim(r);
    }
  } else {                                                       // poly_add_hf, poly.h:67-70: in place, not trimmed
    r = a;
    r.c[0] = (uint8_t)add17(a.c[0], kv);
  }
  lp_store(r, out + i * so, so, olen + i);
}

__global__ void __launch_bounds__(BLOCK_LIGHT) poly_slice_kernel(const uint8_t* __restrict__ p, const uint8_t* __restrict__ plen, int sp,
                                                                 const uint8_t* __restrict__ start, const uint8_t* __restrict__ end,
                                                                 uint8_t* __restrict__ out, uint8_t* __restrict__ olen, int so,
                                                                 uint8_t* __restrict__ status, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  LPoly a, r;
  const bool ok = lp_load(a, p + i * sp, plen[i], sp);
  int s = start[i], e = end[i];
  if (!ok || s >= e || e > a.len) {                              // "Invalid slice indices", poly.h:219-222 -> 1; not a polynomial -> 2
    status[i] = ok ? 1 : ITEM_INVALID; olen[i] = 0;
    for (int j = 0; j < so; j++) out[i * so + j] = 0;
    return;
  }
  status[i] = 0;
  r.len = e - s;
  for (int j = 0; j < r.len; j++) r.c[j] = a.c[s + j];
  lp_trim(r);
  lp_store(r, out + i * so, so, olen + i);
}

// poly_lagrange, poly.h:288-321 (product form, O(len^3)); len <= 16
__global__ void __launch_bounds__(BLOCK_LIGHT) poly_lagrange_kernel(const uint8_t* __restrict__ xs, const uint8_t* __restrict__ ys, int len,
                                                                    uint8_t* __restrict__ out, uint8_t* __restrict__ olen, int so,
                                                                    uint8_t* __restrict__ status, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* x = xs + i * len;
  const uint8_t* y = ys + i * len;
  uint8_t l[17], lj[17], t[17];
  for (int k = 0; k <= len; k++) l[k] = 0;
  bool bytes_ok = true;
  for (int k = 0; k < len; k++) bytes_ok &= x[k] <= 16 && y[k] <= 16;
  if (!bytes_ok) {                                               // not field elements: reported, not computed on
    status[i] = ITEM_INVALID; olen[i] = 0;
    for (int k = 0; k < so; k++) out[i * so + k] = 0;
    return;
  }
  bool dup = false;
  for (int j = 0; j < len && !dup; j++) {
    int ljn = 1;
    lj[0] = 1;
    for (int m = 0; m < len; m++) {
      if (m == j) continue;
      uint32_t dinv = inv17(ft, sub17(x[j], x[m]));
      if (dinv == 0) { dup = true; break; }                      // "x points must be unique", poly.h:298-301
      uint32_t c0 = neg17(mul17(dinv, x[m]));
      for (int k = 0; k <= ljn; k++) t[k] = 0;                  // lj *= (c0 + dinv x)
      for (int k = 0; k < ljn; k++) {
        t[k] = (uint8_t)add17(t[k], mul17(lj[k], c0));
        t[k + 1] = (uint8_t)add17(t[k + 1], mul17(lj[k], dinv));
      }
      ljn++;
      for (int k = 0; k < ljn; k++) lj[k] = t[k];
    }
    if (dup) break;
    for (int k = 0; k < ljn; k++) l[k] = (uint8_t)add17(l[k], mul17(lj[k], y[j]));
  }
  if (dup) {
    status[i] = 1; olen[i] = 0;
    for (int k = 0; k < so; k++) out[i * so + k] = 0;
    return;
  }
  status[i] = 0;
  int ln = len;
  while (ln > 1 && l[ln - 1] == 0) ln--;
  for (int k = 0; k < so; k++) out[i * so + k] = k < ln ? l[k] : 0;
  olen[i] = (uint8_t)ln;
}

__global__ void __launch_bounds__(BLOCK_LIGHT) interpolate_kernel(const __grid_constant__ CircuitConst cc, const uint8_t* __restrict__ vals,
                                                                  uint8_t* __restrict__ out, uint8_t* __restrict__ olen, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w = reinterpret_cast<const uint32_t*>(vals)[i];
  uint32_t v[4] = {w & 0xFFu, (w >> 8) & 0xFFu, (w >> 16) & 0xFFu, w >> 24}, r[4];
  const bool ok = v[0] <= 16u && v[1] <= 16u && v[2] <= 16u && v[3] <= 16u;
  interpolate(cc, v, r);
  reinterpret_cast<uint32_t*>(out)[i] = ok ? r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24) : 0u;
  olen[i] = ok ? (uint8_t)canon_len(r) : 0;                      // a value byte above 16 is not a field element: reported as length 0
}

// matrix_mul, matrix.h:81-98; dims <= 8
__global__ void __launch_bounds__(BLOCK_LIGHT) matrix_mul_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out,
                                                                 int m, int k, int c, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* A = a + i * m * k;
  const uint8_t* B = b + i * k * c;
  uint8_t* O = out + i * m * c;
  for (int r = 0; r < m; r++)
    for (int j = 0; j < c; j++) {
      uint32_t s = 0;
      for (int x = 0; x < k; x++) s += (uint32_t)A[r * k + x] * B[x * c + j];
      O[r * c + j] = (uint8_t)red17(s);
    }
}

// matrix_gauss_jordan, matrix.h:100-149: in-place RREF with the reference's pivot search (walk down the rows of the
// lead column, then advance the column), row swap, normalisation by 1/pivot and elimination of every other row.
PB_D void rref(const FieldTables& ft, uint8_t* M, int rows, int cols) {
  int lead = 0;
  for (int r = 0; r < rows; r++) {
    if (cols <= lead) return;
    int p = r;
    while (M[p * cols + lead] == 0) {
      p++;
      if (p == rows) { p = r; lead++; if (lead == cols) return; }
    }
    if (p != r) for (int c = 0; c < cols; c++) { uint8_t t = M[p * cols + c]; M[p * cols + c] = M[r * cols + c]; M[r * cols + c] = t; }
    uint32_t d = M[r * cols + lead];
    if (d != 0) { uint32_t di = inv17(ft, d); for (int c = 0; c < cols; c++) M[r * cols + c] = (uint8_t)mul17(M[r * cols + c], di); }
    for (int q = 0; q < rows; q++) {
      if (q == r) continue;
      uint32_t mult = M[q * cols + lead];
      for (int c = 0; c < cols; c++) M[q * cols + c] = (uint8_t)sub17(M[q * cols + c], mul17(M[r * cols + c], mult));
    }
    lead++;
  }
}

// rows <= 8, cols <= 16
__global__ void __launch_bounds__(BLOCK_LIGHT) matrix_rref_kernel(uint8_t* __restrict__ a, int rows, int cols, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t M[8 * 16];
  for (int k = 0; k < rows * cols; k++) M[k] = (uint8_t)red17(a[i * rows * cols + k]);   // entries are reduced on load: table indices stay in range
  rref(ft, M, rows, cols);
  for (int k = 0; k < rows * cols; k++) a[i * rows * cols + k] = M[k];
}

// matrix_inv = Gauss-Jordan on (M | I), matrix.h:151-176, without singularity detection; dim <= 8
__global__ void __launch_bounds__(BLOCK_LIGHT) matrix_inv_kernel(const uint8_t* __restrict__ a, uint8_t* __restrict__ out, int dim, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t M[8 * 16];
  const int rows = dim, cols = 2 * dim;
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) M[r * cols + c] = c < dim ? (uint8_t)red17(a[i * dim * dim + r * dim + c]) : (c - dim == r ? 1 : 0);
  rref(ft, M, rows, cols);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < dim; c++) out[i * dim * dim + r * dim + c] = M[r * cols + dim + c];
}

// ------------------------------------------------------------------ family (3)
__global__ void __launch_bounds__(BLOCK_LIGHT) g1_op_kernel(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                            uint8_t* __restrict__ out, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1 x = load_g1(a + 3 * i), r;
  if (op == 0) r = g1_add(ft, x, load_g1(b + 3 * i));
  else if (op == 1) r = g1_double(ft, x);
  else r = g1_neg(x);
  store_g1(out + 3 * i, r);
}

template <typename S>
__global__ void __launch_bounds__(BLOCK) g1_mul_kernel(const uint8_t* __restrict__ p, const S* __restrict__ s, uint8_t* __restrict__ out, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  store_g1(out + 3 * i, g1_mul(ft, load_g1(p + 3 * i), (uint64_t)s[i]));
}

__global__ void __launch_bounds__(BLOCK_LIGHT) g1_on_curve_kernel(const uint8_t* __restrict__ p, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = g1_is_on_curve(load_g1(p + 3 * i)) ? 1 : 0;
}

__global__ void __launch_bounds__(BLOCK_LIGHT) g2_op_kernel(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                            uint8_t* __restrict__ out, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G2 x{a[2 * i], a[2 * i + 1]}, r;
  if (op == 0) r = g2_add(ft, x, G2{b[2 * i], b[2 * i + 1]});
  else r = g2_neg(x);
  out[2 * i] = (uint8_t)r.x; out[2 * i + 1] = (uint8_t)r.y;
}

__global__ void __launch_bounds__(BLOCK) g2_mul_kernel(const uint8_t* __restrict__ p, const uint64_t* __restrict__ s, uint8_t* __restrict__ out, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G2 r = g2_mul(ft, G2{p[2 * i], p[2 * i + 1]}, s[i]);
  out[2 * i] = (uint8_t)r.x; out[2 * i + 1] = (uint8_t)r.y;
}

// T[i][c] = g1_mul(g1s[i], c), the reference's own double-and-add (context creation)
__global__ void srs_table_kernel(const uint8_t* __restrict__ g1s, uint32_t srs_len, uint32_t rows, uint32_t* __restrict__ T) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < rows * 17u; k += blockDim.x) {
    uint32_t i = k / 17u, c = k % 17u;
    G1 r = i < srs_len ? g1_mul(ft, load_g1(g1s + 3 * i), c) : g1_identity();
    T[k] = pack_g1(r.x, r.y, r.inf);
  }
}

// srs_eval_at_s (srs.h:53-68) for arbitrary polynomials against the context's full table
__global__ void __launch_bounds__(BLOCK) commit_kernel(const uint32_t* __restrict__ Tg, uint32_t srs_len, const uint8_t* __restrict__ polys,
                                                       const uint8_t* __restrict__ plen, int sp, uint8_t* __restrict__ out,
                                                       uint8_t* __restrict__ status, size_t n) {
  extern __shared__ uint32_t Ts[];
  __shared__ FieldTables ft;
  build_field_tables(ft);
  for (uint32_t k = threadIdx.x; k < srs_len * 17u; k += blockDim.x) Ts[k] = Tg[k];
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* row = polys + i * sp;
  int len = plen[i];
  bool ok = len <= sp;
  for (int k = 0; k < (ok ? len : 0); k++) ok &= row[k] <= 16;   // a coefficient byte above 16 would index past its table row
  if (!ok) {
    status[i] = ITEM_INVALID;
    out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = 0;
    return;
  }
  while (len > 1 && row[len - 1] == 0) len--;                   // poly_new trims first
  if ((uint32_t)len > srs_len) {                                 // "exceeds SRS size", srs.h:54-57
    status[i] = 1;
    out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = 0;
    return;
  }
  status[i] = 0;
  G1 acc = g1_identity();
  for (int k = 0; k < len; k++) acc = g1_add(ft, acc, unpack_g1(Ts[k * 17 + row[k]]));
  store_g1(out + 3 * i, acc);
}

// srs_eval_at_s (srs.h:53-68) against an SRS handed over per call (no context, no table): the reference's loop as written --
// g1_mul(g1s[i], coeff_i) by double-and-add, summed left to right.  For callers that hold an SRS struct and a handful of
// polynomials (the drop-in srs_eval_at_s): one launch instead of one per term.
__global__ void __launch_bounds__(BLOCK_LIGHT) commit_raw_kernel(const uint8_t* __restrict__ g1s, uint32_t srs_len, const uint8_t* __restrict__ polys,
                                                                 const uint8_t* __restrict__ plen, int sp, int trim, uint8_t* __restrict__ out,
                                                                 uint8_t* __restrict__ status, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* row = polys + i * sp;
  int len = plen[i];
  bool ok = len <= sp;
  for (int k = 0; k < (ok ? len : 0); k++) ok &= row[k] <= 16;
  if (!ok) {                                                     // not a polynomial over F17: reported, not computed on
    status[i] = ITEM_INVALID;
    out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = 0;
    return;
  }
  // trim: the row is passed through poly_new first, as in the other batch entry points; otherwise len is used as given -- the
  // reference's loop runs over POLY.len terms, and a trailing zero term is NOT a no-op when the accumulator is an "identity with
  // coordinates" (g1_add returns *b when a is infinite, g1.h:60)
  while (trim && len > 1 && row[len - 1] == 0) len--;
  if ((uint32_t)len > srs_len) {                                 // "exceeds SRS size", srs.h:54-57
    status[i] = 1;
    out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = 0;
    return;
  }
  status[i] = 0;
  G1 acc = g1_identity();
  for (int k = 0; k < len; k++) acc = g1_add(ft, acc, g1_mul(ft, load_g1(g1s + 3 * k), (uint64_t)row[k]));
  store_g1(out + 3 * i, acc);
}

// ------------------------------------------------------------------ family (4)
__global__ void __launch_bounds__(BLOCK_LIGHT) gtp_mul_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GT r = gt_mul(GT{a[2 * i], a[2 * i + 1]}, GT{b[2 * i], b[2 * i + 1]});
  out[2 * i] = (uint8_t)r.a; out[2 * i + 1] = (uint8_t)r.b;
}
__global__ void __launch_bounds__(BLOCK_LIGHT) gtp_pow_kernel(const uint8_t* __restrict__ a, const uint64_t* __restrict__ e, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GT r = gt_pow(GT{a[2 * i], a[2 * i + 1]}, e[i]);
  out[2 * i] = (uint8_t)r.a; out[2 * i + 1] = (uint8_t)r.b;
}
__global__ void __launch_bounds__(BLOCK_LIGHT) line_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Line l = line_through(load_g1(a + 3 * i), load_g1(b + 3 * i));
  out[3 * i] = (uint8_t)l.x; out[3 * i + 1] = (uint8_t)l.y; out[3 * i + 2] = (uint8_t)l.c;
}
__global__ void __launch_bounds__(BLOCK) pairing_kernel(const uint8_t* __restrict__ p, const uint8_t* __restrict__ q, uint8_t* __restrict__ out, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GT r = pairing17(ft, load_g1(p + 3 * i), G2{q[2 * i], q[2 * i + 1]});
  out[2 * i] = (uint8_t)r.a; out[2 * i + 1] = (uint8_t)r.b;
}
__global__ void __launch_bounds__(BLOCK) pairing_f_kernel(uint64_t r, const uint8_t* __restrict__ p, const uint8_t* __restrict__ q, uint8_t* __restrict__ out, size_t n) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GT f = miller(ft, r, load_g1(p + 3 * i), G2{q[2 * i], q[2 * i + 1]});
  out[2 * i] = (uint8_t)f.a; out[2 * i + 1] = (uint8_t)f.b;
}

// ------------------------------------------------------------------ protocol
__global__ void __launch_bounds__(BLOCK_LIGHT) satisfy_kernel(const __grid_constant__ CircuitConst cc, const uint8_t* __restrict__ wit,
                                                              uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* w = wit + 12 * i;
  bool ok = true;
#pragma unroll
  for (int g = 0; g < 4; g++) {
    uint32_t a = w[g], b = w[4 + g], c = w[8 + g];
    ok &= red17(cc.qv[0][g] * a + cc.qv[1][g] * b + cc.qv[2][g] * c + cc.qv[3][g] * (a * b) + cc.qv[4][g]) == 0u;
  }
  out[i] = ok ? 1 : 0;
}

// constraints_satisfy (constraints.h:145-171) for a gate list of any length: q[5][rows] selectors (q_l q_r q_o q_m q_c),
// a / b / c [n][rows]; first_bad[i] = the first row whose gate equation fails, -1 if all hold (the reference prints that row)
__global__ void __launch_bounds__(BLOCK_LIGHT) satisfy_rows_kernel(const uint8_t* __restrict__ q, uint32_t rows, const uint8_t* __restrict__ a,
                                                                   const uint8_t* __restrict__ b, const uint8_t* __restrict__ c,
                                                                   int32_t* __restrict__ first_bad, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t bad = -1;
  for (uint32_t g = rows; g-- > 0;) {
    const uint32_t av = a[i * rows + g], bv = b[i * rows + g], cv = c[i * rows + g];
    const uint32_t lhs = q[g] * av + q[rows + g] * bv + q[2 * rows + g] * cv + q[3 * rows + g] * (av * bv) + q[4 * rows + g];
    if (red17(lhs) != 0u) bad = (int32_t)g;
  }
  first_bad[i] = bad;
}

// plonk_prove over a batch.  Shared memory: the per-lane-indexed tables, and a staging area through
// which the block's contiguous slice of every input and output array moves with 128-bit accesses
// (item records are 12 / 9 / 5 / 34 bytes -- per-thread byte accesses to global memory would cost one
// sector per byte on the store side).
template <int ITEM, int NT>
PB_D void stage_in(uint8_t* smem, const uint8_t* __restrict__ g, size_t first, size_t n) {
  // bytes [first*ITEM, min(first+NT, n)*ITEM) -> smem[0..]; the block base is 16-byte aligned because
  // first is a multiple of NT and NT*ITEM is a multiple of 16
  const uint8_t* src = g + first * ITEM;
  if (first + NT <= n) {   // a full block: compile-time trip count, 32-bit indices, no tail
    constexpr int NV = NT * ITEM / 16;
#pragma unroll
    for (int k0 = 0; k0 < NV; k0 += NT) {
      const int k = k0 + (int)threadIdx.x;
      if (k0 + NT <= NV || k < NV) reinterpret_cast<uint4*>(smem)[k] = reinterpret_cast<const uint4*>(src)[k];
    }
    return;
  }
  const size_t cnt = (n - first) * ITEM;
  const size_t nv = cnt / 16;
  for (size_t k = threadIdx.x; k < nv; k += NT) reinterpret_cast<uint4*>(smem)[k] = reinterpret_cast<const uint4*>(src)[k];
  for (size_t k = nv * 16 + threadIdx.x; k < cnt; k += NT) smem[k] = src[k];
}
template <int ITEM, int NT>
PB_D void stage_out(uint8_t* __restrict__ g, const uint8_t* smem, size_t first, size_t n) {
  uint8_t* dst = g + first * ITEM;
  if (first + NT <= n) {
    constexpr int NV = NT * ITEM / 16;
#pragma unroll
    for (int k0 = 0; k0 < NV; k0 += NT) {
      const int k = k0 + (int)threadIdx.x;
      if (k0 + NT <= NV || k < NV) reinterpret_cast<uint4*>(dst)[k] = reinterpret_cast<const uint4*>(smem)[k];
    }
    return;
  }
  const size_t cnt = (n - first) * ITEM;
  const size_t nv = cnt / 16;
  for (size_t k = threadIdx.x; k < nv; k += NT) reinterpret_cast<uint4*>(dst)[k] = reinterpret_cast<const uint4*>(smem)[k];
  for (size_t k = nv * 16 + threadIdx.x; k < cnt; k += NT) dst[k] = smem[k];
}

#ifndef PB_PROVE_BLOCK
#define PB_PROVE_BLOCK PB_BLOCK
#endif
constexpr int PBLOCK = PB_PROVE_BLOCK;   // threads per block of the prover (a multiple of 16: the staging copies move 16-byte pieces)
template <typename Tables>
struct __align__(16) ProveSmem {
  __align__(16) Tables tb;
  __align__(16) uint8_t wit[PBLOCK * 12];
  __align__(16) uint8_t rnd[PBLOCK * 9];
  __align__(16) uint8_t chal[PBLOCK * 5];
  __align__(16) uint8_t proof[PBLOCK * 34];
  __align__(16) uint8_t status[PBLOCK];
};

// Tables = ProverTables (any SRS, the reference's order of additions) or ProverPairTables (canonical on-curve SRS).
// done_list / done_count (optional): indices of the completed proofs (status 0) are appended, one atomic per warp,
// so that the verifier runs on a dense list; verdict (optional) gets 0xFF for every item that did not complete.
#ifndef PB_PROVE_PREFETCH
#define PB_PROVE_PREFETCH 370   // blocks ahead (half a resident wave); measured: 0 -> 238.4 us, 370 -> 234.1, 740 -> 235.0, 1480 -> 238.8
#endif
#ifndef PB_PROVE_MINBLOCKS
#define PB_PROVE_MINBLOCKS 5   // 96 registers, 5 blocks of 128 threads per SM: measured best for every table variant (profiles/r1/NOTES.md)
#endif
// FS = true (Fiat-Shamir mode, transcript.cuh): `chal` is not read; chal_out (optional) [n][6] receives the challenges
// drawn before the item's first exit (alpha beta gamma z v u), 0xFF for the ones not drawn.
// PACKED = true (packed wire v2, wire.cuh): `wit` holds the 16-byte packed input records, one 128-bit load per lane and no
// shared-memory staging of the inputs; `rnd` and `chal` are not read.
template <typename Tables, bool FS = false, bool PACKED = false>
__global__ void __launch_bounds__(PBLOCK, PB_PROVE_MINBLOCKS) prove_kernel(const __grid_constant__ CircuitConst cc, const Tables* __restrict__ gtb,
                                                      const uint8_t* __restrict__ wit, const uint8_t* __restrict__ rnd,
                                                      const uint8_t* __restrict__ chal, uint8_t* __restrict__ proofs,
                                                      uint8_t* __restrict__ status, size_t n, uint32_t* __restrict__ done_list,
                                                      uint32_t* __restrict__ done_count, uint8_t* __restrict__ verdict,
                                                      uint8_t* __restrict__ chal_out = nullptr, int wire3 = 0) {
  __shared__ ProveSmem<Tables> sm;
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * PBLOCK;
  static_assert(!(FS && PACKED), "the packed records carry the challenges");
  static_assert(sizeof(Tables) % 16 == 0, "bulk copies move multiples of 16 bytes");
#if PB_BULK
  // tables and (struct inputs of a full block) the three input slices: bulk copies issued by one thread, one mbarrier
  __shared__ __align__(8) uint64_t mbar;
  const bool bulk_in = !PACKED && first + PBLOCK <= n;
  if (tid == 0) mbar_init(&mbar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&mbar, (uint32_t)sizeof(Tables) + (bulk_in ? (uint32_t)PBLOCK * (12u + 9u + (FS ? 0u : 5u)) : 0u));
    bulk_load(&sm.tb, gtb, (uint32_t)sizeof(Tables), &mbar);
    if (bulk_in) {
      bulk_load(sm.wit, wit + first * 12, PBLOCK * 12, &mbar);
      bulk_load(sm.rnd, rnd + first * 9, PBLOCK * 9, &mbar);
      if constexpr (!FS) bulk_load(sm.chal, chal + first * 5, PBLOCK * 5, &mbar);
    }
  }
#else
  constexpr bool bulk_in = false;
  for (int k = tid; k < (int)(sizeof(Tables) / 4); k += PBLOCK)
    reinterpret_cast<uint32_t*>(&sm.tb)[k] = reinterpret_cast<const uint32_t*>(gtb)[k];
#endif
  const bool live = first + tid < n;
  uint32_t wa[4], wb[4], wc[4], r[9], ch[5];
  bool bad = false;
  if constexpr (PACKED) {
#if PB_PROVE_PREFETCH
    {
      const size_t pf = first + (size_t)PB_PROVE_PREFETCH * PBLOCK;   // 16 lines of 128 bytes per block
      if (pf + PBLOCK <= n && tid < (wire3 ? 14 : 16)) asm volatile("prefetch.global.L2 [%0];" ::"l"(wit + pf * (wire3 ? 14 : 16) + tid * 128));
    }
#endif
    uint32_t v[PACKED_VALUES];
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (wire3) {        // v3: 14-byte records, seven 16-bit loads (a warp's records are 448 contiguous bytes)
      uint16_t h[7] = {0, 0, 0, 0, 0, 0, 0};
      const uint16_t* rec = reinterpret_cast<const uint16_t*>(wit) + (first + tid) * 7;
      if (live) {
#pragma unroll
        for (int k = 0; k < 7; k++) h[k] = __ldg(rec + k);
      }
      uint32_t w[4];
      input14_words(h, w);
      q = make_uint4(w[0], w[1], w[2], w[3]);
    } else if (live) {
      q = reinterpret_cast<const uint4*>(wit)[first + tid];
    }
    bad = !unpack_input16(q.x, q.y, q.z, q.w, v);
#pragma unroll
    for (int k = 0; k < 4; k++) { wa[k] = v[k]; wb[k] = v[4 + k]; wc[k] = v[8 + k]; }
#pragma unroll
    for (int k = 0; k < 9; k++) r[k] = v[12 + k];
#pragma unroll
    for (int k = 0; k < 5; k++) ch[k] = v[21 + k];
#if PB_BULK
    mbar_wait(&mbar, 0);   // tables landed
#else
    __syncthreads();   // tables staged
#endif
  } else {
#if PB_PROVE_PREFETCH
  {
    // pull the inputs of the block that will take this block's slot next (~ one resident wave ahead) into L2, so that its
    // staging loads do not wait on DRAM: 26 lines of 128 bytes, one prefetch per lane
    const size_t pf = first + (size_t)PB_PROVE_PREFETCH * PBLOCK;
    if (pf + PBLOCK <= n && tid < 26) {
      const uint8_t* a = tid < 12 ? wit + pf * 12 + tid * 128 : tid < 21 ? rnd + pf * 9 + (tid - 12) * 128 : (FS ? nullptr : chal + pf * 5 + (tid - 21) * 128);
      if (a) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
  }
#endif
  if (!bulk_in) {      // the ragged last block (or PB_BULK=0): per-thread staging copies
    stage_in<12, PBLOCK>(sm.wit, wit, first, n);
    stage_in<9, PBLOCK>(sm.rnd, rnd, first, n);
    if constexpr (!FS) stage_in<5, PBLOCK>(sm.chal, chal, first, n);
    __syncthreads();
  }
#if PB_BULK
  mbar_wait(&mbar, 0);   // tables (and the bulk-copied inputs) landed
#endif
#pragma unroll
  for (int k = 0; k < 4; k++) { wa[k] = sm.wit[tid * 12 + k]; wb[k] = sm.wit[tid * 12 + 4 + k]; wc[k] = sm.wit[tid * 12 + 8 + k]; }
#pragma unroll
  for (int k = 0; k < 9; k++) r[k] = sm.rnd[tid * 9 + k];
#pragma unroll
  for (int k = 0; k < 5; k++) ch[k] = FS ? 0u : sm.chal[tid * 5 + k];
#pragma unroll
  for (int k = 0; k < 4; k++) bad |= wa[k] > 16u || wb[k] > 16u || wc[k] > 16u;
#pragma unroll
  for (int k = 0; k < 9; k++) bad |= r[k] > 16u;
#pragma unroll
  for (int k = 0; k < 5; k++) bad |= ch[k] > 16u;
  }
  if (bad || !live) {   // keep table indices in range; the item is reported as PB_PROVE_BAD_INPUT
#pragma unroll
    for (int k = 0; k < 4; k++) { wa[k] = 0; wb[k] = 0; wc[k] = 0; }
#pragma unroll
    for (int k = 0; k < 9; k++) r[k] = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) ch[k] = 0;
  }
  ProofOut o;
#ifndef PB_PROVE_DEFER_WZ
#define PB_PROVE_DEFER_WZ 1
#endif
  constexpr bool DEFER_WZ = PB_PROVE_DEFER_WZ && !FS && std::is_same<Tables, ProverWideTables>::value;
  prove_one<FS, DEFER_WZ>(cc, sm.tb, wa, wb, wc, r, ch[0], ch[1], ch[2], ch[3], ch[4], o);
  if (bad) o.status = 254u;
  if constexpr (FS) {
    if (chal_out && live) {
      // a challenge exists only if the reference's execution reaches the point where it is drawn
      const uint32_t st = o.status;
      const bool k1 = st == 0u || (st >= 6u && st <= 12u), k2 = st == 0u || (st >= 8u && st <= 12u);
      const bool k3 = st == 0u || (st >= 11u && st <= 12u), k5 = st == 0u;
      uint8_t* co = chal_out + (first + tid) * 6;
      co[0] = k2 ? (uint8_t)o.ch[0] : 0xFF;
      co[1] = k1 ? (uint8_t)o.ch[1] : 0xFF;
      co[2] = k1 ? (uint8_t)o.ch[2] : 0xFF;
      co[3] = k3 ? (uint8_t)o.ch[3] : 0xFF;
      co[4] = k3 ? (uint8_t)o.ch[4] : 0xFF;
      co[5] = k5 ? (uint8_t)o.ch[5] : 0xFF;
    }
  }
  const bool okp = o.status == 0u;     // a failed item's PROOF bytes are zero
  // dense list of the completed proofs: the atomic is issued first and its result used last, so that the round trip to L2
  // overlaps the writing of the record (it was 2.7 % of the kernel's stall samples when the warp waited for it on the spot)
  const bool done = live && okp;
  unsigned dmask = 0u;
  int leader = -1;
  uint32_t base = 0;
  if (done_list) {
    dmask = __ballot_sync(0xFFFFFFFFu, done);
    leader = __ffs(dmask) - 1;
    if (dmask && (tid & 31) == leader) base = atomicAdd(done_count, (uint32_t)__popc(dmask));
  }
  uint8_t* po = sm.proof + tid * 34;
#pragma unroll
  for (int j = 0; j < 9; j++) {
    if (DEFER_WZ && j == 7) continue;                     // [W_z] goes last: its gathers are still in flight
    po[3 * j] = okp ? (uint8_t)o.pts[j].x : 0;
    po[3 * j + 1] = okp ? (uint8_t)o.pts[j].y : 0;
    po[3 * j + 2] = okp ? (uint8_t)o.pts[j].inf : 0;
  }
#pragma unroll
  for (int j = 0; j < 7; j++) po[27 + j] = okp ? (uint8_t)o.sc[j] : 0;
  sm.status[tid] = (uint8_t)o.status;
  if constexpr (DEFER_WZ) {
    const G1 wz = wz_finish(sm.tb, o.wz);
    po[21] = okp ? (uint8_t)wz.x : 0;
    po[22] = okp ? (uint8_t)wz.y : 0;
    po[23] = okp ? (uint8_t)wz.inf : 0;
  }
  if (done_list) {
    base = __shfl_sync(0xFFFFFFFFu, base, leader < 0 ? 0 : leader);
    if (done) done_list[base + __popc(dmask & ((1u << (tid & 31)) - 1u))] = (uint32_t)(first + tid);
    if (live && !okp && verdict) verdict[first + tid] = 0xFF;
  }
#if PB_BULK
  if (first + PBLOCK <= n) {
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_store(proofs + first * 34, sm.proof, PBLOCK * 34);
      bulk_store(status + first, sm.status, PBLOCK);
      bulk_store_commit_wait();
    }
    return;
  }
#endif
  __syncthreads();
  stage_out<34, PBLOCK>(proofs, sm.proof, first, n);
  stage_out<1, PBLOCK>(status, sm.status, first, n);
}

// Fiat-Shamir: the six challenges (alpha beta gamma z v u) of each PROOF record, as a verifier derives them
__global__ void __launch_bounds__(BLOCK_LIGHT) fs_challenges_kernel(const __grid_constant__ FsState seed, const uint8_t* __restrict__ proofs,
                                                                    uint8_t* __restrict__ chal6, size_t n) {
  const size_t i = (size_t)blockIdx.x * BLOCK_LIGHT + threadIdx.x;
  if (i >= n) return;
  uint32_t pbytes[27], op[7], ch[5], u;
#pragma unroll
  for (int k = 0; k < 27; k++) pbytes[k] = proofs[i * 34 + k];
#pragma unroll
  for (int k = 0; k < 7; k++) op[k] = proofs[i * 34 + 27 + k];
  fs_derive(seed, pbytes, op, ch, u);
#pragma unroll
  for (int k = 0; k < 5; k++) chal6[i * 6 + k] = (uint8_t)ch[k];
  chal6[i * 6 + 5] = (uint8_t)u;
}

struct __align__(16) VerifySmem {
  __align__(16) FieldTables ft;
  __align__(16) uint8_t proof[BLOCK * 34];
  __align__(16) uint8_t chal[BLOCK * 5];
};

// status (optional): items whose status byte is non-zero are skipped and get verdict 0xFF
// the word of a packed input record that holds alpha beta gamma z v u: word 3 of a 16-byte v2 record, G of a 14-byte v3 record
PB_D uint32_t packed_tail_word(const uint32_t* __restrict__ packed, size_t item, int wire3) {
  if (!wire3) return __ldg(packed + item * 4 + 3);
  const uint16_t* h = reinterpret_cast<const uint16_t*>(packed) + item * 7;
  return input14_tail_word(__ldg(h + 1), __ldg(h + 3), __ldg(h + 5), __ldg(h + 6));
}
// packed (optional): the packed input records (wire.cuh); the challenges and u are then read from the item's record
// and `chal` / `u` are not read.
__global__ void __launch_bounds__(BLOCK) verify_kernel(const __grid_constant__ VerifyKey key, const uint8_t* __restrict__ proofs,
                                                       const uint8_t* __restrict__ chal, const uint8_t* __restrict__ u,
                                                       const uint8_t* __restrict__ status, uint8_t* __restrict__ verdict,
                                                       uint8_t* __restrict__ gt, size_t n, const uint32_t* __restrict__ packed = nullptr, int wire3 = 0) {
  __shared__ VerifySmem sm;
  const int tid = threadIdx.x;
  build_field_tables(sm.ft);
  const size_t first = (size_t)blockIdx.x * BLOCK;
  const bool fs = chal == nullptr && packed == nullptr;   // Fiat-Shamir mode: the challenges and u come from the proof bytes (transcript.cuh)
  stage_in<34, BLOCK>(sm.proof, proofs, first, n);
  if (chal) stage_in<5, BLOCK>(sm.chal, chal, first, n);
  __syncthreads();
  const size_t i = first + tid;
  if (i >= n) return;
  if (status && status[i] != 0) {
    verdict[i] = 0xFF;
    if (gt) reinterpret_cast<uint32_t*>(gt)[i] = 0u;
    return;
  }
  uint32_t pbytes[27], op[7], ch[5];
#pragma unroll
  for (int k = 0; k < 27; k++) pbytes[k] = sm.proof[tid * 34 + k];
#pragma unroll
  for (int k = 0; k < 7; k++) op[k] = sm.proof[tid * 34 + 27 + k];
  uint32_t uu;
  if (fs) {
    fs_derive(key.fs_seed, pbytes, op, ch, uu);
  } else if (packed) {
    uint32_t d[7];
    unpack7(packed_tail_word(packed, i, wire3), d);      // a non-canonical word never gets here: the prover reports the item as bad input
#pragma unroll
    for (int k = 0; k < 5; k++) ch[k] = d[k];
    uu = d[5];
  } else {
#pragma unroll
    for (int k = 0; k < 5; k++) ch[k] = sm.chal[tid * 5 + k];
    uu = u[i];
  }
  VerifyOut o;
  verify_one(key, sm.ft, pbytes, op, ch, uu, o);
  verdict[i] = (uint8_t)o.verdict;
  if (gt) reinterpret_cast<uint32_t*>(gt)[i] = o.lhs.a | (o.lhs.b << 8) | (o.rhs.a << 16) | (o.rhs.b << 24);
}

// Fast-path verifier (canonical on-curve key).  Two addressing modes: dense (item = global thread index, records
// staged through shared memory) or list (done_list/done_count from the prover: thread t handles item done_list[t],
// so that no lane idles on a proof that never completed).
struct __align__(16) VerifyFastSmem {
  __align__(16) FieldTables ft;
  __align__(16) VerifyTables vt;
  __align__(16) uint8_t proof[BLOCK * 34];
  __align__(16) uint8_t chal[BLOCK * 5];
#if PB_VERIFY_SMEM_PICK
  uint32_t pick[16 * BLOCK];          // per-thread sub-tables of the joint double-and-add, [pair][entry][thread]
#endif
};

#ifndef PB_VERIFY_MINBLOCKS
#define PB_VERIFY_MINBLOCKS 7   // 72 registers; measured 6: 150.4 us, 7: 149.2, default (64 registers): 150.6, 9: 152.0, 10: 154.0 (profiles/r2/NOTES.md)
#endif
template <bool WANT_GT>
__global__ void __launch_bounds__(BLOCK, PB_VERIFY_MINBLOCKS) verify_fast_kernel(const __grid_constant__ VerifyKey key, const VerifyTables* __restrict__ gvt,
                                                            const uint8_t* __restrict__ proofs, const uint8_t* __restrict__ chal,
                                                            const uint8_t* __restrict__ u, const uint32_t* __restrict__ done_list,
                                                            const uint32_t* __restrict__ done_count, uint8_t* __restrict__ verdict,
                                                            uint8_t* __restrict__ gt, size_t n, const uint32_t* __restrict__ packed = nullptr, int wire3 = 0) {
  __shared__ VerifyFastSmem sm;
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * BLOCK;
  const size_t limit = done_list ? (size_t)*done_count : n;
  if (first >= limit) return;                                             // whole block beyond the dense list
  static_assert(sizeof(VerifyTables) % 16 == 0 && sizeof(FieldTables) % 16 == 0, "bulk copies move multiples of 16 bytes");
#if PB_BULK
  // the 5 KB of tables arrive by two bulk copies issued by one thread (the per-thread loop was nine dependent L2 round trips)
  __shared__ __align__(8) uint64_t mbar;
  if (tid == 0) mbar_init(&mbar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&mbar, (uint32_t)(sizeof(VerifyTables) + sizeof(FieldTables)));
    bulk_load(&sm.vt, gvt, (uint32_t)sizeof(VerifyTables), &mbar);
    bulk_load(&sm.ft, &g_field_tables_image, (uint32_t)sizeof(FieldTables), &mbar);
  }
#else
  build_field_tables(sm.ft);
  for (int k = tid; k < (int)(sizeof(VerifyTables) / 4); k += BLOCK)
    reinterpret_cast<uint32_t*>(&sm.vt)[k] = reinterpret_cast<const uint32_t*>(gvt)[k];
#endif
  size_t item = first + tid;
  const bool live = item < limit;
  const bool fs = chal == nullptr && packed == nullptr;   // Fiat-Shamir mode: the challenges and u come from the proof bytes (transcript.cuh)
  uint32_t pbytes[27], op[7], ch[5];
  if (done_list) {
    // dense-list mode: the items of a block are scattered, so each lane reads its own 34-byte record straight into
    // registers (17 two-byte loads: records are 2-byte aligned) instead of staging it through shared memory
#if !PB_BULK
    __syncthreads();                                                      // tables staged
#endif
    if (!live) return;
    item = done_list[item];
    const uint16_t* pr = reinterpret_cast<const uint16_t*>(proofs + item * 34);
    uint32_t b[34];
#pragma unroll
    for (int k = 0; k < 17; k++) { const uint32_t w = pr[k]; b[2 * k] = w & 0xFFu; b[2 * k + 1] = w >> 8; }
#pragma unroll
    for (int k = 0; k < 27; k++) pbytes[k] = b[k];
#pragma unroll
    for (int k = 0; k < 7; k++) op[k] = b[27 + k];
    if (chal) {
#pragma unroll
      for (int k = 0; k < 5; k++) ch[k] = chal[item * 5 + k];
    }
#if PB_BULK
    mbar_wait(&mbar, 0);                                                  // tables landed (the record loads above are in flight meanwhile)
#endif
  } else {
    stage_in<34, BLOCK>(sm.proof, proofs, first, n);
    if (chal) stage_in<5, BLOCK>(sm.chal, chal, first, n);
    __syncthreads();
#if PB_BULK
    mbar_wait(&mbar, 0);
#endif
    if (!live) return;
#pragma unroll
    for (int k = 0; k < 27; k++) pbytes[k] = sm.proof[tid * 34 + k];
#pragma unroll
    for (int k = 0; k < 7; k++) op[k] = sm.proof[tid * 34 + 27 + k];
    if (chal) {
#pragma unroll
      for (int k = 0; k < 5; k++) ch[k] = sm.chal[tid * 5 + k];
    }
  }
  uint32_t uu;
  if (fs) {
    fs_derive(key.fs_seed, pbytes, op, ch, uu);
  } else if (packed) {
    uint32_t d[7];
    unpack7(packed_tail_word(packed, item, wire3), d);   // a non-canonical word never gets here: the prover reports the item as bad input
#pragma unroll
    for (int k = 0; k < 5; k++) ch[k] = d[k];
    uu = d[5];
  } else {
    uu = u[item];
  }
  VerifyOut o;
#if PB_VERIFY_SMEM_PICK
  verify_one_fast<WANT_GT>(key, sm.vt, sm.ft, pbytes, op, ch, uu, o, PickSmem<BLOCK>{sm.pick + tid});
#else
  verify_one_fast<WANT_GT>(key, sm.vt, sm.ft, pbytes, op, ch, uu, o);
#endif
  verdict[item] = (uint8_t)o.verdict;
  if constexpr (WANT_GT) reinterpret_cast<uint32_t*>(gt)[item] = o.lhs.a | (o.lhs.b << 8) | (o.rhs.a << 16) | (o.rhs.b << 24);
}

// Table-path verifier (verifier.cuh: verify_one_log): same addressing modes and the same record / challenge handling as
// verify_fast_kernel; the only shared memory is the 2.2 KB of tables (one bulk copy) and, in the dense mode, the staged records.
// threads (= items) per block of the table-path verifier; the status mode compacts per block (lane numbers are bytes: <= 256).
// Measured, status mode: 128: 53.3 us, 256: 57.4 us per 2^21 attempted items.
constexpr int VLBLOCK = 128;
static_assert(VLBLOCK <= 256, "lane_of holds lane numbers in bytes");
struct __align__(16) VerifyLogSmem {
  __align__(16) VerifyLogTables lt;
  __align__(16) uint8_t proof[VLBLOCK * 34];
  __align__(16) uint8_t chal[VLBLOCK * 5];
};
#ifndef PB_VERIFY_LOG_MINBLOCKS
#define PB_VERIFY_LOG_MINBLOCKS 8   // 60 registers; measured (us per 2^21 attempted items): no cap (91 registers) 55.2, 8: 46.9, 10: 47.2, 12: 51.3, 16: 62.2
#endif
// counts (optional, status mode only): the per-batch counters of pb_tally_dev -- status histogram, accepted proofs, byte sum
// of the proofs -- fall out of what this kernel already holds ("an epilogue block-reduce that produces the
// per-rank counters"): shared-memory atomics per lane, eighteen global atomics per block.
template <bool WANT_GT>
__global__ void __launch_bounds__(VLBLOCK, PB_VERIFY_LOG_MINBLOCKS * 128 / VLBLOCK) verify_log_kernel(const __grid_constant__ VerifyKey key, const VerifyLogTables* __restrict__ glt,
                                                           const uint8_t* __restrict__ proofs, const uint8_t* __restrict__ chal,
                                                           const uint8_t* __restrict__ u, const uint32_t* __restrict__ done_list,
                                                           const uint32_t* __restrict__ done_count, uint8_t* __restrict__ verdict,
                                                           uint8_t* __restrict__ gt, size_t n, const uint32_t* __restrict__ packed = nullptr, int wire3 = 0,
                                                           const uint8_t* __restrict__ status = nullptr, unsigned long long* __restrict__ counts = nullptr) {
  __shared__ VerifyLogSmem sm;
  __shared__ unsigned int tally[18];
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * VLBLOCK;
  const size_t limit = done_list ? (size_t)*done_count : n;
  if (first >= limit) return;                                             // whole block beyond the dense list
  static_assert(sizeof(VerifyLogTables) % 16 == 0, "bulk copies move multiples of 16 bytes");
  // status mode (status given, no list): the block compacts ITS OWN 128 items -- ranks of the completed ones by ballot and
  // a four-entry scan, their lane numbers into a 128-byte table -- and lane r takes the r-th of them.  No list in global
  // memory, no atomics in the prover; the lanes beyond the block's count idle (whole warps of them, mostly).
  const bool by_status = status != nullptr && done_list == nullptr;
  const bool tallying = counts != nullptr && by_status;
  if (tid < 18) tally[tid] = 0u;
#if PB_BULK
  __shared__ __align__(8) uint64_t mbar;
  if (tid == 0) mbar_init(&mbar, 1);
  __syncthreads();
  // status mode, full block: the block's 128 records (and challenge rows) are ONE contiguous 4352-byte (640-byte) slice -- it
  // arrives by bulk copy with the tables, and the compacted lanes read their records from shared memory: 34 wavefronts of
  // the load path per block where 17 strided two-byte loads per lane cost ~160 per warp
  const bool bulk_rec = by_status && first + VLBLOCK <= n;
  if (tid == 0) {
    mbar_arrive_expect_tx(&mbar, (uint32_t)sizeof(VerifyLogTables) + (bulk_rec ? (uint32_t)VLBLOCK * (34u + (chal ? 5u : 0u)) : 0u));
    bulk_load(&sm.lt, glt, (uint32_t)sizeof(VerifyLogTables), &mbar);
    if (bulk_rec) {
      bulk_load(sm.proof, proofs + first * 34, VLBLOCK * 34, &mbar);
      if (chal) bulk_load(sm.chal, chal + first * 5, VLBLOCK * 5, &mbar);
    }
  }
#else
  constexpr bool bulk_rec = false;
  for (int k = tid; k < (int)(sizeof(VerifyLogTables) / 4); k += VLBLOCK)
    reinterpret_cast<uint32_t*>(&sm.lt)[k] = reinterpret_cast<const uint32_t*>(glt)[k];
  __syncthreads();
#endif
  size_t item = first + tid;
  const bool live = item < limit;
  const bool fs = chal == nullptr && packed == nullptr;   // Fiat-Shamir mode: the challenges and u come from the proof bytes (transcript.cuh)
  uint32_t b[34], ch[5];
  bool active = live;
  if (by_status) {
    __shared__ uint8_t lane_of[VLBLOCK];
    __shared__ uint32_t wcnt[VLBLOCK / 32];
    const uint32_t st = live ? status[item] : 1u;
    const bool done = live && st == 0u;
    if (live && !done) verdict[item] = 0xFF;
    if (tallying && live) atomicAdd(&tally[st < 15u ? st : 15u], 1u);
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, done);
    if ((tid & 31) == 0) wcnt[tid >> 5] = (uint32_t)__popc(bal);
    __syncthreads();
    uint32_t rank = (uint32_t)__popc(bal & ((1u << (tid & 31)) - 1u)), cnt = 0;
#pragma unroll
    for (int w = 0; w < VLBLOCK / 32; w++) { if (w < (tid >> 5)) rank += wcnt[w]; cnt += wcnt[w]; }
    if (done) lane_of[rank] = (uint8_t)tid;
    __syncthreads();
    active = (uint32_t)tid < cnt;
    if (active) item = first + lane_of[tid];
  }
  if (bulk_rec) {
    if (active) {
#if PB_BULK
      mbar_wait(&mbar, 0);                                                // tables and the block's records landed
#endif
      const uint32_t src = (uint32_t)(item - first);
      const uint16_t* pr = reinterpret_cast<const uint16_t*>(sm.proof + src * 34u);
#pragma unroll
      for (int k = 0; k < 17; k++) { const uint32_t w = pr[k]; b[2 * k] = w & 0xFFu; b[2 * k + 1] = w >> 8; }
      if (chal) {
#pragma unroll
        for (int k = 0; k < 5; k++) ch[k] = sm.chal[src * 5u + k];
      }
    }
  } else if (done_list || by_status) {
    // dense-list mode (or the ragged last block of the status mode): the items of a block are scattered, so each lane reads
    // its own 34-byte record straight into registers (17 two-byte loads: records are 2-byte aligned)
    if (active) {
      if (done_list) item = done_list[item];
      const uint16_t* pr = reinterpret_cast<const uint16_t*>(proofs + item * 34);
#pragma unroll
      for (int k = 0; k < 17; k++) { const uint32_t w = pr[k]; b[2 * k] = w & 0xFFu; b[2 * k + 1] = w >> 8; }
      if (chal) {
#pragma unroll
        for (int k = 0; k < 5; k++) ch[k] = chal[item * 5 + k];
      }
#if PB_BULK
      mbar_wait(&mbar, 0);                                                // tables landed (the record loads above were in flight meanwhile)
#endif
    }
  } else {
    stage_in<34, VLBLOCK>(sm.proof, proofs, first, n);
    if (chal) stage_in<5, VLBLOCK>(sm.chal, chal, first, n);
    __syncthreads();
    if (active) {
#if PB_BULK
      mbar_wait(&mbar, 0);
#endif
#pragma unroll
      for (int k = 0; k < 34; k++) b[k] = sm.proof[tid * 34 + k];
      if (chal) {
#pragma unroll
        for (int k = 0; k < 5; k++) ch[k] = sm.chal[tid * 5 + k];
      }
    }
  }
  if (active) {
    uint32_t pbytes[27], op[7];
#pragma unroll
    for (int k = 0; k < 27; k++) pbytes[k] = b[k];
#pragma unroll
    for (int k = 0; k < 7; k++) op[k] = b[27 + k];
    uint32_t uu;
    if (fs) {
      fs_derive(key.fs_seed, pbytes, op, ch, uu);
    } else if (packed) {
      uint32_t d[7];
      unpack7(packed_tail_word(packed, item, wire3), d);   // a non-canonical word never gets here: the prover reports the item as bad input
#pragma unroll
      for (int k = 0; k < 5; k++) ch[k] = d[k];
      uu = d[5];
    } else {
      uu = u[item];
    }
    VerifyOut o;
    verify_one_log<WANT_GT>(sm.lt, pbytes, op, ch, uu, o);
    verdict[item] = (uint8_t)o.verdict;
    if constexpr (WANT_GT) reinterpret_cast<uint32_t*>(gt)[item] = o.lhs.a | (o.lhs.b << 8) | (o.rhs.a << 16) | (o.rhs.b << 24);
    if (tallying) {
      uint32_t bsum = 0u;
#pragma unroll
      for (int k = 0; k < 34; k++) bsum += b[k];
      atomicAdd(&tally[17], bsum);                                        // <= 128 * 34 * 255 per block
      if (o.verdict == 1u) atomicAdd(&tally[16], 1u);
    }
  }
  if (tallying) {                                                          // uniform: every lane of the block gets here
    __syncthreads();
    if (tid < 18 && tally[tid] != 0u) atomicAdd(counts + tid, (unsigned long long)tally[tid]);
  }
}

// VerifyLogTables at context creation: the sequential part by one thread, then one entry per thread.  ok[0] = 0 if the
// construction failed (the context then keeps the Straus kernel).
__global__ void __launch_bounds__(256) verify_log_tables_kernel(const __grid_constant__ VerifyKey key, const VerifyTables* __restrict__ vt,
                                                                VerifyLogTables* __restrict__ out, uint32_t* __restrict__ ok) {
  __shared__ FieldTables ft;
  __shared__ VerifyLogTables lt;
  __shared__ uint8_t alog[104];
  __shared__ uint32_t good;
  build_field_tables(ft);
  __syncthreads();
  if (threadIdx.x == 0) good = vlt_group(ft, lt, alog) ? 1u : 0u;
  __syncthreads();
  if (good)
    for (uint32_t t = threadIdx.x; t < VLT_ENTRIES; t += blockDim.x) vlt_entry(ft, key, *vt, alog, lt, t);
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < sizeof(VerifyLogTables) / 4; k += blockDim.x)
    reinterpret_cast<uint32_t*>(out)[k] = reinterpret_cast<const uint32_t*>(&lt)[k];
  if (threadIdx.x == 0) *ok = good;
}

// status bytes -> dense list of the completed items (for pb_plonk_verify_completed_dev, where the list does not come
// from the prover); verdict gets 0xFF elsewhere
__global__ void __launch_bounds__(BLOCK_LIGHT) compact_kernel(const uint8_t* __restrict__ status, size_t n, uint32_t* __restrict__ done_list,
                                                              uint32_t* __restrict__ done_count, uint8_t* __restrict__ verdict) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const bool done = live && status[i] == 0;
  const unsigned m = __ballot_sync(0xFFFFFFFFu, done);
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (m && lane == leader) base = atomicAdd(done_count, (uint32_t)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader < 0 ? 0 : leader);
  if (done) done_list[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
  if (live && !done) verdict[i] = 0xFF;
}

// ------------------------------------------------------------------ compact / packed outputs, synthetic inputs (wire.cuh)
// Only the proofs that exist travel back to the host: a dense array of the completed proofs (status 0) in item order.
// Step 1: offs[g] = number of completed items before group g (GBLOCK items per group), *total = all of them.
constexpr int GBLOCK = 128;
// Step 1a (all SMs): offs[g] = number of completed items of group g; one group of GBLOCK status bytes per thread.
__global__ void __launch_bounds__(256) done_counts_kernel(const uint8_t* __restrict__ status, size_t m, uint32_t* __restrict__ offs) {
  const uint32_t G = (uint32_t)((m + GBLOCK - 1) / GBLOCK);
  const uint32_t g = blockIdx.x * 256u + threadIdx.x;
  if (g >= G) return;
  const bool al = (reinterpret_cast<uintptr_t>(status) & 15u) == 0;
  const size_t lo = (size_t)g * GBLOCK, hi = lo + GBLOCK < m ? lo + GBLOCK : m;
  uint32_t c = 0;
  if (al && hi - lo == GBLOCK) {
#pragma unroll
    for (int k = 0; k < GBLOCK / 16; k++) {
      const uint4 q = reinterpret_cast<const uint4*>(status + lo)[k];
      c += (uint32_t)(__popc(__vcmpeq4(q.x, 0u)) + __popc(__vcmpeq4(q.y, 0u)) + __popc(__vcmpeq4(q.z, 0u)) + __popc(__vcmpeq4(q.w, 0u))) >> 3;
    }
  } else {
    for (size_t i = lo; i < hi; i++) c += status[i] == 0 ? 1u : 0u;
  }
  offs[g] = c;
}
// Step 1b (one block): exclusive scan of the group counts in place, *total = their sum.
// total_host (optional): a second copy of the count in mapped pinned host memory -- the host-pointer pipelines read it after
// the chunk's event instead of queueing a 4-byte D2H copy behind megabytes of proofs on the copy engine.
__global__ void __launch_bounds__(1024) done_offsets_kernel(size_t m, uint32_t* __restrict__ offs, uint32_t* __restrict__ total,
                                                            uint32_t* __restrict__ total_host = nullptr) {
  __shared__ uint32_t wsum[32];
  const uint32_t G = (uint32_t)((m + GBLOCK - 1) / GBLOCK);
  // thread t owns the contiguous groups [t * per, (t + 1) * per)
  const uint32_t per = (G + 1023u) / 1024u;
  const uint32_t lo = threadIdx.x * per < G ? threadIdx.x * per : G, hi = lo + per < G ? lo + per : G;
  uint32_t mine = 0;
  for (uint32_t g = lo; g < hi; g++) mine += offs[g];
  uint32_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((threadIdx.x & 31) >= d) incl += t; }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t w = wsum[threadIdx.x], wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, d); if ((int)threadIdx.x >= d) wi += t; }
    wsum[threadIdx.x] = wi - w;
    if (threadIdx.x == 31) { *total = wi; if (total_host) *total_host = wi; }
  }
  __syncthreads();
  uint32_t run = wsum[threadIdx.x >> 5] + incl - mine;
  for (uint32_t g = lo; g < hi; g++) { const uint32_t c = offs[g]; offs[g] = run; run += c; }
}

// Step 2: block g moves its completed PROOF records to dense[offs[g]...], as 34-byte structs (PACK = 0), as 22-byte
// packed v2 records (PACK = 1) or as 12-byte v3 records (PACK = 2; the points must be on the curve, wire.cuh); with PACK != 0
// sv[i] = status/verdict nibbles is written for every item as well.  The records of a
// block form one contiguous run of the output; it is assembled in shared memory at the same 16-byte phase as its
// destination, so that the body goes out in 128-bit stores.  dense must be 16-byte aligned.
template <int PACK>
__global__ void __launch_bounds__(GBLOCK) gather_done_kernel(const uint8_t* __restrict__ proofs, const uint8_t* __restrict__ status,
                                                             const uint8_t* __restrict__ verdict, size_t m, const uint32_t* __restrict__ offs,
                                                             uint8_t* __restrict__ dense, uint8_t* __restrict__ sv) {
  constexpr int REC = PACK == 2 ? PACKED3_PROOF_BYTES : PACK == 1 ? PACKED_PROOF_BYTES : 34;
  __shared__ __align__(16) uint8_t src[GBLOCK * 34];
  __shared__ __align__(16) uint8_t cbase[104];
  if (PACK == 2 && threadIdx.x < 26) reinterpret_cast<uint32_t*>(cbase)[threadIdx.x] = reinterpret_cast<const uint32_t*>(g_curve_index.base)[threadIdx.x];
  __shared__ __align__(16) uint8_t dst[GBLOCK * REC + 16];
  __shared__ uint32_t wcnt[GBLOCK / 32];
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * GBLOCK;
  const bool live = first + tid < m;
  stage_in<34, GBLOCK>(src, proofs, first, m);
  const uint32_t st = live ? status[first + tid] : 1u;
  if (PACK && live) sv[first + tid] = (uint8_t)sv_byte(st, verdict ? verdict[first + tid] : 0xFFu);
  const bool done = live && st == 0u;
  const unsigned bal = __ballot_sync(0xFFFFFFFFu, done);
  if ((tid & 31) == 0) wcnt[tid >> 5] = (uint32_t)__popc(bal);
  __syncthreads();
  uint32_t rank = (uint32_t)__popc(bal & ((1u << (tid & 31)) - 1u)), cnt = 0;
#pragma unroll
  for (int w = 0; w < GBLOCK / 32; w++) { if (w < (tid >> 5)) rank += wcnt[w]; cnt += wcnt[w]; }
  const size_t g0 = (size_t)offs[blockIdx.x] * REC;
  const uint32_t skew = (uint32_t)((reinterpret_cast<uintptr_t>(dense) + g0) & 15u);   // even: REC is even
  if (done) {
    uint16_t* o = reinterpret_cast<uint16_t*>(dst + skew + rank * REC);
    if (PACK == 2) {
      uint32_t rec[3];
      pack_proof12(src + tid * 34, cbase, rec);
#pragma unroll
      for (int k = 0; k < 3; k++) { o[2 * k] = (uint16_t)(rec[k] & 0xFFFFu); o[2 * k + 1] = (uint16_t)(rec[k] >> 16); }
    } else if (PACK == 1) {
      uint16_t rec[11];
      pack_proof22(src + tid * 34, rec);
#pragma unroll
      for (int k = 0; k < 11; k++) o[k] = rec[k];
    } else {
      const uint16_t* i16 = reinterpret_cast<const uint16_t*>(src + tid * 34);
#pragma unroll
      for (int k = 0; k < 17; k++) o[k] = i16[k];
    }
  }
  __syncthreads();
  const uint32_t total = cnt * REC;
  uint32_t head = (16u - skew) & 15u;
  if (head > total) head = total;
  const uint32_t body = (total - head) & ~15u;
  uint8_t* out = dense + g0;
  const uint8_t* in = dst + skew;
  for (uint32_t k = 2u * tid; k < head; k += 2u * GBLOCK) *reinterpret_cast<uint16_t*>(out + k) = *reinterpret_cast<const uint16_t*>(in + k);
  for (uint32_t k = tid; k < body / 16u; k += GBLOCK) reinterpret_cast<uint4*>(out + head)[k] = reinterpret_cast<const uint4*>(in + head)[k];
  for (uint32_t k = head + body + 2u * tid; k < total; k += 2u * GBLOCK) *reinterpret_cast<uint16_t*>(out + k) = *reinterpret_cast<const uint16_t*>(in + k);
}

// items [start, start + n) of the synthetic stream `seed` (wire.cuh: synth_item = workload.py make_batch) as packed input
// records: one 128-bit store per lane.  wtab: the 289-row witness table [289][12].
__global__ void __launch_bounds__(BLOCK_LIGHT) synth_packed_kernel(unsigned long long seed, unsigned long long start, int variant,
                                                                   const uint8_t* __restrict__ wtab, uint8_t* __restrict__ packed, size_t n) {
  __shared__ uint8_t tab[SYNTH_WITNESS_ROWS * 12];
  for (int k = threadIdx.x; k < SYNTH_WITNESS_ROWS * 12; k += BLOCK_LIGHT) tab[k] = wtab[k];
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * BLOCK_LIGHT + threadIdx.x;
  if (i >= n) return;
  uint32_t v[PACKED_VALUES], w[4];
  synth_item(seed, start + i, variant, tab, v);
  pack_input16(v, w);
  reinterpret_cast<uint4*>(packed)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}
// the same items as the reference's structs: witness[n][12], rnd[n][9], chal[n][5], u[n]
__global__ void __launch_bounds__(BLOCK_LIGHT) synth_struct_kernel(unsigned long long seed, unsigned long long start, int variant,
                                                                   const uint8_t* __restrict__ wtab, uint8_t* __restrict__ wit,
                                                                   uint8_t* __restrict__ rnd, uint8_t* __restrict__ chal, uint8_t* __restrict__ u, size_t n) {
  __shared__ uint8_t tab[SYNTH_WITNESS_ROWS * 12];
  for (int k = threadIdx.x; k < SYNTH_WITNESS_ROWS * 12; k += BLOCK_LIGHT) tab[k] = wtab[k];
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * BLOCK_LIGHT + threadIdx.x;
  if (i >= n) return;
  uint32_t v[PACKED_VALUES];
  synth_item(seed, start + i, variant, tab, v);
  for (int k = 0; k < 12; k++) wit[i * 12 + k] = (uint8_t)v[k];
  for (int k = 0; k < 9; k++) rnd[i * 9 + k] = (uint8_t)v[12 + k];
  for (int k = 0; k < 5; k++) chal[i * 5 + k] = (uint8_t)v[21 + k];
  u[i] = (uint8_t)v[26];
}
// struct arrays <-> packed records on the device (format conversion for callers that hold one and want the other)
__global__ void __launch_bounds__(BLOCK_LIGHT) pack_inputs_kernel(const uint8_t* __restrict__ wit, const uint8_t* __restrict__ rnd,
                                                                  const uint8_t* __restrict__ chal, const uint8_t* __restrict__ u,
                                                                  uint8_t* __restrict__ packed, size_t n) {
  const size_t i = (size_t)blockIdx.x * BLOCK_LIGHT + threadIdx.x;
  if (i >= n) return;
  uint32_t v[PACKED_VALUES], w[4];
  for (int k = 0; k < 12; k++) v[k] = wit[i * 12 + k];
  for (int k = 0; k < 9; k++) v[12 + k] = rnd[i * 9 + k];
  for (int k = 0; k < 5; k++) v[21 + k] = chal[i * 5 + k];
  v[26] = u[i];
  bool ok = true;
  for (int k = 0; k < PACKED_VALUES; k++) ok &= v[k] < 17u;
  pack_input16(v, w);
  if (!ok) w[0] = w[1] = w[2] = w[3] = 0xFFFFFFFFu;     // not an encoding: the prover reports PB_PROVE_BAD_INPUT, as it does for the struct bytes
  reinterpret_cast<uint4*>(packed)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// pair tables from single-point rows: out[j][c0*17 + c1] = g1_add(T[2j][c0], T[2j+1][c1]) (row beyond `rows`: identity)
__global__ void pair_table_kernel(const uint32_t* __restrict__ T, uint32_t rows, uint32_t pairs, uint32_t* __restrict__ out) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < pairs * 289u; k += gridDim.x * blockDim.x) {
    const uint32_t j = k / 289u, c0 = (k % 289u) / 17u, c1 = k % 17u;
    const G1 p = 2 * j < rows ? unpack_g1(T[(2 * j) * 17 + c0]) : g1_identity();
    const G1 q = 2 * j + 1 < rows ? unpack_g1(T[(2 * j + 1) * 17 + c1]) : g1_identity();
    const G1 r = g1_add(ft, p, q);
    out[k] = pack_g1(r.x, r.y, r.inf);
  }
}

// wide tables, step 1: T3[t][c0 + 17 c1 + 289 c2] = (T[3t][c0] + T[3t+1][c1]) + T[3t+2][c2], t = 0..2, 16-bit packed
__global__ void wide_t3_kernel(const uint32_t* __restrict__ T, uint32_t rows, uint16_t* __restrict__ out /*[3][17^3]*/) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < 3u * WIDE_T3_ENTRIES; k += gridDim.x * blockDim.x) {
    const uint32_t t = k / WIDE_T3_ENTRIES, e = k % WIDE_T3_ENTRIES;
    const uint32_t c[3] = {e % 17u, (e / 17u) % 17u, e / 289u};
    G1 acc = g1_identity();
#pragma unroll
    for (uint32_t i = 0; i < 3; i++) {
      const uint32_t row = 3u * t + i;
      acc = g1_add(ft, acc, row < rows ? unpack_g1(T[row * 17u + c[i]]) : g1_identity());
    }
    out[k] = (uint16_t)pack_g1_16(acc);
  }
}
// step 2: T6[i + 17^3 j] = T3[0][i] + T3[1][j]
__global__ void wide_t6_kernel(const uint16_t* __restrict__ T3, uint16_t* __restrict__ T6) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < WIDE_T6_ENTRIES; k += gridDim.x * blockDim.x) {
    const uint32_t i = k % WIDE_T3_ENTRIES, j = k / WIDE_T3_ENTRIES;
    const G1 r = g1_add(ft, unpack_g1_16(T3[i]), unpack_g1_16(T3[WIDE_T3_ENTRIES + j]));
    T6[k] = (uint16_t)pack_g1_16(r);
  }
}

// VerifyTables from the single-point rows of the nine key points, in VerifyKey order qm ql qr qo qc s1 s2 s3 one
__global__ void verify_tables_kernel(const uint32_t* __restrict__ KT /*[9][17]*/, VerifyTables* __restrict__ out) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < 4u * 289u + 17u; k += blockDim.x) {
    G1 r;
    if (k < 4u * 289u) {
      const uint32_t t = k / 289u, a = (k % 289u) / 17u, b = k % 17u;
      const int ia[4] = {0, 2, 4, 5}, ib[4] = {1, 3, 7, 6};            // (qm,ql) (qr,qo) (qc,s3) (s1,s2)
      G1 p = unpack_g1(KT[ia[t] * 17 + a]), q = unpack_g1(KT[ib[t] * 17 + b]);
      if (t == 2) q = g1_neg(q);                                        // - d_s3 [S3]
      r = g1_add(ft, p, q);
      out->P2[t][k % 289u] = pack_g1(r.x, r.y, r.inf);
    } else {
      r = g1_neg(unpack_g1(KT[8 * 17 + (k - 4u * 289u)]));              // - e [1]
      out->one_neg[k - 4u * 289u] = pack_g1(r.x, r.y, r.inf);
    }
  }
}

// the eight preprocessed commitments of the verifier key: srs_eval_at_s of the interpolated selector and
// permutation polynomials (context creation; one thread each)
__global__ void verifier_key_kernel(const uint32_t* __restrict__ Tg, uint32_t srs_len, const uint8_t* __restrict__ polys /*[8][4]*/,
                                    uint32_t* __restrict__ out /*[8] packed, 0xFFFFFFFF = longer than the SRS*/) {
  __shared__ FieldTables ft;
  build_field_tables(ft);
  __syncthreads();
  if (threadIdx.x >= 8) return;
  const uint8_t* row = polys + 4 * threadIdx.x;
  int len = 4;
  while (len > 1 && row[len - 1] == 0) len--;
  if ((uint32_t)len > srs_len) { out[threadIdx.x] = 0xFFFFFFFFu; return; }
  G1 acc = g1_identity();
  for (int k = 0; k < len; k++) acc = g1_add(ft, acc, unpack_g1(Tg[k * 17 + row[k]]));
  out[threadIdx.x] = pack_g1(acc.x, acc.y, acc.inf);
}

// counts[s] += #status==s (s < 15; 15 = other), counts[16] += #verdict==1, counts[17] += sum of proof bytes
// Per-batch counters: status histogram, accepted count, byte sum of the proofs.  HBM-bound by design:
// 36 B per item are read once, 16 bytes per request; the counting is SIMD-in-a-word (__vcmpeq4 / __vsadu4) in registers,
// one warp reduction and one shared-memory atomic per warp and counter, one global atomic per block and counter.
// vec = 0: the pointers are not 16-byte aligned, everything goes through the byte loops.
__global__ void __launch_bounds__(BLOCK_LIGHT) tally_kernel(const uint8_t* __restrict__ proofs, const uint8_t* __restrict__ status,
                                                            const uint8_t* __restrict__ verdict, size_t n, unsigned long long* __restrict__ counts,
                                                            int vec) {
  __shared__ unsigned long long sc[18];
  if (threadIdx.x < 18) sc[threadIdx.x] = 0ull;
  __syncthreads();
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  uint32_t hist[16];
#pragma unroll
  for (int b = 0; b < 16; b++) hist[b] = 0u;
  uint32_t accepted = 0u;
  unsigned long long sum = 0ull;
  const size_t nvec = vec ? n / 16 : 0;           // items covered by 16-byte pieces of status / verdict
  for (size_t k = gtid; k < nvec; k += stride) {
    if (status) {
      const uint4 q = reinterpret_cast<const uint4*>(status)[k];
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      uint32_t low = 0u;
#pragma unroll
      for (int j = 0; j < 4; j++) {
#pragma unroll
        for (uint32_t b = 0; b < 15u; b++) {
          const uint32_t c = (uint32_t)__popc(__vcmpeq4(w[j], b * 0x01010101u)) >> 3;
          hist[b] += c;
          low += c;
        }
      }
      hist[15] += 16u - low;                      // every status byte outside 0..14
    }
    if (verdict) {
      const uint4 q = reinterpret_cast<const uint4*>(verdict)[k];
      accepted += (uint32_t)(__popc(__vcmpeq4(q.x, 0x01010101u)) + __popc(__vcmpeq4(q.y, 0x01010101u)) +
                             __popc(__vcmpeq4(q.z, 0x01010101u)) + __popc(__vcmpeq4(q.w, 0x01010101u))) >> 3;
    }
  }
  for (size_t i = nvec * 16 + gtid; i < n; i += stride) {      // ragged tail, or everything when unaligned
    const uint32_t s = status ? status[i] : 0u;
    const uint32_t bin = s < 15u ? s : 15u;
#pragma unroll
    for (uint32_t b = 0; b < 16u; b++) hist[b] += bin == b ? 1u : 0u;
    if (verdict && verdict[i] == 1) accepted++;
  }
  if (!status && vec) {                            // no status array: every item counts as status 0 (as the byte loop does)
    for (size_t k = gtid; k < nvec; k += stride) hist[0] += 16u;
  }
  if (proofs) {
    const size_t bytes = n * 34;
    const size_t nv = vec ? bytes / 16 : 0;
    uint32_t part = 0u;
    size_t k = gtid;
    for (; k + 3 * stride < nv; k += 4 * stride) {          // four independent 16-byte loads in flight per lane
      uint4 q[4];
#pragma unroll
      for (int j = 0; j < 4; j++) q[j] = reinterpret_cast<const uint4*>(proofs)[k + j * stride];
#pragma unroll
      for (int j = 0; j < 4; j++) part += __vsadu4(q[j].x, 0u) + __vsadu4(q[j].y, 0u) + __vsadu4(q[j].z, 0u) + __vsadu4(q[j].w, 0u);
      if ((part >> 30) != 0u) { sum += part; part = 0u; }   // <= 16320 per round
    }
    for (; k < nv; k += stride) {
      const uint4 q = reinterpret_cast<const uint4*>(proofs)[k];
      part += __vsadu4(q.x, 0u) + __vsadu4(q.y, 0u) + __vsadu4(q.z, 0u) + __vsadu4(q.w, 0u);   // <= 4080 per piece
      if ((part >> 30) != 0u) { sum += part; part = 0u; }
    }
    sum += part;
    for (size_t k = nv * 16 + gtid; k < bytes; k += stride) sum += proofs[k];
  }
  // warp reduction, then one shared atomic per warp and counter
#pragma unroll
  for (int b = 0; b < 16; b++) {
    const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, hist[b]);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(&sc[b], (unsigned long long)t);
  }
  {
    const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, accepted);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(&sc[16], (unsigned long long)t);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xFFFFFFFFu, sum, off);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&sc[17], sum);
  }
  __syncthreads();
  if (threadIdx.x < 18 && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sc[threadIdx.x]);
}

// ------------------------------------------------------------------ roofline probes
// Dependency-free instruction streams (16 independent chains per thread) to measure the issue-rate ceilings the
// integer kernels are charged against.  PROBE_OPS_PER_ITER counted thread-level instructions per loop iteration
// (one per chain per repetition; the SASS of every kind is checked with cuobjdump to be exactly that instruction).
// kind 0 IMAD (32-bit)   1 LOP3 (alu pipe)   2 IMAD and LOP3 alternating   3 LDS.U8 (+ address math, counted as 1)
// kind 4 FFMA            5 HFMA2             6 IDP4A (dp4a)                 7 IMAD.WIDE (32x32+64)
constexpr int PROBE_OPS_PER_ITER = 64;
__global__ void __launch_bounds__(256) peak_probe_kernel(int kind, uint32_t iters, uint32_t* __restrict__ sink) {
  __shared__ uint8_t lut[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) lut[i] = (uint8_t)((i * 7 + 3) & 0xFF);
  __syncthreads();
  uint32_t r[16];
#pragma unroll
  for (int k = 0; k < 16; k++) r[k] = threadIdx.x * 16u + k + blockIdx.x;
  const uint32_t m = 2654435761u + blockIdx.x, c = 40503u + threadIdx.x;
  uint32_t x = 0;
  if (kind == 0) {
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int k = 0; k < 16; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[k]) : "r"(m), "r"(c));
    }
  } else if (kind == 1) {
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int k = 0; k < 16; k++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[k]) : "r"(m), "r"(c));
    }
  } else if (kind == 2) {
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int rep = 0; rep < 2; rep++)
#pragma unroll
        for (int k = 0; k < 16; k++) {
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[k]) : "r"(m), "r"(c));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(k + 8) & 15]) : "r"(m), "r"(c));
        }
    }
  } else if (kind == 3) {
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int k = 0; k < 16; k++) r[k] = lut[(r[k] + k * 61u) & 1023u] + (r[k] >> 3);
    }
  } else if (kind == 4) {
    float f[16];
#pragma unroll
    for (int k = 0; k < 16; k++) f[k] = (float)r[k];
    const float fm = 1.0000001f, fc = 0.5f;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int k = 0; k < 16; k++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fm), "f"(fc));
    }
#pragma unroll
    for (int k = 0; k < 16; k++) x ^= __float_as_uint(f[k]);
  } else if (kind == 5) {
    const uint32_t hm = 0x3C003C00u, hc = 0x38003800u;     // half2(1, 1), half2(0.5, 0.5)
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int k = 0; k < 16; k++) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[k]) : "r"(hm), "r"(hc));
    }
  } else if (kind == 6) {
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int k = 0; k < 16; k++) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(r[k]) : "r"(m), "r"(c));
    }
  } else {
    unsigned long long w[16];
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = r[k];
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int k = 0; k < 16; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(m), "r"(c));
    }
#pragma unroll
    for (int k = 0; k < 16; k++) x ^= (uint32_t)(w[k] ^ (w[k] >> 32));
  }
#pragma unroll
  for (int k = 0; k < 16; k++) x ^= r[k];
  if (x == 0xDEADBEEFu) sink[0] = x;
}

}  // namespace pb
