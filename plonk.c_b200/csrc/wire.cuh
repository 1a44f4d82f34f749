// wire.cuh -- packed wire formats v2 / v3 and the counter-based synthetic stream, as device functions.
//
// Why: end to end the path is bound by the PCIe link, not by the kernels: the reference's structs cost
// 27 B in and 36 B out per proof.  The packed format carries the same information in 16 B in and 22 B per COMPLETED
// proof (+ 1 B per item) out; a proof the reference never produces (it exit()s) does not travel.
//
//   packed input record, 16 bytes = four little-endian u32 words; word k holds seven base-17 digits, least significant
//   first, of the 27 values  a[4] b[4] c[4] (ASSIGNMENTS, constraints.h:57-62) | rand[9] (plonk.h:228) |
//   alpha beta gamma z v (CHALLENGE, plonk.h:16-22) | u (the verifier's batching scalar) | one spare digit (0):
//   word k = sum_j v[7k + j] * 17^j < 17^7 = 410 338 673 < 2^32.  Any other bit pattern (word >= 17^7, spare != 0) is
//   not an encoding and the item is reported as PB_PROVE_BAD_INPUT.
//
//   packed proof record, 22 bytes: nine little-endian u16 points x | y << 7 | infinite << 14 in PROOF field order
//   (plonk.h:24-41; coordinates are < 101 < 2^7), then one little-endian u32 holding the seven openings as base-17
//   digits.  Records are dense: only items with status 0 have one, in item order.
//
//   sv byte, one per item: low nibble = prove status (15 = PB_PROVE_BAD_INPUT), high nibble =
//   verdict (0 reject, 1 accept, 2 bad point, 3 bad scalar, 15 = not verified because the proof does not exist).
//
// Host-side twins of these functions (pb_wire_pack_inputs / pb_wire_unpack_proofs in cabi.cu, plonk.c_b200/wire.py) are
// format conversion only; the tests check all three against each other.
#pragma once
#include "field.cuh"

namespace pb {

constexpr uint32_t P17_7 = 410338673u;   // 17^7
constexpr int PACKED_IN_BYTES = 16;
constexpr int PACKED_PROOF_BYTES = 22;
constexpr int PACKED_VALUES = 27;        // a[4] b[4] c[4] rand[9] chal[5] u

// seven base-17 digits of a 32-bit word, least significant first.  floor(w / 17) for any 32-bit w is
// mulhi(w, 0xF0F0F0F1) >> 4; every later quotient is < 2^28, where the plain Barrett step of red17 is exact.
// Returns false when w is not a canonical encoding (w >= 17^7  <=>  the last quotient is >= 17).
PB_HD bool unpack7(uint32_t w, uint32_t (&d)[7]) {
  uint32_t q = mulhi_u32(w, 0xF0F0F0F1u) >> 4;
  d[0] = w - 17u * q;
#pragma unroll
  for (int k = 1; k < 6; k++) {
    const uint32_t q2 = mulhi_u32(q, M17);
    d[k] = q - 17u * q2;
    q = q2;
  }
  d[6] = q;
  return q < 17u;
}
PB_HD uint32_t pack7(const uint32_t (&d)[7]) {
  uint32_t w = d[6];
#pragma unroll
  for (int k = 5; k >= 0; k--) w = w * 17u + d[k];
  return w;
}

// 16-byte input record -> the 27 values; false = not a canonical encoding
PB_HD bool unpack_input16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t (&v)[PACKED_VALUES]) {
  uint32_t d[4][7];
  bool ok = unpack7(w0, d[0]);
  ok &= unpack7(w1, d[1]);
  ok &= unpack7(w2, d[2]);
  ok &= unpack7(w3, d[3]);
#pragma unroll
  for (int k = 0; k < PACKED_VALUES; k++) v[k] = d[k / 7][k % 7];
  return ok && d[3][6] == 0u;
}
PB_HD void pack_input16(const uint32_t (&v)[PACKED_VALUES], uint32_t (&w)[4]) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t d[7];
#pragma unroll
    for (int j = 0; j < 7; j++) d[j] = 7 * k + j < PACKED_VALUES ? v[7 * k + j] : 0u;
    w[k] = pack7(d);
  }
}

// one PROOF record (34 bytes, any alignment) -> eleven u16 of the packed record
PB_HD void pack_proof22(const uint8_t* rec, uint16_t (&out)[11]) {
#pragma unroll
  for (int j = 0; j < 9; j++) out[j] = (uint16_t)(rec[3 * j] | rec[3 * j + 1] << 7 | (rec[3 * j + 2] ? 1u : 0u) << 14);
  uint32_t d[7];
#pragma unroll
  for (int j = 0; j < 7; j++) d[j] = rec[27 + j];
  const uint32_t w = pack7(d);
  out[9] = (uint16_t)(w & 0xFFFFu);
  out[10] = (uint16_t)(w >> 16);
}
PB_HD bool unpack_proof22(const uint16_t (&in)[11], uint8_t* rec) {
#pragma unroll
  for (int j = 0; j < 9; j++) {
    rec[3 * j] = (uint8_t)(in[j] & 0x7Fu);
    rec[3 * j + 1] = (uint8_t)((in[j] >> 7) & 0x7Fu);
    rec[3 * j + 2] = (uint8_t)(in[j] >> 14);
  }
  uint32_t d[7];
  const bool ok = unpack7((uint32_t)in[9] | (uint32_t)in[10] << 16, d);
#pragma unroll
  for (int j = 0; j < 7; j++) rec[27 + j] = (uint8_t)d[j];
  return ok;
}

// ---- packed wire v3: the same information in 14 bytes in and 12 bytes per completed proof out ---------------------------
//
//   v3 input record, 14 bytes = three little-endian u32 + one little-endian u16.  Bits 0..28 of word k hold values
//   7k .. 7k+6 as seven base-17 digits (17^7 < 2^29); the last six values (alpha beta gamma z v u) form
//   G = sum_j v[21 + j] 17^j < 17^6 < 2^25: bits 0..15 of G are the u16, bits 16+3k .. 18+3k are bits 29..31 of word k.
//   27 log2(17) = 110.4 bits of information in 112.
//
//   v3 proof record, 12 bytes = three little-endian u32, for provers whose SRS lies on the curve (every commitment is
//   then one of the 102 points of E(F_101): y^2 = x^3 + 3).  A point travels as its INDEX in the list of those points --
//   0 = the point at infinity, then the affine points ordered by (x, y) -- which takes 7 bits.  Word k = index of point 3k
//   | index of point 3k+1 << 7 | index of point 3k+2 << 14 | (opening 2k + 17 * opening 2k+1) << 21 | two bits of the
//   seventh opening << 30 (bits 0-1 in word 0, 2-3 in word 1, bit 4 in word 2; bit 31 of word 2 is 0).
//   9 log2(102) + 7 log2(17) = 88.7 bits of information in 96.
constexpr int PACKED3_IN_BYTES = 14;
constexpr int PACKED3_PROOF_BYTES = 12;
constexpr uint32_t P17_6 = 24137569u;    // 17^6
constexpr int CURVE_POINTS = 102;

// the points of E(F_101), built at compile time: base[x] = index of the first affine point with abscissa x (two points
// share an abscissa when x^3 + 3 is a non-zero square: the smaller y comes first), px / py = the list itself
struct CurveIndexImage {
  uint8_t base[104];
  uint8_t px[104], py[104];
  constexpr CurveIndexImage() : base{}, px{}, py{} {
    uint32_t n = 1;                      // index 0: the point at infinity, stored as (0, 0)
    for (uint32_t x = 0; x < 101u; x++) {
      base[x] = (uint8_t)n;
      const uint32_t rhs = (x * x % 101u * x + 3u) % 101u;
      for (uint32_t y = 0; y < 101u; y++)
        if (y * y % 101u == rhs) { px[n] = (uint8_t)x; py[n] = (uint8_t)y; n++; }
    }
    base[101] = (uint8_t)n;              // = CURVE_POINTS on this curve (checked by a static_assert below)
  }
};
static_assert(CurveIndexImage().base[101] == CURVE_POINTS, "E(F_101): y^2 = x^3 + 3 has 102 points");
#ifdef __CUDACC__
__device__ __align__(16) const CurveIndexImage g_curve_index = CurveIndexImage();
#endif

// index of a point of the curve (coordinates as in a PROOF record).  The second point of an abscissa is the one with
// the larger y (the two roots are y and 101 - y).  For a point that is not on the curve the result is meaningless.
PB_HD uint32_t curve_index(const uint8_t* base, uint32_t x, uint32_t y, uint32_t inf) {
  const uint32_t i = base[x < 101u ? x : 0u] + (2u * y > 101u ? 1u : 0u);
  return inf ? 0u : i;
}

// 14-byte input record (as seven u16) -> the four words of the v2 decoder: words 0..2 without their top three bits, and
// G in place of v2's fourth word (same digits: alpha beta gamma z v u and a spare 0), so unpack_input16 finishes the job
PB_HD void input14_words(const uint16_t (&h)[7], uint32_t (&w)[4]) {
  uint32_t G = h[6];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const uint32_t x = (uint32_t)h[2 * k] | (uint32_t)h[2 * k + 1] << 16;
    w[k] = x & 0x1FFFFFFFu;
    G |= (x >> 29) << (16 + 3 * k);
  }
  w[3] = G;
}
// G alone (what the verifier needs) from the four u16 that carry it
PB_HD uint32_t input14_tail_word(uint32_t h1, uint32_t h3, uint32_t h5, uint32_t h6) {
  return h6 | (h1 >> 13) << 16 | (h3 >> 13) << 19 | (h5 >> 13) << 22;
}
PB_HD void pack_input14(const uint32_t (&v)[PACKED_VALUES], uint16_t (&h)[7]) {
  uint32_t G = v[26];
#pragma unroll
  for (int k = 25; k >= 21; k--) G = G * 17u + v[k];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    uint32_t d[7];
#pragma unroll
    for (int j = 0; j < 7; j++) d[j] = v[7 * k + j];
    const uint32_t w = pack7(d) | ((G >> (16 + 3 * k)) & 7u) << 29;
    h[2 * k] = (uint16_t)(w & 0xFFFFu);
    h[2 * k + 1] = (uint16_t)(w >> 16);
  }
  h[6] = (uint16_t)(G & 0xFFFFu);
}
// one PROOF record (34 bytes, any alignment) whose points are on the curve -> three u32
PB_HD void pack_proof12(const uint8_t* rec, const uint8_t* base, uint32_t (&out)[3]) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const uint8_t* p = rec + 3 * (3 * k + j);
      w |= curve_index(base, p[0], p[1], p[2]) << (7 * j);
    }
    w |= ((uint32_t)rec[27 + 2 * k] + 17u * rec[28 + 2 * k]) << 21;
    w |= (((uint32_t)rec[33] >> (2 * k)) & 3u) << 30;
    out[k] = w;
  }
}
// false = not an encoding (a point index >= 102, an opening >= 17)
PB_HD bool unpack_proof12(const uint32_t (&in)[3], const uint8_t* px, const uint8_t* py, uint8_t* rec) {
  bool ok = true;
  uint32_t last = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      uint32_t i = (in[k] >> (7 * j)) & 0x7Fu;
      ok = ok && i < (uint32_t)CURVE_POINTS;
      i = i < (uint32_t)CURVE_POINTS ? i : 0u;
      uint8_t* p = rec + 3 * (3 * k + j);
      p[0] = px[i]; p[1] = py[i]; p[2] = i == 0u ? 1 : 0;
    }
    const uint32_t pair = (in[k] >> 21) & 0x1FFu;
    ok = ok && pair < 289u;
    rec[27 + 2 * k] = (uint8_t)(pair % 17u);
    rec[28 + 2 * k] = (uint8_t)(pair / 17u % 17u);
    last |= (in[k] >> 30) << (2 * k);
  }
  ok = ok && last < 17u;
  rec[33] = (uint8_t)(last & 31u);
  return ok;
}

PB_HD uint32_t sv_byte(uint32_t status, uint32_t verdict) {
  const uint32_t s = status > 14u ? 15u : status, v = verdict > 14u ? 15u : verdict;
  return s | v << 4;
}

// ---- the synthetic stream of plonk.c_b200/workload.py (make_batch), bit for bit: draw j of item i of stream `seed` is
// splitmix64(seed + 16 i + j); j = 0 picks the witness row (mod 289), j = 1..9 the blinding scalars, j = 10..14 the
// challenges, j = 15 the verifier's u.  variant 0 ("U17"): draws mod 17; variant 1 ("NZ"): 1 + draws mod 16.
PB_HD uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
constexpr int SYNTH_WITNESS_ROWS = 289;   // (x, y, z) in F17^3 with x^2 + y^2 = z^2, lexicographic
PB_HD void synth_item(uint64_t seed, uint64_t i, int variant, const uint8_t* wtab /*[289][12]*/, uint32_t (&v)[PACKED_VALUES]) {
  const uint64_t base = seed + i * 16ull;
  const uint32_t row = (uint32_t)(splitmix64(base) % (uint64_t)SYNTH_WITNESS_ROWS);
#pragma unroll
  for (int k = 0; k < 12; k++) v[k] = wtab[row * 12u + k];
#pragma unroll
  for (int j = 1; j < 16; j++) {
    const uint64_t d = splitmix64(base + (uint64_t)j);
    v[11 + j] = variant == 0 ? (uint32_t)(d % 17ull) : (uint32_t)(d % 16ull) + 1u;
  }
}

}  // namespace pb
