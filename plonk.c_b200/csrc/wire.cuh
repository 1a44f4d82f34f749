// wire.cuh -- packed wire format v2 and the counter-based synthetic stream, as device functions.
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
