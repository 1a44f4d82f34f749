// transcript.cuh -- optional Fiat-Shamir mode.
//
// The reference's prover takes its five challenges from the caller (CHALLENGE, src/plonk.h:16-22,227); that mode stays
// the default and is bit-exact with the reference.  In Fiat-Shamir mode the challenges are drawn from a running hash of
// the public parameters and of the proof elements produced so far, so every round DEPENDS on the previous round's
// commitments -- in the kernel that serialises the commitment chains with the polynomial arithmetic of the next round,
// which the explicit-challenge mode is free to overlap.
//
// The hash is a 32-bit mix (the murmur3 finaliser: two IMADs and three shift-xors), a toy like the 17-element scalar
// field it feeds.  Schedule (the specification the tests check against is oracle/fs_spec.inc; this file restates it):
//   state <- seed(circuit bytes, srs_len, SRS G1 bytes, G2 bytes)
//   absorb [a] [b] [c]            -> beta, gamma          absorb [z]                 -> alpha
//   absorb [t_lo] [t_mid] [t_hi]  -> z                    absorb a_z b_z c_z s1_z | s2_z r_z zw_z -> v
//   absorb [W_z] [W_zw]           -> u   (the verifier's opening-batching scalar)
// A point is absorbed as x | y << 8 | infinite << 16 of its PROOF-record bytes, a challenge is floor(state * 17 / 2^32).
#pragma once
#include "curve.cuh"

namespace pb {

PB_HD uint32_t fs_mix(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
PB_HD uint32_t fs_absorb(uint32_t st, uint32_t word) { return fs_mix((st ^ word) + 0x9E3779B9u); }
PB_HD uint32_t fs_challenge(uint32_t& st) {
  st = fs_mix(st + 0x7F4A7C15u);
#ifdef __CUDA_ARCH__
  return __umulhi(st, 17u);
#else
  return (uint32_t)(((uint64_t)st * 17u) >> 32);
#endif
}
PB_HD uint32_t fs_point_word(const G1& p) { return p.x | p.y << 8 | p.inf << 16; }

// running transcript; ch = alpha beta gamma z v u (CHALLENGE field order, plonk.h:16-22, then u)
struct Transcript {
  uint32_t st;
  PB_HD void round1(const G1& a, const G1& b, const G1& c, uint32_t& beta, uint32_t& gamma) {
    st = fs_absorb(st, fs_point_word(a));
    st = fs_absorb(st, fs_point_word(b));
    st = fs_absorb(st, fs_point_word(c));
    beta = fs_challenge(st);
    gamma = fs_challenge(st);
  }
  PB_HD void round2(const G1& z, uint32_t& alpha) {
    st = fs_absorb(st, fs_point_word(z));
    alpha = fs_challenge(st);
  }
  PB_HD void round3(const G1& lo, const G1& mid, const G1& hi, uint32_t& zeta) {
    st = fs_absorb(st, fs_point_word(lo));
    st = fs_absorb(st, fs_point_word(mid));
    st = fs_absorb(st, fs_point_word(hi));
    zeta = fs_challenge(st);
  }
  // sc = a_z b_z c_z s1_z s2_z r_z zw_z (PROOF bytes 27..33)
  PB_HD void round4(const uint32_t (&sc)[7], uint32_t& v) {
    st = fs_absorb(st, sc[0] | sc[1] << 8 | sc[2] << 16 | sc[3] << 24);
    st = fs_absorb(st, sc[4] | sc[5] << 8 | sc[6] << 16);
    v = fs_challenge(st);
  }
  PB_HD void round5(const G1& wz, const G1& wzw, uint32_t& u) {
    st = fs_absorb(st, fs_point_word(wz));
    st = fs_absorb(st, fs_point_word(wzw));
    u = fs_challenge(st);
  }
};

// the verifier's side: all six challenges from the raw bytes of a PROOF record (pb = 27 commitment bytes, op = 7 openings)
PB_HD void fs_derive(uint32_t seed, const uint32_t (&pb)[27], const uint32_t (&op)[7], uint32_t (&ch)[5], uint32_t& u) {
  Transcript t{seed};
  G1 P[9];
#pragma unroll
  for (int j = 0; j < 9; j++) P[j] = G1{pb[3 * j], pb[3 * j + 1], pb[3 * j + 2]};
  t.round1(P[0], P[1], P[2], ch[1], ch[2]);
  t.round2(P[3], ch[0]);
  t.round3(P[4], P[5], P[6], ch[3]);
  t.round4(op, ch[4]);
  t.round5(P[7], P[8], u);
}

// host side of context creation: the seed binds circuit and SRS
inline uint32_t fs_seed_host(const uint8_t* circuit44, const uint8_t* g1s, uint32_t srs_len, const uint8_t* g2) {
  auto le32 = [](const uint8_t* p, size_t avail) {
    uint32_t w = 0;
    for (size_t k = 0; k < 4 && k < avail; k++) w |= (uint32_t)p[k] << (8 * k);
    return w;
  };
  uint32_t st = 0x504C4F4Eu;   // "PLON"
  for (size_t k = 0; k < 44; k += 4) st = fs_absorb(st, le32(circuit44 + k, 4));
  st = fs_absorb(st, srs_len);
  const size_t nb = (size_t)srs_len * 3u;
  for (size_t k = 0; k < nb; k += 4) st = fs_absorb(st, le32(g1s + k, nb - k));
  st = fs_absorb(st, le32(g2, 4));
  return st;
}

}  // namespace pb
