// transcript.cuh -- optional Fiat-Shamir mode.
//
// The reference's prover takes its five challenges from the caller (CHALLENGE, src/plonk.h:16-22,227); that mode stays
// the default and is bit-exact with the reference.  In Fiat-Shamir mode the challenges are drawn from a running hash of
// the public parameters and of the proof elements produced so far, so every round DEPENDS on the previous round's
// commitments -- in the kernel that serialises the commitment chains with the polynomial arithmetic of the next round,
// which the explicit-challenge mode is free to overlap.
//
// The hash is a duplex sponge over a 128-bit ARX state: four 32-bit words updated by the SipRound of HalfSipHash
// (add / rotate / xor; rotations 5, 16, 8, 7, 13, 16 -- restated from its public description, no interoperability with
// HalfSipHash outputs claimed).  Absorbing a word is HalfSipHash's message step (v3 ^= m, two rounds, v0 ^= m: ~30
// instructions), a challenge is its finalisation (v2 ^= 0xff, four rounds, v1 ^ v3) reduced mod 17;
// absorption then continues.  Round 1 shipped a 32-bit murmur3-finaliser mix.  Schedule (the specification the tests check
// against is oracle/fs_spec.inc; this file restates it):
//   state <- seed(circuit bytes, srs_len, SRS G1 bytes, G2 bytes)     (computed once per context, on the host)
//   absorb [a] [b] [c]            -> beta, gamma          absorb [z]                 -> alpha
//   absorb [t_lo] [t_mid] [t_hi]  -> z                    absorb a_z b_z c_z s1_z | s2_z r_z zw_z -> v
//   absorb [W_z] [W_zw]           -> u   (the verifier's opening-batching scalar)
// A point is absorbed as x | y << 8 | infinite << 16 of its PROOF-record bytes.
#pragma once
#include "curve.cuh"

namespace pb {

struct FsState { uint32_t v[4]; };
PB_HD uint32_t fs_rotl(uint32_t x, int b) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(x, x, b);
#else
  return (x << b) | (x >> (32 - b));
#endif
}
PB_HD void fs_sipround(FsState& s) {
  uint32_t v0 = s.v[0], v1 = s.v[1], v2 = s.v[2], v3 = s.v[3];
  v0 += v1; v1 = fs_rotl(v1, 5); v1 ^= v0; v0 = fs_rotl(v0, 16);
  v2 += v3; v3 = fs_rotl(v3, 8); v3 ^= v2;
  v0 += v3; v3 = fs_rotl(v3, 7); v3 ^= v0;
  v2 += v1; v1 = fs_rotl(v1, 13); v1 ^= v2; v2 = fs_rotl(v2, 16);
  s.v[0] = v0; s.v[1] = v1; s.v[2] = v2; s.v[3] = v3;
}
PB_HD FsState fs_absorb(FsState st, uint32_t word) {
  st.v[3] ^= word;
  fs_sipround(st);
  fs_sipround(st);
  st.v[0] ^= word;
  return st;
}
PB_HD uint32_t fs_challenge(FsState& st) {
  st.v[2] ^= 0xFFu;
#pragma unroll
  for (int r = 0; r < 4; r++) fs_sipround(st);
  // x mod 17 by exact division (0xF0F0F0F1 = ceil(2^36 / 17): floor(x / 17) for every 32-bit x).  NOT floor(x * 17 / 2^32):
  // ptxas 12.9 fuses that mul.hi with the additions that consume the challenge into IMAD.HI with a 64-bit addend and, in
  // the 96-register wide-table kernel, reused one such IMAD.HI after the upper half of its addend pair had changed --
  // a wrong [z] on the device only (profiles/r2/NOTES.md, "A ptxas miscompile").
  const uint32_t x = st.v[1] ^ st.v[3];
  return x - 17u * (mulhi_u32(x, 0xF0F0F0F1u) >> 4);
}
PB_HD uint32_t fs_point_word(const G1& p) { return p.x | p.y << 8 | p.inf << 16; }

// running transcript; ch = alpha beta gamma z v u (CHALLENGE field order, plonk.h:16-22, then u)
struct Transcript {
  FsState st;
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
PB_HD void fs_derive(const FsState& seed, const uint32_t (&pb)[27], const uint32_t (&op)[7], uint32_t (&ch)[5], uint32_t& u) {
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

// host side of context creation: the seed is the state after absorbing circuit and SRS
inline FsState fs_seed_host(const uint8_t* circuit44, const uint8_t* g1s, uint32_t srs_len, const uint8_t* g2) {
  auto le32 = [](const uint8_t* p, size_t avail) {
    uint32_t w = 0;
    for (size_t k = 0; k < 4 && k < avail; k++) w |= (uint32_t)p[k] << (8 * k);
    return w;
  };
  const uint32_t k0 = 0x504C4F4Eu, k1 = 0x4B2E6332u;   // "PLON" "K.c2" keyed into HalfSipHash's initial constants
  FsState st{{k0, k1, 0x6c796765u ^ k0, 0x74656462u ^ k1}};
  for (size_t k = 0; k < 44; k += 4) st = fs_absorb(st, le32(circuit44 + k, 4));
  st = fs_absorb(st, srs_len);
  const size_t nb = (size_t)srs_len * 3u;
  for (size_t k = 0; k < nb; k += 4) st = fs_absorb(st, le32(g1s + k, nb - k));
  st = fs_absorb(st, le32(g2, 4));
  return st;
}

}  // namespace pb
