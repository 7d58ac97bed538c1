// Host-side RNG bookkeeping of the C ABI (plain C++, no CUDA).
//
// Mirrors the static state the reference keeps around its RngStream objects:
//   RngStream::nextSeed (the "package seed")  src/RngStream.cpp:328-355, 391-398
//   default_stream / current_stream            src/microsimulation.cc:19-75
//   R's Mersenne-Twister state + set.seed      libR main/RNG.c
// so that a sequence of .C/.Call-equivalent calls leaves the same generator
// state behind as the reference would.
#pragma once
#include <stdint.h>
#include <string.h>

#include "mrg32k3a.cuh"

namespace msim {

inline uint32_t mrg_modpow(uint32_t a, uint64_t e, uint32_t m) {
  uint64_t r = 1, b = a % m;
  while (e) {
    if (e & 1) r = (r * b) % m;
    b = (b * b) % m;
    e >>= 1;
  }
  return (uint32_t)r;
}

// inverse of the one-step matrix (RngStream.cpp:53-63 "InvA1/InvA2"), derived
// from the recurrence: m1 and m2 are prime, so 1/a = a^(m-2).
inline MrgJump mrg_jump_one_step_inverse() {
  MrgJump J;
  memset(&J, 0, sizeof J);
  uint32_t i13 = mrg_modpow(810728u, (uint64_t)MRG_M1 - 2, MRG_M1);
  uint32_t i23 = mrg_modpow(1370589u, (uint64_t)MRG_M2 - 2, MRG_M2);
  J.a[0] = (uint32_t)(((uint64_t)1403580u * i13) % MRG_M1);
  J.a[2] = MRG_M1 - i13;
  J.a[3] = 1;
  J.a[7] = 1;
  J.b[1] = (uint32_t)(((uint64_t)527612u * i23) % MRG_M2);
  J.b[2] = MRG_M2 - i23;
  J.b[3] = 1;
  J.b[7] = 1;
  return J;
}
inline const MrgJump &mrg_substream_jump_inverse() {
  static const MrgJump J = mrg_jump_two_pow(mrg_jump_one_step_inverse(), 76);
  return J;
}

inline void seed_to_u32(const double seed[6], uint32_t s[6]) {
  for (int i = 0; i < 6; i++) s[i] = (uint32_t)seed[i];
}

// a host RngStream: Cg (current), Bg (substream start), Ig (stream start)
struct HostStream {
  Mrg32k3a Cg, Bg, Ig;
  void set_all(const uint32_t s[6]) {
    for (int i = 0; i < 6; i++) Cg.s[i] = Bg.s[i] = Ig.s[i] = s[i];
  }
  void reset_next_substream() {  // RngStream.cpp:381-387
    Bg.jump(mrg_substream_jump());
    Cg = Bg;
  }
  void advance_substream(int32_t e, int32_t c) {  // RngStream.cpp:418-458
    MrgJump C = c >= 0 ? mrg_jump_pow(mrg_substream_jump(), (uint64_t)c)
                       : mrg_jump_pow(mrg_substream_jump_inverse(), (uint64_t)(-(int64_t)c));
    if (e > 0) C = mrg_jump_mul(mrg_jump_two_pow(mrg_substream_jump(), e), C);
    else if (e < 0) C = mrg_jump_mul(mrg_jump_two_pow(mrg_substream_jump_inverse(), -e), C);
    Cg.jump(C);
    Bg = Cg;
  }
};

struct HostRngState {
  uint32_t next_seed[6];  // RngStream::nextSeed
  bool have_default;
  HostStream default_stream;
  uint32_t mt[625];  // R's dummy[]: position, then the 624 state words
  bool mt_seeded;

  HostRngState() : have_default(false), mt_seeded(false) {
    for (int i = 0; i < 6; i++) next_seed[i] = 12345u;
    memset(mt, 0, sizeof mt);
  }
  // `new RngStream()`: takes nextSeed, then nextSeed *= A^(2^127)  (RngStream.cpp:337-355)
  HostStream new_stream() {
    HostStream s;
    s.set_all(next_seed);
    Mrg32k3a t;
    memcpy(t.s, next_seed, sizeof next_seed);
    t.jump(mrg_stream_jump());
    memcpy(next_seed, t.s, sizeof next_seed);
    return s;
  }
  // set.seed(seed) with kind Mersenne-Twister
  void mt_set_seed(uint32_t seed) {
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
    for (int j = 0; j < 625; j++) {
      seed = 69069u * seed + 1u;
      mt[j] = seed;
    }
    mt[0] = 624;  // FixupSeeds: regenerate on first use
    mt_seeded = true;
  }
};

}  // namespace msim
