// Integer MRG32k3a (L'Ecuyer) with the reference's stream / substream
// structure, for host and device.
//
// Replaces the double-precision arithmetic of the reference's RngStream
// (src/RngStream.cpp:122-159 MultModM/MatVecModM, :276-299 U01, :381-387
// ResetNextSubstream, :418-458 AdvanceSubstream).  The recurrences are exact
// integer recurrences mod m1 = 2^32-209 and m2 = 2^32-22853, so every state
// and every uniform is bit-identical to the reference's; reductions use
// 2^32 = C (mod m) folding instead of division.
#pragma once
#include <stdint.h>

// MSIM_DEV marks per-person engine/model code.  Under nvcc it is device code;
// compiled by a plain C++ compiler (tests/host_emulation only) the same
// templates run on the CPU so that heap, RNG and report logic can be checked
// against the oracle on a GPU-less machine.  The product library is always
// built by nvcc and has no host execution path.
#ifdef __CUDACC__
#define MSIM_HD __host__ __device__ __forceinline__
#define MSIM_DEV __device__ __forceinline__
#define MSIM_DEV_NOINLINE __device__ __noinline__
#else
#define MSIM_HD inline
#define MSIM_DEV inline
#define MSIM_DEV_NOINLINE inline
#endif

namespace msim {

constexpr uint32_t MRG_M1 = 4294967087u;  // 2^32 - 209
constexpr uint32_t MRG_M2 = 4294944443u;  // 2^32 - 22853
constexpr uint32_t MRG_C1 = 209u;
constexpr uint32_t MRG_C2 = 22853u;
constexpr double MRG_NORM = 2.328306549295727688e-10;  // 1/(m1+1), RngStream.cpp:41

// Reduction mod m = 2^32 - C by folding 2^32 = C (mod m).  Written so that the compiler emits the short
// sequences (profiles/r02_*): a canonical residue costs one 32-bit multiply-add, two compares and a predicated add
// once the value is below (C+1)*2^32.
//
// (hi*2^32 + lo) mod m, canonical, for hi*C < 2^32.
//   r = hi*C + lo (mod 2^32).  If the sum carried, the value is 2^32 + r = r + C (mod m) with r < hi*C, so r + C < m.
//   Otherwise r < 2^32 < 2m: subtract m once if r >= m, i.e. add C and drop the carry.
template <uint32_t C>
MSIM_HD uint32_t mrg_fold_small(uint32_t hi, uint32_t lo) {
  uint32_t r = hi * C + lo;
  if ((r < lo) | (r >= (0u - C))) r += C;
  return r;
}

// one fold of a 64-bit value: (x >> 32)*C + (x & 0xffffffff) < (C+1)*2^32, congruent to x
template <uint32_t C>
MSIM_HD uint64_t mrg_fold64(uint64_t x) {
  return (x >> 32) * C + (uint32_t)x;
}

// x mod m, canonical, for any 64-bit x: after one fold the high word is <= C and C*C < 2^32
template <uint32_t C>
MSIM_HD uint32_t mrg_reduce(uint64_t x) {
  const uint64_t y = mrg_fold64<C>(x);
  return mrg_fold_small<C>((uint32_t)(y >> 32), (uint32_t)y);
}
template <uint32_t C>
MSIM_HD uint32_t mrg_fold2(uint64_t x) {  // (older name)
  return mrg_reduce<C>(x);
}

// (a0*s0 + a1*s1 + a2*s2) mod m for canonical inputs (all < m).  A product of two residues is at most
// (2^32 - C - 1)^2 = 2^64 - 2(C+1)*2^32 + (C+1)^2 and a folded value is below (C+1)*2^32, so
// fold(acc) + a*s < 2^64 - (C+1)*2^32 + (C+1)^2 < 2^64: the running sum never overflows.
template <uint32_t C>
MSIM_HD uint32_t mrg_dot3(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t s0, uint32_t s1, uint32_t s2) {
  uint64_t acc = (uint64_t)a0 * s0;
  acc = mrg_fold64<C>(acc) + (uint64_t)a1 * s1;
  acc = mrg_fold64<C>(acc) + (uint64_t)a2 * s2;
  return mrg_reduce<C>(acc);
}

// 3x3 transition matrix pair (component 1 mod m1, component 2 mod m2), row-major.
struct MrgJump {
  uint32_t a[9];
  uint32_t b[9];
};

struct Mrg32k3a {
  uint32_t s[6];  // s10 s11 s12 s20 s21 s22 (the reference's Cg[0..5])

  // RngStream::U01 (RngStream.cpp:276-299), anti = false, incPrec = false
  MSIM_HD double u01() {
    // p1 = (1403580*s11 - 810728*s10) mod m1, with -x = (m1 - x)
    uint64_t t1 = 1403580ull * s[1] + 810728ull * (uint64_t)(MRG_M1 - s[0]);  // < 2^54: high word * C1 < 2^30
    uint32_t p1 = mrg_fold_small<MRG_C1>((uint32_t)(t1 >> 32), (uint32_t)t1);
    s[0] = s[1];
    s[1] = s[2];
    s[2] = p1;
    // p2 = (527612*s22 - 1370589*s20) mod m2
    uint64_t t2 = 527612ull * s[5] + 1370589ull * (uint64_t)(MRG_M2 - s[3]);  // < 2^54
    uint32_t p2 = mrg_reduce<MRG_C2>(t2);
    s[3] = s[4];
    s[4] = s[5];
    s[5] = p2;
    // (p1 - p2) or (p1 - p2 + m1): an integer in [1, m1], exact in double
    uint32_t d = p1 - p2;
    if (p1 <= p2) d += MRG_M1;
#ifdef __CUDA_ARCH__
    // d as a double without an integer conversion: 2^52 + d is the bit pattern (0x43300000, d); the multiply-add
    // (2^52 + d) * norm - 2^52 * norm is exact before its one rounding, so this is RN(d * norm), the reference's value
    return __fma_rn(__hiloint2double(0x43300000, (int)d), MRG_NORM, -(4503599627370496.0 * MRG_NORM));
#else
    return (double)d * MRG_NORM;
#endif
  }

  // state <- J * state (MatVecModM on both components)
  MSIM_HD void jump(const MrgJump &J) {
    uint32_t x0 = mrg_dot3<MRG_C1>(J.a[0], J.a[1], J.a[2], s[0], s[1], s[2]);
    uint32_t x1 = mrg_dot3<MRG_C1>(J.a[3], J.a[4], J.a[5], s[0], s[1], s[2]);
    uint32_t x2 = mrg_dot3<MRG_C1>(J.a[6], J.a[7], J.a[8], s[0], s[1], s[2]);
    uint32_t y0 = mrg_dot3<MRG_C2>(J.b[0], J.b[1], J.b[2], s[3], s[4], s[5]);
    uint32_t y1 = mrg_dot3<MRG_C2>(J.b[3], J.b[4], J.b[5], s[3], s[4], s[5]);
    uint32_t y2 = mrg_dot3<MRG_C2>(J.b[6], J.b[7], J.b[8], s[3], s[4], s[5]);
    s[0] = x0; s[1] = x1; s[2] = x2;
    s[3] = y0; s[4] = y1; s[5] = y2;
  }
};

constexpr int JUMP_TAB = 1024;  // radix of the per-thread start-state tables

// start state of global thread g: J_hi[g / JUMP_TAB] * J_lo[g % JUMP_TAB] * base
struct StreamStart {
  uint32_t base[6];        // substream state of the shard's first person
  const MrgJump *lo, *hi;  // M^j and M^(JUMP_TAB*j), j < JUMP_TAB, M = A^(2^76)
  MrgJump stride;          // M^(gridDim.x*BLOCK)
};

MSIM_HD Mrg32k3a thread_start_state(const StreamStart &st, unsigned g) {
  Mrg32k3a r;
  for (int k = 0; k < 6; k++) r.s[k] = st.base[k];
  MrgJump J = st.lo[g % JUMP_TAB];
  r.jump(J);
  unsigned h = g / JUMP_TAB;
  if (h) {
    J = st.hi[h];
    r.jump(J);
  }
  return r;
}

// ---- host-side matrix algebra (exact, 64-bit) -------------------------

inline void mrg_matmul(const uint32_t A[9], const uint32_t B[9], uint32_t Cm[9], uint32_t m) {
  uint32_t W[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      uint64_t acc = 0;
      for (int k = 0; k < 3; k++) acc = (acc + ((uint64_t)A[3 * i + k] * B[3 * k + j]) % m) % m;
      W[3 * i + j] = (uint32_t)acc;
    }
  for (int i = 0; i < 9; i++) Cm[i] = W[i];
}

inline MrgJump mrg_jump_mul(const MrgJump &X, const MrgJump &Y) {
  MrgJump Z;
  mrg_matmul(X.a, Y.a, Z.a, MRG_M1);
  mrg_matmul(X.b, Y.b, Z.b, MRG_M2);
  return Z;
}

inline MrgJump mrg_jump_identity() {
  MrgJump I;
  for (int i = 0; i < 9; i++) I.a[i] = I.b[i] = (i % 4 == 0);
  return I;
}

// One step of the generator as a matrix (RngStream.cpp:65-75 "A1p0/A2p0").
inline MrgJump mrg_jump_one_step() {
  MrgJump J = {{0, 1, 0, 0, 0, 1, MRG_M1 - 810728u, 1403580u, 0}, {0, 1, 0, 0, 0, 1, MRG_M2 - 1370589u, 0, 527612u}};
  return J;
}

inline MrgJump mrg_jump_two_pow(MrgJump J, int e) {  // J^(2^e)
  for (int i = 0; i < e; i++) J = mrg_jump_mul(J, J);
  return J;
}

inline MrgJump mrg_jump_pow(MrgJump J, uint64_t n) {  // J^n
  MrgJump R = mrg_jump_identity();
  while (n) {
    if (n & 1) R = mrg_jump_mul(J, R);
    J = mrg_jump_mul(J, J);
    n >>= 1;
  }
  return R;
}

// Substream jump A^(2^76) and stream jump A^(2^127) (RngStream.cpp:77-99).
inline const MrgJump &mrg_substream_jump() {
  static const MrgJump J = mrg_jump_two_pow(mrg_jump_one_step(), 76);
  return J;
}
inline const MrgJump &mrg_stream_jump() {
  static const MrgJump J = mrg_jump_two_pow(mrg_jump_one_step(), 127);
  return J;
}

// CheckSeed (RngStream.cpp:234-268)
inline bool mrg_seed_ok(const double seed[6]) {
  for (int i = 0; i < 3; i++)
    if (!(seed[i] >= 0) || seed[i] >= 4294967087.0) return false;
  for (int i = 3; i < 6; i++)
    if (!(seed[i] >= 0) || seed[i] >= 4294944443.0) return false;
  if (seed[0] == 0 && seed[1] == 0 && seed[2] == 0) return false;
  if (seed[3] == 0 && seed[4] == 0 && seed[5] == 0) return false;
  return true;
}

}  // namespace msim
