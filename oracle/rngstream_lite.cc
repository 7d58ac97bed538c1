// ORACLE — TEST INFRASTRUCTURE ONLY (see rngstream_lite.h).
#include "rngstream_lite.h"

namespace ssim {

namespace {

typedef uint64_t u64;
typedef u64 Mat[3][3];

const u64 M1 = 4294967087ULL;  // 2^32 - 209       (RngStream.cpp:39)
const u64 M2 = 4294944443ULL;  // 2^32 - 22853     (RngStream.cpp:40)
const double NORM = 1.0 / (4294967087.0 + 1.0);  // RngStream.cpp:41
const double FACT = 5.9604644775390625e-8;       // 2^-24, RngStream.cpp:48

// One-step transition matrices and their inverses (RngStream.cpp:53-75);
// negative entries are stored as residues.
const Mat A1 = {{0, 1, 0}, {0, 0, 1}, {M1 - 810728ULL, 1403580ULL, 0}};
const Mat A2 = {{0, 1, 0}, {0, 0, 1}, {M2 - 1370589ULL, 0, 527612ULL}};
const Mat IA1 = {{184888585ULL, 0, 1945170933ULL}, {1, 0, 0}, {0, 1, 0}};
const Mat IA2 = {{0, 360363334ULL, 4225571728ULL}, {1, 0, 0}, {0, 1, 0}};

inline u64 mulmod(u64 a, u64 b, u64 m) { return (a * b) % m; }  // a,b < 2^32

void matvec(const Mat A, const u64 s[3], u64 v[3], u64 m) {
  u64 x[3];
  for (int i = 0; i < 3; i++)
    x[i] = (mulmod(A[i][0], s[0], m) + mulmod(A[i][1], s[1], m) + mulmod(A[i][2], s[2], m)) % m;
  for (int i = 0; i < 3; i++) v[i] = x[i];
}

void matmat(const Mat A, const Mat B, Mat C, u64 m) {
  Mat W;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      W[i][j] = (mulmod(A[i][0], B[0][j], m) + mulmod(A[i][1], B[1][j], m) +
                 mulmod(A[i][2], B[2][j], m)) % m;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i][j] = W[i][j];
}

void copy(const Mat A, Mat B) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B[i][j] = A[i][j];
}

void two_pow(const Mat A, Mat B, u64 m, int e) {  // B = A^(2^e)
  copy(A, B);
  for (int i = 0; i < e; i++) matmat(B, B, B, m);
}

void power(const Mat A, Mat B, u64 m, int64_t n) {  // B = A^n
  Mat W;
  copy(A, W);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B[i][j] = (i == j);
  while (n > 0) {
    if (n & 1) matmat(W, B, B, m);
    matmat(W, W, W, m);
    n >>= 1;
  }
}

// Base jump matrices: which = 0 single step, 1 substream (2^76 steps),
// 2 stream (2^127 steps); RngStream.cpp:53-116.
struct Jumps {
  Mat f1[3], f2[3], b1[3], b2[3];
  Jumps() {
    const int ex[3] = {0, 76, 127};
    for (int k = 0; k < 3; k++) {
      two_pow(A1, f1[k], M1, ex[k]);
      two_pow(A2, f2[k], M2, ex[k]);
      two_pow(IA1, b1[k], M1, ex[k]);
      two_pow(IA2, b2[k], M2, ex[k]);
    }
  }
};
const Jumps &jumps() {
  static Jumps J;
  return J;
}

int bad_seed(const double s[6]) {  // RngStream.cpp:234-268
  for (int i = 0; i < 3; i++)
    if (s[i] >= 4294967087.0) return -1;
  for (int i = 3; i < 6; i++)
    if (s[i] >= 4294944443.0) return -1;
  if (s[0] == 0 && s[1] == 0 && s[2] == 0) return -1;
  if (s[3] == 0 && s[4] == 0 && s[5] == 0) return -1;
  return 0;
}

}  // namespace

u64 RngStream::nextSeed[6] = {12345, 12345, 12345, 12345, 12345, 12345};  // RngStream.cpp:328-331

RngStream::RngStream(const char *s) : name(s) {  // RngStream.cpp:337-355
  anti = false;
  incPrec = false;
  for (int i = 0; i < 6; i++) Bg[i] = Cg[i] = Ig[i] = nextSeed[i];
  matvec(jumps().f1[2], nextSeed, nextSeed, M1);
  matvec(jumps().f2[2], &nextSeed[3], &nextSeed[3], M2);
}

void RngStream::ResetStartStream() {
  for (int i = 0; i < 6; i++) Cg[i] = Bg[i] = Ig[i];
}
void RngStream::ResetStartSubstream() {
  for (int i = 0; i < 6; i++) Cg[i] = Bg[i];
}
void RngStream::ResetNextSubstream() {  // RngStream.cpp:381-387
  matvec(jumps().f1[1], Bg, Bg, M1);
  matvec(jumps().f2[1], &Bg[3], &Bg[3], M2);
  for (int i = 0; i < 6; i++) Cg[i] = Bg[i];
}

bool RngStream::SetPackageSeed(const double seed[6]) {  // RngStream.cpp:391-398
  if (bad_seed(seed)) return false;
  for (int i = 0; i < 6; i++) nextSeed[i] = (u64)seed[i];
  return true;
}
bool RngStream::SetSeed(const double seed[6]) {  // RngStream.cpp:402-409
  if (bad_seed(seed)) return false;
  for (int i = 0; i < 6; i++) Cg[i] = Bg[i] = Ig[i] = (u64)seed[i];
  return true;
}

// n = ±2^|e| + c steps of the chosen base jump (RngStream.cpp:418-447)
void RngStream::jump(int32_t e, int32_t c, int which) {
  const Jumps &J = jumps();
  Mat B1, B2, C1, C2;
  if (e > 0) {
    two_pow(J.f1[which], B1, M1, e);
    two_pow(J.f2[which], B2, M2, e);
  } else if (e < 0) {
    two_pow(J.b1[which], B1, M1, -e);
    two_pow(J.b2[which], B2, M2, -e);
  }
  if (c >= 0) {
    power(J.f1[which], C1, M1, c);
    power(J.f2[which], C2, M2, c);
  } else {
    power(J.b1[which], C1, M1, -(int64_t)c);
    power(J.b2[which], C2, M2, -(int64_t)c);
  }
  if (e) {
    matmat(B1, C1, C1, M1);
    matmat(B2, C2, C2, M2);
  }
  matvec(C1, Cg, Cg, M1);
  matvec(C2, &Cg[3], &Cg[3], M2);
}

void RngStream::AdvanceState(int32_t e, int32_t c) { jump(e, c, 0); }
void RngStream::AdvanceSubstream(int32_t e, int32_t c) {  // RngStream.cpp:454-458
  jump(e, c, 1);
  for (int i = 0; i < 6; i++) Bg[i] = Cg[i];
}
void RngStream::AdvanceStream(int32_t e, int32_t c) {
  jump(e, c, 2);
  for (int i = 0; i < 6; i++) Ig[i] = Bg[i] = Cg[i];
}

void RngStream::GetState(double seed[6]) const {
  for (int i = 0; i < 6; i++) seed[i] = (double)Cg[i];
}
void RngStream::IncreasedPrecis(bool incp) { incPrec = incp; }
void RngStream::SetAntithetic(bool a) { anti = a; }

double RngStream::U01() {  // RngStream.cpp:276-299
  int64_t d1 = (int64_t)(1403580ULL * Cg[1]) - (int64_t)(810728ULL * Cg[0]);
  int64_t p1 = d1 % (int64_t)M1;
  if (p1 < 0) p1 += (int64_t)M1;
  Cg[0] = Cg[1];
  Cg[1] = Cg[2];
  Cg[2] = (u64)p1;
  int64_t d2 = (int64_t)(527612ULL * Cg[5]) - (int64_t)(1370589ULL * Cg[3]);
  int64_t p2 = d2 % (int64_t)M2;
  if (p2 < 0) p2 += (int64_t)M2;
  Cg[3] = Cg[4];
  Cg[4] = Cg[5];
  Cg[5] = (u64)p2;
  double u = (p1 > p2) ? (double)(p1 - p2) * NORM : (double)(p1 - p2 + (int64_t)M1) * NORM;
  return anti ? (1 - u) : u;
}

double RngStream::U01d() {  // RngStream.cpp:305-318
  double u = U01();
  if (anti) {
    u += (U01() - 1.0) * FACT;
    return (u < 0.0) ? u + 1.0 : u;
  }
  u += U01() * FACT;
  return (u < 1.0) ? u : (u - 1.0);
}

double RngStream::RandU01() { return incPrec ? U01d() : U01(); }
int RngStream::RandInt(int low, int high) {
  return low + static_cast<int>((high - low + 1.0) * RandU01());
}

}  // namespace ssim
