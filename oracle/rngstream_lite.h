// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Restatement of the reference's RngStream (L'Ecuyer MRG32k3a with streams
// and substreams): inst/include/RngStream.h:30-111, src/RngStream.cpp:32-583.
// The reference carries the state as doubles and reduces with floating-point
// division (RngStream.cpp:122-139, 276-299).  The recurrences are exact
// integer recurrences, so this restatement keeps the same public interface
// (state still exchanged as double[6]) but computes in 64-bit integers.
// tests/test_oracle_ref.py checks it draw for draw against the reference's
// own file compiled into oracle/_ref.
#ifndef MSIM_ORACLE_RNGSTREAM_LITE_H
#define MSIM_ORACLE_RNGSTREAM_LITE_H

#include <stdint.h>
#include <string>

namespace ssim {

class RngStream {
 public:
  RngStream(const char *name = "");
  virtual ~RngStream() {}
  static bool SetPackageSeed(const double seed[6]);
  void ResetStartStream();
  void ResetStartSubstream();
  void ResetNextSubstream();
  void SetAntithetic(bool a);
  void IncreasedPrecis(bool incp);
  bool SetSeed(const double seed[6]);
  void AdvanceState(int32_t e, int32_t c);
  void AdvanceSubstream(int32_t e, int32_t c);
  void AdvanceStream(int32_t e, int32_t c);
  void GetState(double seed[6]) const;
  double RandU01();
  int RandInt(int i, int j);

 private:
  uint64_t Cg[6], Bg[6], Ig[6];  // current / substream start / stream start
  bool anti, incPrec;
  std::string name;
  static uint64_t nextSeed[6];
  double U01();
  double U01d();
  void jump(int32_t e, int32_t c, int which);
};

}  // namespace ssim

#endif
