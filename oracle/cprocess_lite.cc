// ORACLE — TEST INFRASTRUCTURE ONLY (see cprocess_lite.h).
// Restates src/microsimulation.cc:6-98 of the reference.
#include "cprocess_lite.h"

namespace ssim {

double rweibullHR(double shape, double scale, double hr) {  // microsimulation.cc:6-8
  return R::rweibull(shape, scale * pow(hr, 1.0 / shape));
}
Time now() { return Sim::clock(); }
Time simTime() { return Sim::clock(); }

static Rng *default_stream = 0, *current_stream = 0;
static double rn = 0.0;
static int counter_id = 0;

Rng::Rng() : RngStream() { id = ++counter_id; }
Rng::~Rng() {  // microsimulation.cc:22-25
  if (current_stream && current_stream->id == this->id) current_stream = default_stream;
}
void Rng::set() { current_stream = this; }

extern "C" {

void r_create_current_stream() {
  default_stream = new Rng();
  current_stream = default_stream;
}
void r_remove_current_stream() {
  delete default_stream;
  default_stream = 0;
}
void r_set_user_random_seed(double *inseed) {  // microsimulation.cc:44-51
  double seed[6];
  for (int i = 0; i < 6; i++) seed[i] = inseed[i];
  Rng::SetPackageSeed(seed);
  default_stream->SetSeed(seed);
}
void r_get_user_random_seed(double *outseed) { default_stream->GetState(outseed); }
void r_next_rng_substream() { default_stream->ResetNextSubstream(); }
void r_rng_advance_substream(double *inoutseed, int *n) {  // microsimulation.cc:65-75
  RngStream r;
  double seed[6];
  for (int i = 0; i < 6; i++) seed[i] = inoutseed[i];
  r.SetSeed(seed);
  r.AdvanceSubstream(0, *n);
  r.GetState(seed);
  for (int i = 0; i < 6; i++) inoutseed[i] = seed[i];
}
double *user_unif_rand() {  // microsimulation.cc:77-85
  if (!current_stream) return 0;
  rn = current_stream->RandU01();
  return &rn;
}

}  // extern "C"

}  // namespace ssim
