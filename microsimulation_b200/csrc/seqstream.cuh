// Sequential-stream regime helpers shared by the kernels and the host
// emulation used in CPU tests: R's Mersenne-Twister output transform, the
// per-wave regeneration step, and the chunk walks that resolve each person's
// offset into the single global stream.
#pragma once
#include <stdint.h>

#include "mrg32k3a.cuh"

namespace msim {

// ---- Mersenne-Twister stream (R's default generator) ---------------------
//
// R: MT19937, 624-word state regenerated in place; unif_rand() = fixup(temper(
// mt[mti++]) * 2.3283064365386963e-10)  (SURVEY.md Appendix A.1).  The device
// keeps the RAW state words of consecutive regenerations in one array
// (batch 0 = the state passed in); a consumer tempers on the fly, and the
// state R must be left in afterwards is simply the batch its last draw fell in.

MSIM_HD double mt_word_to_unif(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  double x = (double)y * 2.3283064365386963e-10;  // [0,1)
  const double i2_32m1 = 2.328306437080797e-10;
  if (x <= 0.0) return 0.5 * i2_32m1;
  if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
  return x;
}

struct StreamSource {
  const uint32_t *w;
  int64_t pos, end;
  int overrun;
  MSIM_HD double unif_rand() {
    if (pos >= end) {
      overrun = 1;
      return 0.5;
    }
    return mt_word_to_unif(w[pos++]);
  }
};

// One lane's share of an in-place MT19937 regeneration.  The 624 words are
// renewed in three dependent waves, kk in [0,227), [227,454), [454,624): wave 1
// reads wave 0's new words, wave 2 wave 1's (and, for kk = 623, the new
// mt[0]).  Lane t < 227 of wave w owns kk = 227*w + t; all lanes compute from
// the array as it stands, then all store (barrier in between on the device).
MSIM_HD bool mt_wave_value(const uint32_t *mt, int wave, int t, int *kk_out, uint32_t *v_out) {
  int kk = wave * 227 + t;
  int hi = (wave == 2) ? 624 : (wave + 1) * 227;
  if (t >= 227 || kk >= hi) return false;
  uint32_t a = mt[kk], nb = mt[(kk + 1 == 624) ? 0 : kk + 1];
  uint32_t y = (a & 0x80000000u) | (nb & 0x7fffffffu);
  int src = (kk < 227) ? kk + 397 : kk - 227;
  *kk_out = kk;
  *v_out = mt[src] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  return true;
}

constexpr int CHAIN_CHUNK = 2048;   // positions per chunk
constexpr int CHAIN_MAX_DRAWS = 8;  // entry offsets tracked per chunk (>= the model's max draws per person)

// Walk p += len[p] through one chunk starting at entry offset e.  Returns the
// offset at which the chain enters the next chunk (255 = broken) and the
// number of persons that start inside this chunk.
MSIM_HD int chain_walk(const unsigned char *chunk_len, int e, int *count) {
  int p = e, cnt = 0;
  while (p < CHAIN_CHUNK) {
    int l = chunk_len[p];
    if (l == 255 || l == 0) {
      *count = cnt;
      return 255;
    }
    p += l;
    cnt++;
  }
  *count = cnt;
  return p - CHAIN_CHUNK;
}

// One step of the serial pass over chunks.
MSIM_HD void chain_scan_step(const unsigned char *exit_c, const int *count_c, int64_t n, int *e, int64_t *b) {
  if (*e == 255 || *b >= n) {
    *e = 255;  // persons beyond n are never needed
    return;
  }
  *b += count_c[*e];
  *e = exit_c[*e];
}

// Second walk: word index of every person's first draw, for person indices < n.
// Returns false if the chain is broken inside the chunk.
MSIM_HD bool chain_emit(const unsigned char *chunk_len, int e, int64_t b, int64_t n, int64_t chunk_first_word,
                        int64_t *starts, int64_t *end_word) {
  int p = e;
  while (p < CHAIN_CHUNK && b < n) {
    int l = chunk_len[p];
    if (l == 255 || l == 0) return false;
    starts[b] = chunk_first_word + p;
    if (b == n - 1) *end_word = chunk_first_word + p + l;
    p += l;
    b++;
  }
  return true;
}

}  // namespace msim
