// Per-packet random number streams.
//
//  * RNG_PHILOX (production): counter-based Philox4x32-10 (Salmon et al. 2011). key = (seed, packet number),
//    counter = (draw block within the timestep, timestep, high word of the seed, rank). A packet's stream therefore depends on
//    nothing but its identity and the timestep: no generator state has to survive between timesteps, and
//    the result is independent of how packets are scheduled onto threads, launches or GPUs.
//  * RNG_XOSHIRO (parity): the reference's own per-packet generator in its GPU_ON build, Xoshiro128++ seeded
//    through SplitMix32 (reference random.h:65-138), with its 16-byte state carried in the 256-byte Packet
//    (packet.h:110-114). Used to compare packet histories one-to-one against the oracle.
//
// Both return U[0,1) as a FLOAT built from 24 random bits, as the reference does (random.h:140-192):
// every random number on the hot path is float32 before it enters the double-precision physics.
#pragma once
#include "hd.h"

namespace ab {

constexpr int RNG_PHILOX = 0;
constexpr int RNG_XOSHIRO = 1;

AHD unsigned int rotl32(const unsigned int x, const unsigned int k) { return (x << k) | (x >> (32U - k)); }

AHD unsigned int mulhi32(const unsigned int a, const unsigned int b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return static_cast<unsigned int>((static_cast<unsigned long long>(a) * b) >> 32U);
#endif
}

struct RngSetup {
  int mode;
  unsigned int key0;  // philox key word 0 (seed)
  unsigned int ctr1;  // philox counter word 1 (timestep)
  unsigned int ctr2;  // philox counter word 2 (high word of the seed)
  unsigned int ctr3;  // philox counter word 3 (rank)
};

struct Rng {
  unsigned int s0, s1, s2, s3;  // xoshiro state, or philox (draw counter, packet number, -, -)
  // run-wide settings are read where they are needed (kernel parameter space) instead of living in registers
  const RngSetup* setup;
  unsigned int b0, b1, b2, b3;  // philox: the four outputs of block (s0 >> 2); valid when have_block != 0
  int have_block;

  AHD void philox_block(const unsigned int blockindex) {
    // Philox4x32-10 (Salmon et al. 2011): counter (block, timestep, seed high word, rank), key (seed, packet number)
    unsigned int c0 = blockindex;
    unsigned int c1 = setup->ctr1;
    unsigned int c2 = setup->ctr2;
    unsigned int c3 = setup->ctr3;
    unsigned int k0 = setup->key0;
    unsigned int k1 = s1;
#pragma unroll
    for (int round = 0; round < 10; round++) {
      const unsigned int hi0 = mulhi32(0xD2511F53U, c0);
      const unsigned int lo0 = 0xD2511F53U * c0;
      const unsigned int hi1 = mulhi32(0xCD9E8D57U, c2);
      const unsigned int lo1 = 0xCD9E8D57U * c2;
      c0 = hi1 ^ c1 ^ k0;
      c1 = lo1;
      c2 = hi0 ^ c3 ^ k1;
      c3 = lo0;
      k0 += 0x9E3779B9U;
      k1 += 0xBB67AE85U;
    }
    b0 = c0;
    b1 = c1;
    b2 = c2;
    b3 = c3;
    have_block = 1;
  }

  AHD unsigned int next_u32() {
    if (setup->mode == RNG_XOSHIRO) {
      // Xoshiro128++ (Blackman & Vigna), same output function as reference random.h:124-135
      const unsigned int result = rotl32(s0 + s3, 7U) + s0;
      const unsigned int t = s1 << 9U;
      s2 ^= s0;
      s3 ^= s1;
      s1 ^= s2;
      s0 ^= s3;
      s2 ^= t;
      s3 = rotl32(s3, 11U);
      return result;
    }
    // draw number s0 is word (s0 & 3) of Philox block (s0 >> 2): one 10-round evaluation serves four draws, and
    // the value of a draw depends only on its index, not on where a history was split across kernel launches
    const unsigned int word = s0 & 3U;
    if (word == 0U || have_block == 0) {
      philox_block(s0 >> 2U);
    }
    s0 += 1U;
    return (word == 0U) ? b0 : ((word == 1U) ? b1 : ((word == 2U) ? b2 : b3));
  }

  // U[0,1) float with 24 random bits (random.h:140-166); never returns 1
  AHD float uniform() { return static_cast<float>(next_u32() >> 8U) * 0x1.0p-24F; }

  // U(0,1): rejects zero (random.h:194-201)
  AHD float uniform_pos() {
    while (true) {
      const float z = uniform();
      if (z > 0) {
        return z;
      }
    }
  }
};

// SplitMix32-seeded Xoshiro128++ state from a 32-bit seed (reference random.h:44-121), host-side helper
inline void xoshiro_seed(const unsigned int seed, unsigned int out[4]) {
  unsigned long long state = static_cast<unsigned long long>(seed) + 0x9E3779B97f4A7C15ULL;
  state = (state ^ (state >> 30U)) * 0xBF58476D1CE4E5B9ULL;
  state = (state ^ (state >> 27U)) * 0x94D049BB133111EBULL;
  auto s = static_cast<unsigned int>(state ^ (state >> 31U));
  for (int i = 0; i < 4; i++) {
    unsigned int r = (s += 0x9e3779b9U);
    r = (r ^ (r >> 16U)) * 0x21f0aaadU;
    r = (r ^ (r >> 15U)) * 0x735a2d97U;
    out[i] = r ^ (r >> 15U);
  }
}

}  // namespace ab
