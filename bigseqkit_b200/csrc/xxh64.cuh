// xxh64.cuh -- XXH64 (seed-parameterised) on the device, bit-exact with cespare/xxhash/v2
// Sum64 (seed 0), the hash behind RmDupPrepare's key (bigseqkit-lib/rmdup.go:67-86).
// The byte source is a functor so that the same code hashes raw or lower-cased subjects.
#pragma once
#include "kernels.h"

namespace bsk {

#define XXP1 11400714785074694791ULL
#define XXP2 14029467366897019727ULL
#define XXP3 1609587929392839161ULL
#define XXP4 9650029242287828579ULL
#define XXP5 2870177450012600261ULL

__host__ __device__ __forceinline__ u64 xx_rotl(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
__host__ __device__ __forceinline__ u64 xx_round(u64 acc, u64 in) {
  acc += in * XXP2;
  acc = xx_rotl(acc, 31);
  return acc * XXP1;
}
__host__ __device__ __forceinline__ u64 xx_merge(u64 acc, u64 v) {
  v = xx_round(0, v);
  acc ^= v;
  return acc * XXP1 + XXP4;
}

struct XxState {
  u64 v1, v2, v3, v4, seed;
  __host__ __device__ __forceinline__ void init(u64 s) {
    seed = s;
    v1 = s + XXP1 + XXP2;
    v2 = s + XXP2;
    v3 = s;
    v4 = s - XXP1;
  }
  __host__ __device__ __forceinline__ void stripe(u64 a, u64 b, u64 c, u64 d) {
    v1 = xx_round(v1, a);
    v2 = xx_round(v2, b);
    v3 = xx_round(v3, c);
    v4 = xx_round(v4, d);
  }
  __host__ __device__ __forceinline__ u64 converge(bool had_stripes) const {
    if (!had_stripes) return seed + XXP5;
    u64 h = xx_rotl(v1, 1) + xx_rotl(v2, 7) + xx_rotl(v3, 12) + xx_rotl(v4, 18);
    h = xx_merge(h, v1);
    h = xx_merge(h, v2);
    h = xx_merge(h, v3);
    h = xx_merge(h, v4);
    return h;
  }
};

__host__ __device__ __forceinline__ u64 xx_avalanche(u64 h) {
  h ^= h >> 33;
  h *= XXP2;
  h ^= h >> 29;
  h *= XXP3;
  h ^= h >> 32;
  return h;
}

// little-endian word readers over a byte functor get(i)
template <class Get>
__host__ __device__ __forceinline__ u64 xx_rd64(const Get &get, u32 i) {
  u64 v = 0;
#pragma unroll
  for (int b = 0; b < 8; b++) v |= (u64)get(i + b) << (8 * b);
  return v;
}
template <class Get>
__host__ __device__ __forceinline__ u64 xx_rd32(const Get &get, u32 i) {
  u64 v = 0;
#pragma unroll
  for (int b = 0; b < 4; b++) v |= (u64)get(i + b) << (8 * b);
  return v;
}

// Two hashes (two seeds) in one pass over the bytes.
template <class Get>
__host__ __device__ __forceinline__ void xxh64_pair(const Get &get, u32 len, u64 seed_a, u64 seed_b, u64 &ha, u64 &hb,
                                                    bool want_b) {
  XxState a, b;
  a.init(seed_a);
  b.init(seed_b);
  u32 p = 0;
  const bool stripes = len >= 32;
  if (stripes) {
    for (; p + 32 <= len; p += 32) {
      const u64 w0 = xx_rd64(get, p), w1 = xx_rd64(get, p + 8), w2 = xx_rd64(get, p + 16), w3 = xx_rd64(get, p + 24);
      a.stripe(w0, w1, w2, w3);
      if (want_b) b.stripe(w0, w1, w2, w3);
    }
  }
  u64 h = a.converge(stripes) + (u64)len;
  u64 g = b.converge(stripes) + (u64)len;
  for (; p + 8 <= len; p += 8) {
    const u64 w = xx_rd64(get, p);
    h ^= xx_round(0, w);
    h = xx_rotl(h, 27) * XXP1 + XXP4;
    if (want_b) {
      g ^= xx_round(0, w);
      g = xx_rotl(g, 27) * XXP1 + XXP4;
    }
  }
  if (p + 4 <= len) {
    const u64 w = xx_rd32(get, p);
    h ^= w * XXP1;
    h = xx_rotl(h, 23) * XXP2 + XXP3;
    if (want_b) {
      g ^= w * XXP1;
      g = xx_rotl(g, 23) * XXP2 + XXP3;
    }
    p += 4;
  }
  for (; p < len; p++) {
    const u64 c = get(p);
    h ^= c * XXP5;
    h = xx_rotl(h, 11) * XXP1;
    if (want_b) {
      g ^= c * XXP5;
      g = xx_rotl(g, 11) * XXP1;
    }
  }
  ha = xx_avalanche(h);
  hb = want_b ? xx_avalanche(g) : 0;
}

// ---- the 128-bit fingerprint of rmdup: {XXH64 seed 0 (the reference's key), FP64}.  FP64 is an independent 64-bit hash
// of the same bytes built for cheap evaluation next to XXH64: four multiply-xor lanes over the 32-byte stripes (lane i
// takes word i of every stripe, so four threads can share one subject) + a serial tail, folded with the length.
static const u64 kFpSeed = 0x9E3779B97F4A7C15ull;
__host__ __device__ __forceinline__ u64 fp_lane_init(u32 i) { return kFpSeed + (u64)i * XXP3; }
__host__ __device__ __forceinline__ u64 fp_lane_step(u64 g, u64 w) { return xx_rotl((g ^ w) * XXP2, 29); }
__host__ __device__ __forceinline__ u64 fp_finish(u64 g0, u64 g1, u64 g2, u64 g3, u64 t, u32 len) {
  return xx_avalanche(g0 + xx_rotl(g1, 17) + xx_rotl(g2, 31) + xx_rotl(g3, 47) + t + (u64)len * XXP4);
}
// the bytes behind the last whole stripe (p .. len): XXH64's tail steps on h, FP64's on t
template <class Get>
__host__ __device__ __forceinline__ void key_fp_tail(const Get &get, u32 p, u32 len, u64 &h, u64 &t) {
  for (; p + 8 <= len; p += 8) {
    const u64 w = xx_rd64(get, p);
    h ^= xx_round(0, w);
    h = xx_rotl(h, 27) * XXP1 + XXP4;
    t = xx_rotl((t ^ w) * XXP1, 27);
  }
  if (p + 4 <= len) {
    const u64 w = xx_rd32(get, p);
    h ^= w * XXP1;
    h = xx_rotl(h, 23) * XXP2 + XXP3;
    t = xx_rotl((t ^ w) * XXP2, 23);
    p += 4;
  }
  for (; p < len; p++) {
    const u64 c = get(p);
    h ^= c * XXP5;
    h = xx_rotl(h, 11) * XXP1;
    t = xx_rotl((t ^ c) * XXP5, 11);
  }
}
// one thread, one subject: key = XXH64 seed 0, fp = FP64
template <class Get>
__host__ __device__ __forceinline__ void xxh64_key_fp(const Get &get, u32 len, u64 &key, u64 &fp) {
  XxState a;
  a.init(0);
  u64 g0 = fp_lane_init(0), g1 = fp_lane_init(1), g2 = fp_lane_init(2), g3 = fp_lane_init(3);
  u32 p = 0;
  const bool stripes = len >= 32;
  for (; p + 32 <= len; p += 32) {
    const u64 w0 = xx_rd64(get, p), w1 = xx_rd64(get, p + 8), w2 = xx_rd64(get, p + 16), w3 = xx_rd64(get, p + 24);
    a.stripe(w0, w1, w2, w3);
    g0 = fp_lane_step(g0, w0);
    g1 = fp_lane_step(g1, w1);
    g2 = fp_lane_step(g2, w2);
    g3 = fp_lane_step(g3, w3);
  }
  u64 h = a.converge(stripes) + (u64)len, t = XXP5;
  key_fp_tail(get, p, len, h, t);
  key = xx_avalanche(h);
  fp = fp_finish(g0, g1, g2, g3, t, len);
}

}  // namespace bsk
