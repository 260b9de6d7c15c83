// Device-side building blocks shared by the sampler kernels: Philox counter RNG, the draw transforms, and the
// bucket-accelerated / warp-cooperative searches over the trajectory boundary table.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ogb {

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  key = seed, counter = (row, batch_lo, batch_hi, purpose | stream<<8).
// A draw is a pure function of (seed, stream, batch index, row, purpose): no RNG state lives in memory, any
// kernel can regenerate any draw, and the sampler's whole RNG state is one 64-bit batch counter.
// oracle/philox_np.py restates exactly this for the tests.
// ---------------------------------------------------------------------------------------------------------
enum Purpose : uint32_t {
  PURPOSE_IDX = 0,       // .x,.y -> transition index position; .z,.w -> value-goal mix coins (u_traj, u_cur)
  PURPOSE_GOAL = 1,      // .x,.y -> the value goal's 64 bits; .z,.w -> the actor goal's 64 bits (a goal spends its bits
                         //          either on the random-goal position or on the geometric / distance uniform)
  PURPOSE_GOAL_LOW = 2,  // .x,.y -> the low-value goal's 64 bits; .z,.w -> its mix coins (HGC + low_discount)
  PURPOSE_CROP = 3,      // .x -> crop shift (cy, cx jointly); only drawn for augmented batches
  PURPOSE_MIX = 4,       // .x,.y -> actor-goal mix coins (u_traj, u_cur); only drawn when that mix is a real choice
  PURPOSE_TRL_MID = 6,   // .x,.y -> TRL midpoint position in [idx, value goal)
  PURPOSE_COIN = 7       // row = 0xFFFFFFFF: .x,.y -> the per-batch augmentation coin
};

// The ten round keys (key + i * Weyl constant) are the same for every thread: the host computes them once and the
// kernel reads them from the constant bank, which removes two adds per round from every draw.
struct RngKey {
  uint32_t round_key[10][2];
  uint32_t stream;
  uint32_t pad_;
};

__host__ inline RngKey make_rng_key(uint64_t seed, uint32_t stream) {
  RngKey k;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    k.round_key[i][0] = k0;
    k.round_key[i][1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  k.stream = stream;
  k.pad_ = 0;
  return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const RngKey& key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;   // one IMAD.WIDE each
    c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ key.round_key[round][0], (uint32_t)p1,
                   (uint32_t)(p0 >> 32) ^ c.w ^ key.round_key[round][1], (uint32_t)p0);
  }
  return c;
}

__device__ __forceinline__ uint4 draw4(const RngKey& key, uint64_t batch, uint32_t row, uint32_t purpose) {
  const uint4 ctr = make_uint4(row, (uint32_t)batch, (uint32_t)(batch >> 32), purpose | (key.stream << 8));
  return philox4x32_10(ctr, key);
}

// the same draw, out of line: for the draws that only some configurations make (actor mix coins, TRL midpoints, crop
// shifts, the per-batch coin), so that their ten Philox rounds are not inlined into the hot path's instruction footprint
__device__ __noinline__ uint4 draw4_cold(const RngKey& key, uint64_t batch, uint32_t row, uint32_t purpose) {
  return draw4(key, batch, row, purpose);
}

// uniform integer in [0, n), n < 2^32, from 64 random bits: floor(((hi << 32) | lo) * n / 2^64), bias < n / 2^64
__device__ __forceinline__ uint32_t bounded_u32n(uint32_t hi, uint32_t lo, uint32_t n) {
  const uint64_t t = (uint64_t)lo * n;
  const uint64_t u = (uint64_t)hi * n + (t >> 32);
  return (uint32_t)(u >> 32);
}

// uniform double in [0, 1) with 53 random bits, the same construction numpy's legacy rand() uses on two words
__device__ __forceinline__ double unit_double(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// geometric(p) by inversion on U = unit_double(a, b), support [1, inf): ceil(log(1-U) / log(1-p)); log_1mp = log(1-p)
// comes from the host.  A float32 estimate decides unless the quotient q lies so close to an integer that the
// estimate's error could change the ceiling; only then (about one row in 2,500) is the float64 expression
// evaluated.  Both branches return the value of the float64 expression.  Error budget of the estimate, with
// uf = float(a) * 2^-32:  |uf - U| <= 2^-27 + 2^-24 U  ->  |dq| <= q * 6e-8 / (1-U) + 7.5e-9 / ((1-U) |log_1mp|);
// log1pf 1 ulp, __fdividef 2 ulp, float(log_1mp) 0.5 ulp -> 4.2e-7 * q.  The margin below is twice that bound.
// the rare float64 path, kept out of line: its log() expansion is ~200 instructions and would otherwise be inlined at
// every goal set (the index code has to stay inside the 32 KB instruction cache level to issue well)
__device__ __noinline__ double geometric_exact(uint32_t a, uint32_t b, double log_1mp) {
  return ceil(log(1.0 - unit_double(a, b)) / log_1mp);
}

__device__ __forceinline__ int64_t geometric_from_words(uint32_t a, uint32_t b, double log_1mp, float abs_margin) {
  const float uf = __uint2float_rn(a) * 2.3283064365386962891e-10f;
  const float qf = __fdividef(log1pf(-uf), (float)log_1mp);
  const float cf = ceilf(qf);
  const float inv = __fdividef(1.0f, 1.0f - uf);
  const float margin = inv * fmaf(1.0f + qf, 1.2e-7f, abs_margin) + (1.0f + qf) * 8.4e-7f;   // abs_margin = 1.5e-8 / |log_1mp|
  double x;
  if (uf < 0.9999f && cf - qf > margin && qf - (cf - 1.0f) > margin) {
    x = (double)cf;
  } else {
    x = geometric_exact(a, b, log_1mp);
  }
  return x < 1.0 ? 1 : (int64_t)x;
}

// ---------------------------------------------------------------------------------------------------------
// Searches.  `table` is sorted int32.  bucket[b] = lower_bound(table, b << shift), so a key's answer lies in
// [bucket[key>>shift], bucket[(key>>shift)+1]] -- for trajectory tables that range is 0-2 entries wide.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_bucketed(const int32_t* __restrict__ table, const int32_t* __restrict__ bucket,
                                                    int shift, int key) {
  const int b = key >> shift;
  int lo = __ldg(bucket + b), hi = __ldg(bucket + b + 1);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(table + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Warp-cooperative searchsorted: all 32 lanes of a warp resolve ONE key with a 32-ary search (each round the
// lanes probe 32 evenly spaced splitters and a ballot picks the sub-range), ~log32(n) dependent loads instead of
// log2(n).  side_right == 0: first j with table[j] >= key (np.searchsorted side='left'); 1: first j with
// table[j] > key.  Returns the same value in every lane.
template <typename T>
__device__ __forceinline__ int64_t warp_searchsorted(const T* __restrict__ table, int64_t n, T key, bool side_right) {
  const unsigned lane = threadIdx.x & 31u;
  int64_t lo = 0, hi = n;  // invariant: the answer lies in [lo, hi]
  while (lo < hi) {
    const int64_t step = (hi - lo + 31) / 32;         // lane l probes position lo + (l+1)*step - 1
    const int64_t pos = lo + (int64_t)(lane + 1) * step - 1;
    bool before = false;                              // "table[pos] sorts before the answer"; false beyond hi
    if (pos < hi) {
      const T v = table[pos];
      before = side_right ? !(key < v) : (v < key);
    }
    // the predicate is monotone over lanes, so the ballot is a run of ones followed by zeros
    const int n_before = __popc(__ballot_sync(0xffffffffu, before));
    const int64_t cap = lo + (int64_t)(n_before + 1) * step - 1;
    lo = lo + (int64_t)n_before * step;
    hi = cap < hi ? cap : hi;                         // new span <= step - 1, so step == 1 ends the loop
  }
  return lo;
}

}  // namespace ogb
