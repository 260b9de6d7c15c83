// The index kernel and the row-gather kernel of the replay sampler.
//
// relabel_index_kernel (one thread per batch row) restates the reference's per-row index algebra
// (impls/utils/datasets.py:296-327 sample_goals, :478-491 compute_high_next_idxs, :250-252 and :533-582
// rewards/masks) from either injected draws (validation mode) or Philox draws.  It writes the
// scalar keys (masks, rewards, offsets, steps) and one int32 row-index vector per slot.
//
// gather_rows_kernel<V> (one warp per 32 batch rows) gathers the dataset rows those vectors name
// (datasets.py:78-83 get_subset, :341-357 get_observations / get_goal_observations) into the dense output arrays.
// The two are separate launches so that the gather runs at its own (high) occupancy: the index vectors are
// 4 bytes per row and slot and stay in L2 between the launches.
#pragma once
#include "device_common.cuh"

namespace ogb {

constexpr int kMaxSlots = 10;
constexpr int kMaxRowJobs = 24;
constexpr int kRelabelThreads = 256;
constexpr int kGatherMinBlocks = 4;   // resident CTAs per SM the gather kernel is compiled for (register cap 64)

// index-vector slots
enum : int {
  SLOT_IDX = 0, SLOT_NEXT = 1,
  GC_VALUE_GOAL = 2, GC_ACTOR_GOAL = 3, GC_NUM_SLOTS = 4,
  GC_TRL_MID = 4, GC_TRL_PLUS1 = 5, GC_TRL_NUM_SLOTS = 6,
  HGC_HV_GOAL = 2, HGC_HV_NEXT = 3, HGC_LV_NEXT = 4, HGC_HA_GOAL = 5, HGC_HA_NEXT = 6, HGC_LA_GOAL = 7,
  HGC_LA_NEXT = 8, HGC_LV_GOAL = 9, HGC_NUM_SLOTS = 10
};

struct GoalSpec {
  double thr_traj;   // p_trajgoal / (1.0 - p_curgoal), float64 as the reference evaluates it (datasets.py:321)
  double p_cur;
  double log_1mp;    // log(1 - (1 - discount)) for the geometric inversion
  int32_t geom;      // geometric (1) or uniform-in-remainder (0) future goals
  int32_t cur_only;  // p_curgoal == 1.0 short-circuit (datasets.py:317-318)
};

struct GoalInject {
  const int64_t* rand_pos;
  const int64_t* offset;
  const double* dist;
  const double* u_traj;
  const double* u_cur;
};

struct RowJob {
  const uint8_t* src;   // field base in HBM
  uint8_t* dst;         // dense output [total_rows, row_bytes]
  uint32_t src_stride;  // padded row stride of the resident copy
  uint32_t row_bytes;   // dense output row size
  uint16_t epr;         // elements (of 1 << vec_log2 bytes) per row
  uint16_t n_coliter;   // ceil(epr / lanes-per-row)
  uint8_t vec_log2;     // element = 1,2,4,8,16 bytes
  uint8_t lpr_log2;     // lanes per row = 1 << lpr_log2 (<= 32)
  uint8_t slot;         // which index vector names the source rows
  uint8_t pad_;
};

// A field whose whole row is at most 16 bytes (terminals, valids, 2-D observations ...) is gathered by the index
// kernel itself: the thread that computed row g's index vectors copies those few bytes straight away.
constexpr int kMaxTinyJobs = 16;
struct TinyJob {
  const uint8_t* src;
  uint8_t* dst;
  uint8_t slot;
  uint8_t size_log2;   // element size
  uint8_t n_elem;      // row = n_elem elements (<= 16 bytes in total)
  uint8_t pad_;
  uint32_t pad2_;
};

struct RelabelParams {
  // ---- dataset-side tables (all int32, resident) ----
  const int32_t* term;         // terminal_locs (datasets.py:186)
  const int32_t* term_bucket;  // lower_bound(term, b << term_shift)
  const int32_t* valid_table;  // valid_idxs (datasets.py:63) -- valid_mode 1
  const int32_t* gap_c;        // valid_mode 2: c[m] = (m-th invalid row) - m; valid_idxs[j] = j + #{m: c[m] <= j}
  const int32_t* gap_bucket;
  int64_t n_choices;           // len(valid_idxs), or size when the dataset has no 'valids'
  int32_t n_rows_ds;
  int32_t term_shift;
  int32_t gap_shift;
  int32_t valid_mode;          // 0: no 'valids'; 1: table; 2: gap ranks
  // ---- sampler config ----
  GoalSpec goal[3];            // value, low-value, actor
  const double* neg_lut;       // -(1 - discount**s)/(1 - discount)
  const double* pow_lut;       // discount**s
  int32_t kind;                // 0 GC, 1 HGC, 2 PLAIN, 3 ATC
  int32_t next_offset;         // SLOT_NEXT = idx + next_offset (1; the temporal offset k for ATC)
  int32_t trl;                 // TRL branch of GCDataset.sample (datasets.py:254-276)
  const int64_t* in_trl_mid;   // injected randint(idxs, value_goal_idxs)
  int64_t* trl_offsets;
  int64_t* trl_mid_offsets;
  int32_t has_low_goal;
  int32_t k_val, k_act, k_lo;
  int32_t gc_negative;
  int32_t stacked_next;        // frame_stack set: next_observations uses un-clamped idx+1 (datasets.py:231)
  int32_t need_mix;            // some goal set has 0 < p_cur < 1 or a real traj/random choice: draw the mix uniforms
  int32_t aug_mode;            // draw the per-batch coin (p_aug is not None and not evaluation)
  int32_t crop_pad;
  double p_aug;
  // ---- randomness ----
  RngKey key;
  uint64_t batch0;
  const int64_t* in_idx_pos;
  GoalInject in_goal[3];
  const int64_t* in_crop;
  double in_coin;
  const int64_t* given_idxs;
  // ---- launch shape ----
  int64_t batch;               // rows per sample() call
  int64_t total_rows;          // batch * n_batches (also the stride of the [slot][row] index vectors)
  int64_t row_begin, row_end;  // rows this launch handles (a big launch is split into chunks, see ogb_sampler.cu)
  int32_t n_slots;
  int32_t pad0_;
  // ---- scalar outputs (float64 / int64 like the reference) ----
  double* masks;
  double* rewards;
  int64_t* hv_offsets;
  int64_t* hv_steps;
  int64_t* lv_steps;
  double* hv_masks;
  double* hv_rewards;
  double* lv_masks;
  double* lv_rewards;
  // ---- index outputs: [slot][total_rows] ----
  int32_t* vec_rows;           // row index vectors, consumed by the gather kernels
  int32_t* vec_init;           // first row of each row's trajectory segment (frame stacking only; may be null)
  int8_t* crop_out;            // [total_rows][2] (dy, dx) or -128 when the batch is not augmented (may be null)
  // ---- rows of <= 16 bytes, copied by the index kernel ----
  int32_t n_tiny;
  int32_t write_vecs;          // 0: no later kernel needs the index vectors (everything was tiny) and debug is off
  TinyJob tiny[kMaxTinyJobs];
};

__device__ __forceinline__ int32_t valid_row(const RelabelParams& p, int64_t pos) {
  if (p.valid_mode == 0) return (int32_t)pos;
  if (p.valid_mode == 1) return __ldg(p.valid_table + pos);
  const int j = (int)pos;
  return j + lower_bound_bucketed(p.gap_c, p.gap_bucket, p.gap_shift, j + 1);  // upper_bound(c, j)
}

// `mix` carries the two 32-bit goal-mix uniforms of this goal set (Philox mode; ignored when draws are injected).
template <bool kInject>
__device__ __forceinline__ int32_t pick_goal(const RelabelParams& p, const int gs, const int32_t i, const int32_t fin,
                                             const uint64_t batch_id, const uint32_t r, const int64_t g, const uint2 mix) {
  const GoalSpec& s = p.goal[gs];
  if (s.cur_only) return i;                                  // p_curgoal == 1.0  (datasets.py:317-318)
  if (kInject) {
    const GoalInject& in = p.in_goal[gs];
    if (in.u_cur[g] < s.p_cur) return i;                     // np.where(rand < p_cur, idxs, ...)  :325
    if (!(in.u_traj[g] < s.thr_traj)) return valid_row(p, in.rand_pos[g]);   // random goal  :303,:320-322
    if (s.geom) {                                            // :309-310
      const int64_t t = (int64_t)i + in.offset[g];
      return (int32_t)(t < (int64_t)fin ? t : (int64_t)fin);
    }
    const int32_t lo = (i + 1 < fin) ? i + 1 : fin;          // :313-316 -- float64, no FMA contraction, half-even
    const double d = in.dist[g];
    return (int32_t)rint(__dadd_rn(__dmul_rn((double)lo, d), __dmul_rn((double)fin, __dsub_rn(1.0, d))));
  }
  if (unit_from_word(mix.y) < s.p_cur) return i;
  const uint4 a = draw4(p.key, batch_id, r, PURPOSE_GOAL + (uint32_t)gs);
  if (!(unit_from_word(mix.x) < s.thr_traj)) return valid_row(p, bounded_u64(a.x, a.y, (uint64_t)p.n_choices));
  const double u = unit_double(a.z, a.w);
  if (s.geom) {
    const int64_t t = (int64_t)i + geometric_from_unit(u, s.log_1mp);
    return (int32_t)(t < (int64_t)fin ? t : (int64_t)fin);
  }
  const int32_t lo = (i + 1 < fin) ? i + 1 : fin;
  return (int32_t)rint(__dadd_rn(__dmul_rn((double)lo, u), __dmul_rn((double)fin, __dsub_rn(1.0, u))));
}

// datasets.py:478-491
__device__ __forceinline__ void subgoal_step(int32_t i, int32_t fin, int32_t goal, int32_t k, int32_t& next, int32_t& s) {
  s = fin - i < k ? fin - i : k;
  const int32_t d = goal - i;
  if (0 <= d && d < s) s = d;
  next = i + s;
}

__device__ __forceinline__ int32_t trajectory_first_row(const RelabelParams& p, int32_t x) {
  // initial_locs[searchsorted(initial_locs, x, 'right') - 1] (datasets.py:361) expressed on terminal_locs:
  // with t = lower_bound(term, x) the start is 0 when t == 0, else term[t-1] + 1.
  x = x < p.n_rows_ds ? x : p.n_rows_ds - 1;
  const int t = lower_bound_bucketed(p.term, p.term_bucket, p.term_shift, x);
  return t == 0 ? 0 : __ldg(p.term + t - 1) + 1;
}

__device__ __forceinline__ void put_slot(const RelabelParams& p, int32_t* sr, const int slot, const int64_t g, const int32_t x) {
  sr[slot] = x;  // `slot` is a compile-time constant at every call site, so sr[] stays in registers
  if (p.write_vecs) p.vec_rows[(int64_t)slot * p.total_rows + g] = x;
  if (p.vec_init != nullptr) p.vec_init[(int64_t)slot * p.total_rows + g] = trajectory_first_row(p, x);
}

__device__ __forceinline__ int32_t pick_slot(const int32_t* sr, const int slot) {
  int32_t x = sr[0];
#pragma unroll
  for (int v = 1; v < kMaxSlots; ++v) x = (slot == v) ? sr[v] : x;
  return x;
}

template <bool kInject>
__global__ void __launch_bounds__(kRelabelThreads) relabel_index_kernel(const __grid_constant__ RelabelParams p) {
  for (int64_t g = p.row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < p.row_end; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t kb = g / p.batch;
    const uint32_t r = (uint32_t)(g - kb * p.batch);
    const uint64_t batch_id = p.batch0 + (uint64_t)kb;
    int32_t sr[kMaxSlots];
#pragma unroll
    for (int v = 0; v < kMaxSlots; ++v) sr[v] = 0;

    uint4 w0 = make_uint4(0, 0, 0, 0);
    if (!kInject) w0 = draw4(p.key, batch_id, r, PURPOSE_IDX);
    int32_t i;
    if (p.given_idxs != nullptr) {
      i = (int32_t)p.given_idxs[g];
    } else {
      const int64_t pos = kInject ? p.in_idx_pos[g] : bounded_u64(w0.x, w0.y, (uint64_t)p.n_choices);
      i = valid_row(p, pos);                                            // datasets.py:65-70
    }
    put_slot(p, sr, SLOT_IDX, g, i);
    const int32_t nxt = p.stacked_next ? i + p.next_offset                                           // :231 / :408
                                       : (i + p.next_offset < p.n_rows_ds ? i + p.next_offset : p.n_rows_ds - 1);  // :82
    put_slot(p, sr, SLOT_NEXT, g, nxt);

    if (p.kind == 0 || p.kind == 1) {
      uint4 mix = make_uint4(0, 0, 0, 0);
      if (!kInject && p.need_mix) mix = draw4(p.key, batch_id, r, PURPOSE_MIX);
      const int tl = lower_bound_bucketed(p.term, p.term_bucket, p.term_shift, i);
      const int32_t fin = __ldg(p.term + tl);                            // final_state_idxs  :306,:505
      const double neg = p.gc_negative ? 1.0 : 0.0;
      if (p.kind == 0) {
        const int32_t vg = pick_goal<kInject>(p, 0, i, fin, batch_id, r, g, make_uint2(mix.x, mix.y));
        const int32_t ag = pick_goal<kInject>(p, 2, i, fin, batch_id, r, g, make_uint2(mix.z, mix.w));
        put_slot(p, sr, GC_VALUE_GOAL, g, vg);
        put_slot(p, sr, GC_ACTOR_GOAL, g, ag);
        const double succ = (i == vg) ? 1.0 : 0.0;                       // :250-252
        p.masks[g] = 1.0 - succ;
        p.rewards[g] = succ - neg;
        if (p.trl) {                                                     // :259-267
          const int64_t span = vg > i ? (int64_t)vg - i : 1;             // the reference asserts idxs != value_goal_idxs
          int32_t mid;
          if (kInject) {
            mid = (int32_t)p.in_trl_mid[g];
          } else {
            const uint4 t = draw4(p.key, batch_id, r, PURPOSE_TRL_MID);
            mid = i + (int32_t)bounded_u64(t.x, t.y, (uint64_t)span);    // randint(idxs, value_goal_idxs): [i, vg)
          }
          put_slot(p, sr, GC_TRL_MID, g, mid);
          put_slot(p, sr, GC_TRL_PLUS1, g, i + 1);
          p.trl_offsets[g] = (int64_t)vg - i;
          p.trl_mid_offsets[g] = (int64_t)mid - i;
        }
      } else {
        const int32_t hv = pick_goal<kInject>(p, 0, i, fin, batch_id, r, g, make_uint2(mix.x, mix.y));   // :508-514
        int32_t hv_next, hv_s, lv_next, lv_s;
        subgoal_step(i, fin, hv, p.k_val, hv_next, hv_s);                             // :519-524
        subgoal_step(i, fin, hv, p.k_lo, lv_next, lv_s);                              // :544-549
        put_slot(p, sr, HGC_HV_GOAL, g, hv);
        put_slot(p, sr, HGC_HV_NEXT, g, hv_next);
        put_slot(p, sr, HGC_LV_NEXT, g, lv_next);
        p.hv_offsets[g] = (int64_t)hv - (int64_t)i;                                   // :531
        p.hv_steps[g] = hv_s;
        p.lv_steps[g] = lv_s;
        const double hv_succ = hv_s < p.k_val ? 1.0 : 0.0;                            // :533
        const double lv_succ = lv_s < p.k_lo ? 1.0 : 0.0;                             // :552
        p.hv_masks[g] = 1.0 - hv_succ;
        p.hv_rewards[g] = p.gc_negative ? __ldg(p.neg_lut + hv_s) : __dmul_rn(__ldg(p.pow_lut + hv_s), hv_succ);
        double lv_mask = 1.0 - lv_succ;
        double lv_rew = p.gc_negative ? __ldg(p.neg_lut + lv_s) : __dmul_rn(__ldg(p.pow_lut + lv_s), lv_succ);
        if (p.has_low_goal) {                                                         // :563-576
          uint4 mix_low = make_uint4(0, 0, 0, 0);
          if (!kInject) mix_low = draw4(p.key, batch_id, r, PURPOSE_MIX_LOW);
          const int32_t lvg = pick_goal<kInject>(p, 1, i, fin, batch_id, r, g, make_uint2(mix_low.x, mix_low.y));
          put_slot(p, sr, HGC_LV_GOAL, g, lvg);
          const double s = (i == lvg) ? 1.0 : 0.0;
          lv_mask = 1.0 - s;
          lv_rew = s - neg;
        }
        p.lv_masks[g] = lv_mask;
        p.lv_rewards[g] = lv_rew;
        const double succ = (i == hv) ? 1.0 : 0.0;                                    // :579-582
        p.masks[g] = 1.0 - succ;
        p.rewards[g] = succ - neg;
        const int32_t ha = pick_goal<kInject>(p, 2, i, fin, batch_id, r, g, make_uint2(mix.z, mix.w));   // :585-591
        int32_t ha_next, la_next, unused;
        subgoal_step(i, fin, ha, p.k_act, ha_next, unused);                           // :595-600
        subgoal_step(i, fin, ha, p.k_lo, la_next, unused);                            // :613-618
        const int64_t la = (int64_t)i + p.k_act;                                      // :610
        put_slot(p, sr, HGC_HA_GOAL, g, ha);
        put_slot(p, sr, HGC_HA_NEXT, g, ha_next);
        put_slot(p, sr, HGC_LA_GOAL, g, (int32_t)(la < (int64_t)fin ? la : (int64_t)fin));
        put_slot(p, sr, HGC_LA_NEXT, g, la_next);
      }
    }

    // rows of <= 16 bytes: this thread copies them now (datasets.py:78-83 for the per-transition fields)
#pragma unroll 1
    for (int j = 0; j < p.n_tiny; ++j) {
      const TinyJob& job = p.tiny[j];
      const int row_bytes = (int)job.n_elem << job.size_log2;
      const uint8_t* sp = job.src + (size_t)pick_slot(sr, job.slot) * row_bytes;
      uint8_t* dp = job.dst + (size_t)g * row_bytes;
      switch (job.size_log2) {
        case 4: *reinterpret_cast<uint4*>(dp) = __ldg(reinterpret_cast<const uint4*>(sp)); break;
        case 3:
          for (int e = 0; e < job.n_elem; ++e) reinterpret_cast<uint2*>(dp)[e] = __ldg(reinterpret_cast<const uint2*>(sp) + e);
          break;
        case 2:
          for (int e = 0; e < job.n_elem; ++e) reinterpret_cast<uint32_t*>(dp)[e] = __ldg(reinterpret_cast<const uint32_t*>(sp) + e);
          break;
        case 1:
          for (int e = 0; e < job.n_elem; ++e) reinterpret_cast<uint16_t*>(dp)[e] = __ldg(reinterpret_cast<const uint16_t*>(sp) + e);
          break;
        default:
          for (int e = 0; e < job.n_elem; ++e) dp[e] = __ldg(sp + e);
          break;
      }
    }

    if (p.crop_out != nullptr) {
      int dy = -128, dx = -128;
      if (p.aug_mode) {                                                               // :278-279, :621-622
        double coin;
        if (kInject) {
          coin = p.in_coin;
        } else {
          const uint4 c = draw4(p.key, batch_id, 0xFFFFFFFFu, PURPOSE_COIN);
          coin = unit_double(c.x, c.y);
        }
        if (coin < p.p_aug) {                                                         // :333
          const uint32_t span = 2u * (uint32_t)p.crop_pad + 1u;
          const uint32_t joint = __umulhi(w0.z, span * span);            // (cy, cx) jointly uniform on span x span
          const int cy = kInject ? (int)p.in_crop[2 * g] : (int)(joint / span);
          const int cx = kInject ? (int)p.in_crop[2 * g + 1] : (int)(joint % span);
          dy = cy - p.crop_pad;
          dx = cx - p.crop_pad;
        }
      }
      p.crop_out[2 * g] = (int8_t)dy;
      p.crop_out[2 * g + 1] = (int8_t)dx;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Row gather.  Each warp owns 32 consecutive batch rows and walks the job list; lane l holds the source row of
// batch row l for the current job and broadcasts it with a shuffle.  A warp-wide access covers 32 >> lpr_log2
// rows of (1 << lpr_log2) elements (short rows) or one 32-element slice of one row (long rows).  Every warp
// issues kU independent loads before the matching stores so that enough bytes are in flight to cover HBM
// latency; there is no block-level synchronisation at all.
// ---------------------------------------------------------------------------------------------------------
struct GatherParams {
  const int32_t* vec_rows;     // [slot][total_rows]
  int64_t total_rows;
  int64_t row_begin, row_end;  // row_begin is a multiple of 32
  int32_t n_jobs;
  RowJob jobs[kMaxRowJobs];
};

// read-once data: do not allocate in L1 (keeps the small index / table lines resident)
__device__ __forceinline__ uint32_t load_stream(const uint32_t* p) {
  uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ uint2 load_stream(const uint2* p) {
  uint2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v;
}
__device__ __forceinline__ uint4 load_stream(const uint4* p) {
  uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}
__device__ __forceinline__ uint16_t load_stream(const uint16_t* p) { return __ldg(p); }
__device__ __forceinline__ uint8_t load_stream(const uint8_t* p) { return __ldg(p); }

template <typename V>
struct GatherTuning {
  static constexpr int kU = sizeof(V) >= 16 ? 4 : 8;  // independent loads in flight per warp
};

template <typename V>
__global__ void __launch_bounds__(kRelabelThreads, kGatherMinBlocks) gather_rows_kernel(const __grid_constant__ GatherParams p) {
  constexpr int kU = GatherTuning<V>::kU;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps_global = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_warp_tiles = (p.row_end + 31) >> 5;

#pragma unroll 1
  for (int64_t wt = (p.row_begin >> 5) + warp_global; wt < n_warp_tiles; wt += n_warps_global) {
    const int64_t g0 = wt << 5;
    const int n = (int)(p.row_end - g0 < 32 ? p.row_end - g0 : 32);
    // lane l holds the source row of batch row g0 + l; the vector of the next job is fetched one job ahead
    int32_t next_row = lane < n ? __ldg(p.vec_rows + (int64_t)p.jobs[0].slot * p.total_rows + g0 + lane) : 0;
#pragma unroll 1
    for (int j = 0; j < p.n_jobs; ++j) {
      const RowJob& job = p.jobs[j];
      const int32_t my_row = next_row;
      if (j + 1 < p.n_jobs) next_row = lane < n ? __ldg(p.vec_rows + (int64_t)p.jobs[j + 1].slot * p.total_rows + g0 + lane) : 0;
      const uint8_t* __restrict__ src = job.src;
      uint8_t* __restrict__ dst = job.dst + (size_t)g0 * job.row_bytes;
      const uint32_t stride = job.src_stride, row_bytes = job.row_bytes;
      const int epr = job.epr;
      if (job.n_coliter == 1) {
        const int lpr_log2 = job.lpr_log2;
        const int sub = lane >> lpr_log2;                       // row within the pass
        const int col = lane & ((1 << lpr_log2) - 1);
        const int rpp_log2 = 5 - lpr_log2;
        const int n_pass = (n + (1 << rpp_log2) - 1) >> rpp_log2;
        const bool col_ok = col < epr;
        // lanes beyond the row (col >= epr) and rows beyond the tile re-read a valid element instead of being
        // predicated off: no extra sectors are touched and every val[] register is unconditionally defined
        const uint32_t src_off = (uint32_t)(col_ok ? col : epr - 1) * (uint32_t)sizeof(V);
        const uint32_t dst_off = (uint32_t)sub * row_bytes + (uint32_t)col * (uint32_t)sizeof(V);
        const uint32_t pass_bytes = row_bytes << rpp_log2;
#pragma unroll 1
        for (int pass0 = 0; pass0 < n_pass; pass0 += kU) {
          V val[kU];
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const int row = min(((pass0 + u) << rpp_log2) + sub, n - 1);
            const uint32_t src_row = (uint32_t)__shfl_sync(0xffffffffu, my_row, row);
            val[u] = load_stream(reinterpret_cast<const V*>(src + (size_t)src_row * stride + src_off));
          }
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const int row = ((pass0 + u) << rpp_log2) + sub;
            if (col_ok && row < n) *reinterpret_cast<V*>(dst + (uint32_t)(pass0 + u) * pass_bytes + dst_off) = val[u];
          }
        }
      } else {
        // long rows: 2 rows x up to 4 column slices per sweep
        constexpr int kRowsPerSweep = 2, kSlices = (kU > 8 ? 8 : kU) / kRowsPerSweep;
        const int n_coliter = job.n_coliter;
#pragma unroll 1
        for (int row0 = 0; row0 < n; row0 += kRowsPerSweep) {
          const uint8_t* sp[kRowsPerSweep];
#pragma unroll
          for (int rr = 0; rr < kRowsPerSweep; ++rr)
            sp[rr] = src + (size_t)(uint32_t)__shfl_sync(0xffffffffu, my_row, min(row0 + rr, n - 1)) * stride;
#pragma unroll 1
          for (int c0 = 0; c0 < n_coliter; c0 += kSlices) {
            V val[kRowsPerSweep][kSlices];
#pragma unroll
            for (int rr = 0; rr < kRowsPerSweep; ++rr)
#pragma unroll
              for (int c = 0; c < kSlices; ++c) {
                const int col = min(((c0 + c) << 5) + lane, epr - 1);
                val[rr][c] = load_stream(reinterpret_cast<const V*>(sp[rr]) + col);
              }
#pragma unroll
            for (int rr = 0; rr < kRowsPerSweep; ++rr)
#pragma unroll
              for (int c = 0; c < kSlices; ++c) {
                const int col = ((c0 + c) << 5) + lane;
                if (row0 + rr < n && col < epr) reinterpret_cast<V*>(dst + (size_t)(row0 + rr) * row_bytes)[col] = val[rr][c];
              }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Row gather, asynchronous variant (the default for rows of 17 .. kAsyncMaxStride bytes).
//
// The register-staged kernel above keeps one register per in-flight load, which caps the bytes a warp can have
// in flight.  Here the source rows go HBM -> shared memory with cp.async (LDGSTS, 16 bytes per lane, L1
// bypassed) into a per-warp ring of kAsyncStages stages, so several KB per warp are in flight without holding
// registers, and the drain is shared-memory loads + *flat* coalesced stores: the dense output tile of a job is
// one contiguous span, so lanes store consecutive elements no matter how long a row is.
// Work item = (warp tile of 32 batch rows, job, sub-range of rows that fits one stage).
// ---------------------------------------------------------------------------------------------------------
constexpr int kAsyncStages = 3;
constexpr int kAsyncWarps = 8;
constexpr int kAsyncMaxStride = 4096;

struct AsyncJob {
  const uint8_t* src;
  uint8_t* dst;
  uint32_t stride;        // resident row stride, multiple of 16
  uint32_t row_bytes;
  uint32_t cpr;           // 16-byte chunks copied per row = ceil(row_bytes / 16)
  uint32_t cpr_magic;     // ceil(2^32 / cpr), or 0 when cpr == 1: e / cpr == umulhi(e, magic) for e * cpr < 2^32
  uint32_t epr;           // output elements per row (row_bytes >> vec_log2)
  uint32_t epr_magic;
  uint16_t rows_per_item; // rows of one stage
  uint8_t vec_log2;
  uint8_t slot;
};

struct AsyncGatherParams {
  const int32_t* vec_rows;
  int64_t total_rows;
  int64_t row_begin, row_end;  // row_begin is a multiple of 32
  int32_t n_jobs;
  int32_t stage_bytes;
  int32_t flat_drain;          // 1: flat element index (coalesced 128-byte stores), 0: row by row
  int32_t pad_;
  AsyncJob jobs[kMaxRowJobs];
};

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t fast_div(uint32_t e, uint32_t magic) { return magic ? __umulhi(e, magic) : e; }

// flat drain: the item's dense output is one contiguous span, lanes store consecutive elements (full 128-byte lines)
template <typename V>
__device__ __forceinline__ void drain_flat(const uint8_t* __restrict__ sbase, uint8_t* __restrict__ dbase, const uint32_t n_elem,
                                           const uint32_t stride, const uint32_t epr, const uint32_t epr_magic, const int lane) {
  for (uint32_t e = lane; e < n_elem; e += 32) {
    const uint32_t r = fast_div(e, epr_magic), col = e - r * epr;
    reinterpret_cast<V*>(dbase)[e] = *reinterpret_cast<const V*>(sbase + r * stride + col * (uint32_t)sizeof(V));
  }
}

// 4-byte elements, the common case (float32 rows whose size is not a multiple of 16): each lane assembles 16
// consecutive output bytes from four shared-memory words (they straddle at most one row boundary) and issues one
// 16-byte store, so a warp-wide store covers 512 contiguous bytes of the dense output.
__device__ __forceinline__ void drain_flat_quads(const uint8_t* __restrict__ sbase, uint8_t* __restrict__ dbase, const uint32_t n_words,
                                                 const uint32_t stride, const uint32_t epr, const uint32_t epr_magic, const int lane) {
  const uint32_t n_quads = n_words >> 2;
  const uint32_t row_gap = stride - epr * 4u;             // padding bytes between two rows in the stage
  for (uint32_t q = lane; q < n_quads; q += 32) {
    const uint32_t w0 = q << 2;
    const uint32_t r = fast_div(w0, epr_magic);
    uint32_t col = w0 - r * epr;
    uint32_t off = r * stride + col * 4u;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = *reinterpret_cast<const uint32_t*>(sbase + off);
      off += 4u;
      if (++col == epr) { col = 0; off += row_gap; }
    }
    reinterpret_cast<uint4*>(dbase)[q] = make_uint4(v[0], v[1], v[2], v[3]);
  }
  for (uint32_t e = (n_quads << 2) + lane; e < n_words; e += 32) {   // ragged last tile only
    const uint32_t r = fast_div(e, epr_magic), col = e - r * epr;
    reinterpret_cast<uint32_t*>(dbase)[e] = *reinterpret_cast<const uint32_t*>(sbase + r * stride + col * 4u);
  }
}

template <typename V>
__device__ __forceinline__ void drain_rows(const uint8_t* __restrict__ sbase, uint8_t* __restrict__ dbase, const int rows,
                                           const uint32_t stride, const uint32_t row_bytes, const uint32_t epr, const int lane) {
  if (epr <= 32) {
    const bool ok = (uint32_t)lane < epr;
    const uint32_t off = (uint32_t)lane * (uint32_t)sizeof(V);
    int r = 0;
    for (; r + 4 <= rows; r += 4) {
      V v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (ok) v[q] = *reinterpret_cast<const V*>(sbase + (uint32_t)(r + q) * stride + off);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (ok) *reinterpret_cast<V*>(dbase + (uint32_t)(r + q) * row_bytes + off) = v[q];
    }
    for (; r < rows; ++r)
      if (ok) *reinterpret_cast<V*>(dbase + (uint32_t)r * row_bytes + off) = *reinterpret_cast<const V*>(sbase + (uint32_t)r * stride + off);
  } else {
    for (int r = 0; r < rows; ++r)
      for (uint32_t c = lane; c < epr; c += 32)
        reinterpret_cast<V*>(dbase + (uint32_t)r * row_bytes)[c] = reinterpret_cast<const V*>(sbase + (uint32_t)r * stride)[c];
  }
}

struct ItemCursor {
  int64_t wt;   // warp tile
  int32_t j;    // job
  int32_t sub;  // first row of the sub-range within the warp tile
};

__global__ void __launch_bounds__(kAsyncWarps * 32) gather_rows_async_kernel(const __grid_constant__ AsyncGatherParams p) {
  extern __shared__ __align__(128) uint8_t smem_ring[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * kAsyncWarps + warp;
  const int64_t n_warps_global = (int64_t)gridDim.x * kAsyncWarps;
  const int64_t n_warp_tiles = (p.row_end + 31) >> 5;
  const int64_t first_tile = (p.row_begin >> 5) + warp_global;
  uint8_t* ring = smem_ring + (size_t)warp * kAsyncStages * p.stage_bytes;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);

  auto advance = [&](ItemCursor& c) {
    c.sub += p.jobs[c.j].rows_per_item;
    if (c.sub >= 32) {
      c.sub = 0;
      if (++c.j == p.n_jobs) { c.j = 0; c.wt += n_warps_global; }
    }
  };
  // lane l keeps the source row of batch row (wt*32 + l) for the (tile, job) pair being issued; the vector of the
  // pair after it is fetched one pair ahead, so its L2/HBM latency is hidden behind a whole job of work
  auto load_rows = [&](int64_t wt, int j) -> int32_t {
    if (wt >= n_warp_tiles) return 0;
    const int64_t g0 = wt << 5;
    const int n = (int)(p.row_end - g0 < 32 ? p.row_end - g0 : 32);
    return __ldg(p.vec_rows + (int64_t)p.jobs[j].slot * p.total_rows + g0 + min(lane, n - 1));
  };
  int32_t issue_rows = load_rows(first_tile, 0);
  int64_t pref_wt = p.n_jobs > 1 ? first_tile : first_tile + n_warps_global;
  int32_t pref_job = p.n_jobs > 1 ? 1 : 0;
  int32_t pref_rows = load_rows(pref_wt, pref_job);
  int64_t issue_rows_wt = first_tile;
  int32_t issue_rows_job = 0;

  auto issue = [&](const ItemCursor& c, int stage) {
    if (c.wt < n_warp_tiles) {
      const AsyncJob& job = p.jobs[c.j];
      const int64_t g0 = c.wt << 5;
      const int n = (int)(p.row_end - g0 < 32 ? p.row_end - g0 : 32);
      if (issue_rows_wt != c.wt || issue_rows_job != c.j) {   // entering the next pair: rotate the prefetched vector in
        issue_rows = pref_rows;
        issue_rows_wt = c.wt;
        issue_rows_job = c.j;
        pref_job = c.j + 1 < p.n_jobs ? c.j + 1 : 0;
        pref_wt = pref_job ? c.wt : c.wt + n_warps_global;
        pref_rows = load_rows(pref_wt, pref_job);
      }
      const int rows = min((int)job.rows_per_item, n - c.sub);     // may be <= 0 for the ragged last tile
      const uint32_t n_chunks = rows > 0 ? (uint32_t)rows * job.cpr : 0u;
      const uint32_t base = ring_u32 + (uint32_t)stage * (uint32_t)p.stage_bytes;
      for (uint32_t e0 = 0; e0 < n_chunks; e0 += 32) {
        const uint32_t e = e0 + lane;
        const uint32_t r = min(fast_div(e, job.cpr_magic), (uint32_t)rows - 1);
        const int32_t src_row = __shfl_sync(0xffffffffu, issue_rows, c.sub + (int)r);
        if (e < n_chunks) {
          const uint32_t ch = e - r * job.cpr;
          cp_async16(base + r * job.stride + (ch << 4), job.src + (size_t)(uint32_t)src_row * job.stride + (ch << 4));
        }
      }
    }
    cp_async_commit();
  };

  auto drain = [&](const ItemCursor& c, int stage) {
    const AsyncJob& job = p.jobs[c.j];
    const int64_t g0 = c.wt << 5;
    const int n = (int)(p.row_end - g0 < 32 ? p.row_end - g0 : 32);
    const int rows = min((int)job.rows_per_item, n - c.sub);
    if (rows <= 0) return;
    const uint8_t* sbase = ring + (size_t)stage * p.stage_bytes;
    uint8_t* dbase = job.dst + (size_t)(g0 + c.sub) * job.row_bytes;
    if (p.flat_drain) {
      const uint32_t n_elem = (uint32_t)rows * job.epr;
      switch (job.vec_log2) {
        case 4: drain_flat<uint4>(sbase, dbase, n_elem, job.stride, job.epr, job.epr_magic, lane); break;
        case 3: drain_flat<uint2>(sbase, dbase, n_elem, job.stride, job.epr, job.epr_magic, lane); break;
        case 2:
          if ((reinterpret_cast<uintptr_t>(dbase) & 15) == 0) drain_flat_quads(sbase, dbase, n_elem, job.stride, job.epr, job.epr_magic, lane);
          else drain_flat<uint32_t>(sbase, dbase, n_elem, job.stride, job.epr, job.epr_magic, lane);
          break;
        case 1: drain_flat<uint16_t>(sbase, dbase, n_elem, job.stride, job.epr, job.epr_magic, lane); break;
        default: drain_flat<uint8_t>(sbase, dbase, n_elem, job.stride, job.epr, job.epr_magic, lane); break;
      }
      return;
    }
    // one warp-wide shared load + global store per 32 elements of a row; rows are unrolled by 4 for ILP
    switch (job.vec_log2) {
      case 4: drain_rows<uint4>(sbase, dbase, rows, job.stride, job.row_bytes, job.epr, lane); break;
      case 3: drain_rows<uint2>(sbase, dbase, rows, job.stride, job.row_bytes, job.epr, lane); break;
      case 2: drain_rows<uint32_t>(sbase, dbase, rows, job.stride, job.row_bytes, job.epr, lane); break;
      case 1: drain_rows<uint16_t>(sbase, dbase, rows, job.stride, job.row_bytes, job.epr, lane); break;
      default: drain_rows<uint8_t>(sbase, dbase, rows, job.stride, job.row_bytes, job.epr, lane); break;
    }
  };

  ItemCursor head{first_tile, 0, 0};  // next item to issue
  ItemCursor tail{first_tile, 0, 0};  // next item to drain
  int head_stage = 0, tail_stage = 0;
#pragma unroll 1
  for (int s = 0; s < kAsyncStages - 1; ++s) {
    issue(head, head_stage);
    if (head.wt < n_warp_tiles) advance(head);
    head_stage = head_stage + 1 == kAsyncStages ? 0 : head_stage + 1;
  }
#pragma unroll 1
  while (tail.wt < n_warp_tiles) {
    issue(head, head_stage);
    if (head.wt < n_warp_tiles) advance(head);
    head_stage = head_stage + 1 == kAsyncStages ? 0 : head_stage + 1;
    cp_async_wait<kAsyncStages - 1>();   // everything but the newest kAsyncStages-1 groups has landed
    __syncwarp();                        // ... for every lane of this warp
    drain(tail, tail_stage);
    __syncwarp();                        // the stage may be overwritten by the next issue
    advance(tail);
    tail_stage = tail_stage + 1 == kAsyncStages ? 0 : tail_stage + 1;
  }
}

}  // namespace ogb
