// relabel_rows_kernel -- the fused hot path for vector-valued fields.
//
// One CTA owns a tile of consecutive batch rows.  Phase 1 (one thread per row) restates the reference's per-row
// index algebra (impls/utils/datasets.py:296-327 sample_goals, :478-491 compute_high_next_idxs, :250-252 and
// :533-582 rewards/masks) from either injected draws (validation mode) or Philox draws,
// and leaves every index vector of the tile in shared memory.  Phase 2 (whole CTA, warp-per-row-group) gathers
// the dataset rows those vectors name (datasets.py:78-83 get_subset, :341-357 get_observations /
// get_goal_observations) into the dense output arrays.  Nothing but the final batch is written to HBM.
#pragma once
#include "device_common.cuh"

namespace ogb {

constexpr int kMaxSlots = 10;
constexpr int kMaxRowJobs = 24;
constexpr int kRelabelThreads = 256;
constexpr int kMaxTileRows = 256;
constexpr int kGatherUnroll = 4;

// index-vector slots
enum : int {
  SLOT_IDX = 0, SLOT_NEXT = 1,
  GC_VALUE_GOAL = 2, GC_ACTOR_GOAL = 3, GC_NUM_SLOTS = 4,
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
  int32_t kind;                // 0 GC, 1 HGC, 2 PLAIN
  int32_t has_low_goal;
  int32_t k_val, k_act, k_lo;
  int32_t gc_negative;
  int32_t stacked_next;        // frame_stack set: next_observations uses un-clamped idx+1 (datasets.py:231)
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
  int64_t total_rows;          // batch * n_batches
  int32_t tile_rows;
  int32_t n_slots;
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
  // ---- optional index outputs for the frame kernels / debug: [slot][total_rows] ----
  int32_t* vec_rows;
  int32_t* vec_init;
  int8_t* crop_out;            // [total_rows][2] (dy, dx) or -128 when the batch is not augmented
  // ---- row gathers ----
  int32_t n_jobs;
  int32_t total_items;
  int32_t item_start[kMaxRowJobs + 1];
  RowJob jobs[kMaxRowJobs];
};

__device__ __forceinline__ int32_t valid_row(const RelabelParams& p, int64_t pos) {
  if (p.valid_mode == 0) return (int32_t)pos;
  if (p.valid_mode == 1) return __ldg(p.valid_table + pos);
  const int j = (int)pos;
  return j + lower_bound_bucketed(p.gap_c, p.gap_bucket, p.gap_shift, j + 1);  // upper_bound(c, j)
}

template <bool kInject>
__device__ __forceinline__ int32_t pick_goal(const RelabelParams& p, const int gs, const int32_t i, const int32_t fin,
                                             const uint64_t batch_id, const uint32_t r, const int64_t g) {
  const GoalSpec& s = p.goal[gs];
  int64_t rand_pos, offset = 0;
  double dist = 0.0, u_traj = 0.0, u_cur = 0.0;
  if (kInject) {
    const GoalInject& in = p.in_goal[gs];
    rand_pos = in.rand_pos[g];
    if (s.geom) offset = in.offset[g]; else dist = in.dist[g];
    if (!s.cur_only) { u_traj = in.u_traj[g]; u_cur = in.u_cur[g]; }
  } else {
    const uint4 a = draw4(p.key, batch_id, r, PURPOSE_GOAL_A + 2u * (uint32_t)gs);
    rand_pos = bounded_u64(a.x, a.y, (uint64_t)p.n_choices);
    const double u = unit_double(a.z, a.w);
    if (s.geom) offset = geometric_from_unit(u, s.log_1mp); else dist = u;
    if (!s.cur_only) {
      const uint4 b = draw4(p.key, batch_id, r, PURPOSE_GOAL_B + 2u * (uint32_t)gs);
      u_traj = unit_double(b.x, b.y);
      u_cur = unit_double(b.z, b.w);
    }
  }
  if (s.cur_only) return i;
  if (u_cur < s.p_cur) return i;                      // np.where(rand < p_cur, idxs, ...)  :325
  if (!(u_traj < s.thr_traj)) return valid_row(p, rand_pos);  // random goal  :303,:320-322
  if (s.geom) {                                       // :309-310
    const int64_t t = (int64_t)i + offset;
    return (int32_t)(t < (int64_t)fin ? t : (int64_t)fin);
  }
  // :313-316 -- float64, separate multiply/add (no FMA contraction), round-half-even
  const int32_t lo = (i + 1 < fin) ? i + 1 : fin;
  const double x = __dadd_rn(__dmul_rn((double)lo, dist), __dmul_rn((double)fin, __dsub_rn(1.0, dist)));
  return (int32_t)rint(x);
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

template <bool kInject>
__global__ void __launch_bounds__(kRelabelThreads) relabel_rows_kernel(const __grid_constant__ RelabelParams p) {
  __shared__ int32_t s_row[kMaxSlots][kMaxTileRows];
  const int tile_rows = p.tile_rows;
  const int64_t n_tiles = (p.total_rows + tile_rows - 1) / tile_rows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t g0 = tile * tile_rows;
    const int n = (int)((p.total_rows - g0) < tile_rows ? (p.total_rows - g0) : tile_rows);

    // ------------------------------- phase 1: per-row index algebra -------------------------------
    if ((int)threadIdx.x < n) {
      const int t = threadIdx.x;
      const int64_t g = g0 + t;
      const int64_t kb = g / p.batch;
      const uint32_t r = (uint32_t)(g - kb * p.batch);
      const uint64_t batch_id = p.batch0 + (uint64_t)kb;

      uint4 w0 = make_uint4(0, 0, 0, 0);
      if (!kInject) w0 = draw4(p.key, batch_id, r, PURPOSE_IDX);
      int32_t i;
      if (p.given_idxs != nullptr) {
        i = (int32_t)p.given_idxs[g];
      } else {
        const int64_t pos = kInject ? p.in_idx_pos[g] : bounded_u64(w0.x, w0.y, (uint64_t)p.n_choices);
        i = valid_row(p, pos);                                            // datasets.py:65-70
      }
      s_row[SLOT_IDX][t] = i;
      const int32_t nxt = p.stacked_next ? i + 1 : (i + 1 < p.n_rows_ds ? i + 1 : p.n_rows_ds - 1);  // :82 / :231
      s_row[SLOT_NEXT][t] = nxt;

      if (p.kind != 2) {
        const int tl = lower_bound_bucketed(p.term, p.term_bucket, p.term_shift, i);
        const int32_t fin = __ldg(p.term + tl);                            // final_state_idxs  :306,:505
        const double neg = p.gc_negative ? 1.0 : 0.0;
        if (p.kind == 0) {
          const int32_t vg = pick_goal<kInject>(p, 0, i, fin, batch_id, r, g);
          const int32_t ag = pick_goal<kInject>(p, 2, i, fin, batch_id, r, g);
          s_row[GC_VALUE_GOAL][t] = vg;
          s_row[GC_ACTOR_GOAL][t] = ag;
          const double succ = (i == vg) ? 1.0 : 0.0;                       // :250-252
          p.masks[g] = 1.0 - succ;
          p.rewards[g] = succ - neg;
        } else {
          const int32_t hv = pick_goal<kInject>(p, 0, i, fin, batch_id, r, g);          // :508-514
          int32_t hv_next, hv_s, lv_next, lv_s;
          subgoal_step(i, fin, hv, p.k_val, hv_next, hv_s);                             // :519-524
          subgoal_step(i, fin, hv, p.k_lo, lv_next, lv_s);                              // :544-549
          s_row[HGC_HV_GOAL][t] = hv;
          s_row[HGC_HV_NEXT][t] = hv_next;
          s_row[HGC_LV_NEXT][t] = lv_next;
          p.hv_offsets[g] = (int64_t)hv - (int64_t)i;                                   // :531
          p.hv_steps[g] = hv_s;
          p.lv_steps[g] = lv_s;
          const double hv_succ = hv_s < p.k_val ? 1.0 : 0.0;                            // :533
          const double lv_succ = lv_s < p.k_lo ? 1.0 : 0.0;                             // :552
          p.hv_masks[g] = 1.0 - hv_succ;
          p.hv_rewards[g] = p.gc_negative ? __ldg(p.neg_lut + hv_s) : __dmul_rn(__ldg(p.pow_lut + hv_s), hv_succ);
          double lv_mask = 1.0 - lv_succ;
          double lv_rew = p.gc_negative ? __ldg(p.neg_lut + lv_s) : __dmul_rn(__ldg(p.pow_lut + lv_s), lv_succ);
          int32_t lvg = i;
          if (p.has_low_goal) {                                                         // :563-576
            lvg = pick_goal<kInject>(p, 1, i, fin, batch_id, r, g);
            const double s = (i == lvg) ? 1.0 : 0.0;
            lv_mask = 1.0 - s;
            lv_rew = s - neg;
          }
          s_row[HGC_LV_GOAL][t] = lvg;
          p.lv_masks[g] = lv_mask;
          p.lv_rewards[g] = lv_rew;
          const double succ = (i == hv) ? 1.0 : 0.0;                                    // :579-582
          p.masks[g] = 1.0 - succ;
          p.rewards[g] = succ - neg;
          const int32_t ha = pick_goal<kInject>(p, 2, i, fin, batch_id, r, g);          // :585-591
          int32_t ha_next, la_next, unused;
          subgoal_step(i, fin, ha, p.k_act, ha_next, unused);                           // :595-600
          subgoal_step(i, fin, ha, p.k_lo, la_next, unused);                            // :613-618
          const int64_t la = (int64_t)i + p.k_act;                                      // :610
          s_row[HGC_HA_GOAL][t] = ha;
          s_row[HGC_HA_NEXT][t] = ha_next;
          s_row[HGC_LA_GOAL][t] = (int32_t)(la < (int64_t)fin ? la : (int64_t)fin);
          s_row[HGC_LA_NEXT][t] = la_next;
        }
      }

      if (p.vec_rows != nullptr) {
        for (int v = 0; v < p.n_slots; ++v) {
          const int32_t x = s_row[v][t];
          p.vec_rows[(int64_t)v * p.total_rows + g] = x;
          if (p.vec_init != nullptr) p.vec_init[(int64_t)v * p.total_rows + g] = trajectory_first_row(p, x);
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
            const int cy = kInject ? (int)p.in_crop[2 * g] : (int)__umulhi(w0.z, span);
            const int cx = kInject ? (int)p.in_crop[2 * g + 1] : (int)__umulhi(w0.w, span);
            dy = cy - p.crop_pad;
            dx = cx - p.crop_pad;
          }
        }
        p.crop_out[2 * g] = (int8_t)dy;
        p.crop_out[2 * g + 1] = (int8_t)dx;
      }
    }
    __syncthreads();

    // ------------------------------- phase 2: row gathers -------------------------------
    // An "item" is one warp-wide access: 32 >> lpr_log2 rows x (1 << lpr_log2) elements.  Items of all jobs are
    // one flat list; every warp issues kGatherUnroll independent loads before the matching stores.
    for (int it0 = warp; it0 < p.total_items; it0 += n_warps * kGatherUnroll) {
      uint4 val[kGatherUnroll];
      uint8_t* dptr[kGatherUnroll];
      int vlog[kGatherUnroll];
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u) {
        vlog[u] = -1;
        const int it = it0 + u * n_warps;
        if (it < p.total_items) {
          int j = 0;
          while (it >= p.item_start[j + 1]) ++j;  // warp-uniform
          const RowJob& job = p.jobs[j];
          const int local = it - p.item_start[j];
          const int pass = job.n_coliter == 1 ? local : local / job.n_coliter;
          const int ci = local - pass * job.n_coliter;
          const int row = (pass << (5 - job.lpr_log2)) + (lane >> job.lpr_log2);
          const int col = (ci << job.lpr_log2) + (lane & ((1 << job.lpr_log2) - 1));
          if (row < n && col < job.epr) {
            const int32_t src_row = s_row[job.slot][row];
            const uint8_t* sp = job.src + (size_t)src_row * job.src_stride + ((size_t)col << job.vec_log2);
            dptr[u] = job.dst + (size_t)(g0 + row) * job.row_bytes + ((size_t)col << job.vec_log2);
            vlog[u] = job.vec_log2;
            switch (job.vec_log2) {
              case 4: val[u] = __ldg(reinterpret_cast<const uint4*>(sp)); break;
              case 3: { const uint2 q = __ldg(reinterpret_cast<const uint2*>(sp)); val[u].x = q.x; val[u].y = q.y; } break;
              case 2: val[u].x = __ldg(reinterpret_cast<const uint32_t*>(sp)); break;
              case 1: val[u].x = __ldg(reinterpret_cast<const uint16_t*>(sp)); break;
              default: val[u].x = __ldg(sp); break;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u) {
        switch (vlog[u]) {
          case 4: *reinterpret_cast<uint4*>(dptr[u]) = val[u]; break;
          case 3: *reinterpret_cast<uint2*>(dptr[u]) = make_uint2(val[u].x, val[u].y); break;
          case 2: *reinterpret_cast<uint32_t*>(dptr[u]) = val[u].x; break;
          case 1: *reinterpret_cast<uint16_t*>(dptr[u]) = (uint16_t)val[u].x; break;
          case 0: *dptr[u] = (uint8_t)val[u].x; break;
          default: break;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace ogb
