// Internal interface between msm.cu (scalar side: digits, counting sort - independent of the group) and the four
// per-group translation units msm_g_*.cu (bucket accumulation / reduction kernels, instantiated from msm_group.cuh).
// Split this way so the slow-to-compile group code builds in parallel.
#pragma once
#include <stdint.h>
#include <atomic>
#include <functional>
#include <vector>
#include "common.cuh"
#include "msm.h"

namespace b200 {

struct MsmPlan;
// Equal bases inside one MSM's point range, found once per key (b200_params_precompute). sum s_i P = (sum s_i) P for the
// members of a group (G1 has prime order r, so the scalar sum may be reduced mod r), hence before the digits are
// extracted the members' scalars are added into the representative's and zeroed: the A query of every key made by
// the reference's generator has m/2 copies of one point (pitfall 2) and costs half as much this way.
// A group is cut into segments of <= kDedupSegment members so that one huge group is summed by many blocks.
struct MsmDedup {
  size_t merged = 0;      // bases folded into a representative (0: nothing to do)
  uint32_t nsegments = 0, ngroups = 0;
  DevBuf members;         // uint32[]: indices into the range, grouped, representative first
  DevBuf segments;        // uint32[3 * nsegments]: first position in `members`, length, group
  DevBuf groups;          // uint32[3 * ngroups]: representative, first segment, number of segments
  DevBuf segment_sums;    // Fr[nsegments] scratch
  void reset() {
    merged = 0;
    nsegments = ngroups = 0;
    members.release();
    segments.release();
    groups.release();
    segment_sums.release();
  }
};
constexpr uint32_t kDedupSegment = 1024;
struct MsmWorkspace {
  DevBuf digits, counts, offsets, cursor, entries, order, counts_sorted, iota, cub_tmp, buckets, red_a, red_b, plan;
  DevBuf ntasks, task_off, task_bucket, task_len, task_len_sorted, partials;
  DevBuf scalar_out, fold_cnt, fold_off, fold_bucket, fold_partials;  // fold_* hold two ping-pong halves
  DevBuf aff_cnt, aff_off, aff_totals, aff_pts[2], aff_scratch;        // batch-affine accumulation (msm_affine_*)
  DevBuf aff_pairs, aff_oflag[2], base_flags;                          // operand pairs of all rounds, O flags per round
  DevBuf merged_scalars;                                               // scalars after equal-base merging (MsmDedup)
  // Two streams per MSM: `stream` (HIGH priority) carries the short, latency-bound kernels - digits, counting sort,
  // task lists, fold/combine, bucket reduction, tree sums; `acc_stream` (LOW priority) carries the long accumulation
  // kernel. The block scheduler serves pending blocks of high-priority streams first, so the latency-bound phases of
  // one MSM (and compute_H) run in the slots that free up while another MSM's accumulation saturates the multiplier.
  cudaStream_t stream = nullptr, acc_stream = nullptr;
  // The scalar-side preparation (digits, counting sort, task lists, batch-affine bookkeeping) runs on a stream of the
  // HIGHEST priority: its kernels are tiny, the issuing host thread has to wait for two of their results (task count,
  // largest bucket), and on the MSM's own low-priority stream they queue behind the accumulation grids of the MSMs
  // issued before - at 8 GPUs (150 K points per rank) the host spent 25 ms of an 88 ms proof blocked there.
  cudaStream_t prep_stream = nullptr;
  cudaEvent_t aff_ready = nullptr;
  cudaEvent_t acc_done = nullptr;
  cudaEvent_t tm_ev[4] = {nullptr, nullptr, nullptr, nullptr};  // diagnostics: digits start/end, sort start/end
  // ring of pinned staging buffers + events for the asynchronous copy of the window sums to the host
  struct Staging {
    void *pinned = nullptr;
    size_t bytes = 0;
    cudaEvent_t done = nullptr;
    cudaEvent_t ta = nullptr, t0 = nullptr, t1 = nullptr;  // accumulate starts at ta, reduce spans t0..t1
    int slot = 0;                                           // workspace slot that issued it (timeline diagnostics)
    cudaEvent_t prep_ev[4] = {nullptr, nullptr, nullptr, nullptr};  // the issuing workspace's tm_ev (null: shared prep)
  } ring[4];
  int ring_pos = 0;
  Staging *next_staging(size_t bytes);
  // sharing of the scalar-side preparation (digits, counting sort, tasks) between MSMs over the SAME scalars and the
  // same window plan (A, B1 and B2 of a proof all use w): the producer records prep_done, consumers wait on it
  cudaEvent_t prep_done = nullptr;
  MsmPlan *prepared = nullptr;  // plan as completed by msm_prepare (ntasks, task_len, ...)
};
// One workspace + stream per MSM of a proof: b200_prove issues its five MSMs on five streams so that the latency-bound
// phases of one (counting sort, bucket reduction) overlap the throughput-bound accumulation of the others.
// msm_select_slot() picks the one used by subsequent calls on this host thread (default 0).
constexpr int kMsmSlots = 5;
MsmWorkspace &msm_workspace();
MsmWorkspace &msm_workspace_slot(int slot);
void msm_select_slot(int slot);
int msm_current_slot();
// MSMs prepared by the calling host thread wait for work already enqueued on this stream (the producer of their
// scalars: copies, compute_H). Default: the legacy default stream.
void msm_set_input_stream(cudaStream_t st);

// Reuse of another MSM's scalar-side preparation (digits, counting sort, task lists): `slot` = the workspace that made
// it for the SAME window plan (-1: prepare on its own). When the consumer's points are numbered differently - the L
// query reads w[i + 2], i.e. its point i belongs to the producer's scalar i + 2 - the producer's entries are re-indexed
// into the consumer's own entry buffer: entry (window j, scalar s) -> j * n + (s - shift), dropped (kSkipEntry) when
// s < shift or s - shift >= n. n_src = the producer's n (0: same numbering, entries are used in place).
struct MsmShare {
  int slot = -1;
  uint32_t n_src = 0, shift = 0;
};
constexpr uint32_t kSkipEntry = 0xffffffffu;
int msm_share_entries(const uint32_t *src_entries, size_t total, uint32_t n_src, uint32_t n_dst, uint32_t shift,
                      uint32_t *dst_entries, cudaStream_t st);

struct MsmPlan {
  int c = 0, W = 0;
  bool merged = false;          // all windows share one bucket set (entries index a table of pre-shifted bases)
  uint32_t nb = 0;              // buckets per window = 2^(c-1)
  size_t nbuckets = 0;          // W * nb, or nb when merged
  std::vector<uint32_t> windows;  // start_bit | width << 16
  uint32_t task_len = 0;        // T: entries per accumulation task (set by msm_prepare)
  size_t ntasks = 0;            // number of tasks of this call (set by msm_prepare)
  uint32_t max_tasks_per_bucket = 0;  // largest number of task sums any bucket has (set by msm_prepare)
  uint32_t max_count = 0;             // largest bucket (set by msm_prepare)
  // bucket reduction by bit planes (msm_reduce_rows_kernel, one bucket set): the device returns S and O_0 .. O_{planes-1},
  // the host finishes S + red_K * sum_b 2^b O_b. -1: the device returns one finished sum per bucket set.
  int red_planes = -1;
  uint32_t red_K = 0;
};

// Diagnostics: when a base event has been set (msm_timeline_begin), msm_collect stores for the issuing slot the times
// (ms since the base) at which the accumulation started, the reduction started and the reduction ended.
void msm_timeline_begin();
void msm_timeline_note(int slot, cudaEvent_t ta, cudaEvent_t t0, cudaEvent_t t1);
void msm_timeline_get(double *out15);
// diagnostics shared by all host threads (two proofs may be in flight): relaxed atomics, no ordering implied
extern std::atomic<double> g_msm_phase_ms[5];         // last call: digits, sort, accumulate, reduce, host tail
extern std::atomic<double> g_msm_phase_total[2][5];   // accumulated, [0] G1 calls, [1] G2 calls
static inline void msm_stat_add(std::atomic<double> &a, double v) {
  a.store(a.load(std::memory_order_relaxed) + v, std::memory_order_relaxed);
}

// Phase 1+2 (group independent): window plan, signed digits, histogram, counting sort of point indices by
// (window, bucket), bucket visiting order by descending size. fr_tag: 0 = modulus A, 1 = modulus B.
// plan.W == 0 on entry: choose a per-window plan for n. Otherwise the caller's plan (e.g. the one a table was built for).
int msm_prepare(int fr_tag, const void *d_scalars, size_t n, MsmPlan &plan, const MsmDedup *dedup = nullptr);
int msm_make_plan(size_t n, bool merged, MsmPlan &plan);
int msm_affine_levels(const uint32_t *counts, const uint32_t *offsets, uint32_t nbuckets, uint32_t max_count,
                      std::vector<size_t> &totals);
int msm_base_flags(const void *d_points, size_t n, size_t point_bytes, DevBuf &flags, cudaStream_t st);
int msm_affine_pairs(MsmWorkspace &ws, const uint32_t *cnt, const uint32_t *off, uint32_t nbuckets,
                     const std::vector<size_t> &totals, const uint32_t *entries, const uint8_t *base_is_O, size_t n_bases,
                     std::vector<size_t> &pair_off);
bool msm_use_batch_affine();
int msm_accum_mode();
bool msm_affine_wins(int degree, size_t entries);
bool msm_affine_split_tail();
bool msm_use_coop(size_t total_buckets, int degree);
void msm_note_concurrent_proofs(int n);
constexpr uint32_t kFoldWidth = 32;  // a bucket with more task sums than this is folded in parallel first
int msm_fold_level(const uint32_t *cnt_in, uint32_t nbuckets, uint32_t width, uint32_t *cnt_out, uint32_t *off_out,
                   size_t &total_out);
// pre-shifted base tables (merged buckets), see msm_group.cuh
int msm_precompute_dispatch(int curve, int group, const void *d_points, size_t n, MsmPlan &plan, DevBuf &table);
int msm_table_dispatch_deferred(int curve, int group, const void *d_scalars, const void *d_table, size_t n,
                                const MsmPlan &plan, void *h_out, MsmTail &tail, MsmShare share = MsmShare(),
                                const MsmDedup *dedup = nullptr);

}  // namespace b200
