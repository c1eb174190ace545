// Pippenger multi-scalar multiplication on one B200 for MNT4753 / MNT6753 G1 and G2.
//
// Replaces libff::multi_exp_with_mixed_addition (depends/libff/libff/algebra/scalar_multiplication/multiexp.tcc:443-496
// -> multi_exp :402-441 -> multi_exp_inner<BDLO12> :165-282) as called by B::multiexp_G1/G2
// (libsnark/prover_reference_functions.cpp:247-265). Same sum, different schedule:
//
//   K3  msm_digits_kernel      scalar -> integer (one Montgomery reduction, fp.tcc:227-238), signed c-bit digits,
//                              per-(window,bucket) histogram                      [1 thread / scalar]
//   K4  scan + msm_scatter     counting sort of (window, |digit|) -> contiguous lists of signed point indices
//   K5  msm_accumulate_kernel  bucket sums by mixed addition, buckets visited largest-first so the threads of a
//                              warp run equally long loops                        [1 thread / bucket]
//   K7  msm_reduce_kernel      running-sum reduction of K consecutive buckets + lo*sum fix-up [1 thread / chunk]
//       msm_sum_kernel         tree sum of the chunk results per window
//       host                   Horner combine of the <= 95 window sums (753 doublings, serial, microseconds each)
//
// Scalars equal to 0 fall out naturally (all digits 0), scalars equal to 1 land in bucket 1 of window 0; points at
// infinity in the bases (y == 0 on the wire) are skipped; P+P and P+(-P) inside a bucket take the same branches as
// the reference's mixed_add (curve.cuh).
#include <cub/cub.cuh>
#include <cstdlib>
#include <functional>
#include "common.cuh"
#include "field.cuh"
#include "msm.h"
#include "msm_internal.h"

namespace b200 {

std::atomic<double> g_msm_phase_ms[5];
std::atomic<double> g_msm_phase_total[2][5];
static std::atomic<int> g_forced_window{0};
// B200_BATCH_AFFINE=1: bucket accumulation by rounds of batched affine additions (experimental, see msm_group.cuh)
static std::atomic<int> g_batch_affine{-1};  // -1: take B200_BATCH_AFFINE from the environment on first use
// 0 = XYZZ mixed additions, 1 = batched affine additions, 2 = automatic (default): batched affine where it is the faster
// one on this hardware - see msm_affine_wins()
int msm_accum_mode() {
  if (g_batch_affine < 0) {
    const char *e = getenv("B200_BATCH_AFFINE");
    g_batch_affine = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
  }
  return g_batch_affine;
}
bool msm_use_batch_affine() { return msm_accum_mode() == 1; }
void msm_set_batch_affine(int on) { g_batch_affine = on < 0 || on > 2 ? 2 : on; }
// Measured on B200 (profiles/r02_summary.md): the batch-affine rounds beat the XYZZ kernel for G2 over Fq2 once a bucket
// set holds millions of entries (2^20 points: 145 vs 157 ms); for G1 the two are within 1 %, and for small MSMs the
// ~7 dependent rounds (one inversion latency each) lose to the single XYZZ launch (2^15 points: 9.3 vs 2.8 ms).
// A 19/64 slice of the MNT4753 B2 query (12.4 M entries, what a GPU of the 8-GPU plan gets): 56.3 ms against 53.7 ms for
// XYZZ - the threshold sits between that and the 24.9 M entries of a 40/64 slice.
bool msm_affine_wins(int degree, size_t entries) { return degree == 2 && entries >= ((size_t)20 << 20); }
// Lane-cooperative bucket reduction (coop.cuh): B200_COOP=0 never, 1 always, default: bucket sets of at most 2^18 buckets,
// where the reduction is a latency problem (sharded proofs, the MNT6753 proof); at 2^20 buckets it is a throughput
// problem and the thread-per-chunk kernel does the same work with all 32 lanes.
// A cooperative addition keeps 5 of 8 ... 30 of 32 lanes busy in its multiplication levels and none in the others: it
// trades multiplier throughput (1.5-2x the pipe time of the thread-per-chunk kernel) for latency (5-15x shorter chains).
// Measured (profiles/r02_summary.md): the MNT6753 2^15 proof alone 57.7 -> 47.3 ms (its G2 reduction 16.9 -> 6.3 ms), but the
// two-proof step, where the small proof's kernels share the multiplier with the large proof's accumulations, 428 -> 438 ms.
// So: only when this is the only proof in flight (b200_prove_batch notes how many it runs).
static std::atomic<int> g_concurrent_proofs{1};
void msm_note_concurrent_proofs(int n) { g_concurrent_proofs = n; }
// G2 over Fq2 at 2^18 buckets (the 19/64 slice of B2): 9.5 ms cooperative against 7.6 ms with a thread per chunk, so its
// limit is one bit lower.
bool msm_use_coop(size_t total_buckets, int degree) {
  static const int mode = getenv("B200_COOP") ? atoi(getenv("B200_COOP")) : 2;
  if (mode == 0) return false;
  if (mode == 1) return true;
  return total_buckets <= ((size_t)1 << (degree == 2 ? 17 : 18)) && g_concurrent_proofs <= 1;
}
// B200_AFF_SPLIT=1: cut large batch-affine rounds into an 80 % and a 20 % region (see AffRegions) to fill the tail of the
// single wave. Measured on B200 and left off: the second region's shorter batches pay more per addition for the shared
// inversion than the tail costs (G1 2^20: 52.9 vs 49.2 ms, G2: 149.5 vs 145.2 ms).
bool msm_affine_split_tail() {
  static const bool on = getenv("B200_AFF_SPLIT") && getenv("B200_AFF_SPLIT")[0] == '1';
  return on;
}
void msm_set_window(int c) { g_forced_window = c; }
void msm_phase_totals(double *out10, int reset) {
  for (int g = 0; g < 2; g++)
    for (int i = 0; i < 5; i++) {
      out10[g * 5 + i] = g_msm_phase_total[g][i];
      if (reset) g_msm_phase_total[g][i] = 0;
    }
}
static cudaEvent_t g_timeline_base = nullptr;
static std::atomic<double> g_timeline[kMsmSlots][3];
void msm_timeline_begin() {
  if (!g_timeline_base) cudaEventCreate(&g_timeline_base);
  cudaEventRecord(g_timeline_base, 0);
  for (int s = 0; s < kMsmSlots; s++)
    for (int k = 0; k < 3; k++) g_timeline[s][k] = -1;
}
void msm_timeline_note(int slot, cudaEvent_t ta, cudaEvent_t t0, cudaEvent_t t1) {
  if (!g_timeline_base || slot < 0 || slot >= kMsmSlots) return;
  cudaEvent_t ev[3] = {ta, t0, t1};
  for (int k = 0; k < 3; k++) {
    float ms = -1;
    if (cudaEventElapsedTime(&ms, g_timeline_base, ev[k]) != cudaSuccess) {
      cudaGetLastError();
      ms = -1;
    }
    g_timeline[slot][k] = ms;
  }
}
void msm_timeline_get(double *out15) {
  for (int s = 0; s < kMsmSlots; s++)
    for (int k = 0; k < 3; k++) out15[s * 3 + k] = g_timeline[s][k];
}
static std::atomic<int> g_last_plan[3];
void msm_last_plan(int *out3) {
  for (int i = 0; i < 3; i++) out3[i] = g_last_plan[i];
}
void msm_last_phase_ms(double *out5) {
  for (int i = 0; i < 5; i++) out5[i] = g_msm_phase_ms[i];
}

// ---------------------------------------------------------------------------------------------- kernels
// Window plan: W windows tile bits [0, 753) of the scalar. plan[j] = start_bit | width << 16. All windows but the top
// one are recoded to signed digits in [-2^(w-1), 2^(w-1)] (carry into the next window); the top window has width c-1
// and stays unsigned, so it needs the same 2^(c-1) buckets as a full signed window and never carries out. Widths
// are c or c-1, chosen so that no window is a narrow left-over: uniformly distributed scalars give uniformly filled
// buckets in every window (a short top window would put ~n/2 points into one bucket).
template <class FrP>
__global__ void __launch_bounds__(128) msm_digits_kernel(const Fp<FrP> *__restrict__ scalars, uint32_t n, int c, int W,
                                                         int merged, const uint32_t *__restrict__ plan,
                                                         int32_t *__restrict__ digits, uint32_t *__restrict__ counts) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<FrP> s = scalars[i];
  Fp<FrP>::from_mont(s, s);
  const uint32_t nb = 1u << (c - 1);
  uint32_t carry = 0;
  for (int j = 0; j < W; j++) {
    const uint32_t pj = plan[j];
    const uint32_t bitpos = pj & 0xffffu, cw = pj >> 16;
    const uint32_t word = bitpos >> 5, off = bitpos & 31;
    uint64_t two = word < (uint32_t)kLimbs ? s.l[word] : 0u;
    if (word + 1 < (uint32_t)kLimbs) two |= (uint64_t)s.l[word + 1] << 32;
    uint32_t v = ((uint32_t)(two >> off) & ((1u << cw) - 1)) + carry;
    int32_t d;
    if (j < W - 1 && v > (1u << (cw - 1))) {
      d = (int32_t)v - (int32_t)(1u << cw);
      carry = 1;
    } else {
      d = (int32_t)v;
      carry = 0;
    }
    digits[(size_t)j * n + i] = d;
    if (d != 0) {
      uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
      atomicAdd(&counts[(merged ? (size_t)0 : (size_t)j * nb) + (mag - 1)], 1u);
    }
  }
}

// Equal-base merging, step 1: one block per segment of a group adds the members' scalars (Montgomery form, so the sum
// is just the field sum) and zeroes them in the working copy; step 2: one block per group adds the segment sums and
// stores the total as the representative's scalar.
template <class FrP>
__global__ void __launch_bounds__(128) msm_dedup_segment_kernel(const Fp<FrP> *__restrict__ scalars,
                                                                const uint32_t *__restrict__ members,
                                                                const uint32_t *__restrict__ segments,
                                                                Fp<FrP> *__restrict__ merged,
                                                                Fp<FrP> *__restrict__ segment_sums) {
  typedef Fp<FrP> F;
  __shared__ F part[128];
  const uint32_t first = segments[3 * blockIdx.x], len = segments[3 * blockIdx.x + 1];
  F acc, zero;
  F::set_zero(acc);
  F::set_zero(zero);
  for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) {
    const uint32_t i = members[first + k];
    F s = scalars[i];
    F::add(acc, acc, s);
    merged[i] = zero;
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t w = blockDim.x / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) F::add(part[threadIdx.x], part[threadIdx.x], part[threadIdx.x + w]);
    __syncthreads();
  }
  if (threadIdx.x == 0) segment_sums[blockIdx.x] = part[0];
}
template <class FrP>
__global__ void __launch_bounds__(128) msm_dedup_group_kernel(const uint32_t *__restrict__ groups,
                                                              const Fp<FrP> *__restrict__ segment_sums,
                                                              Fp<FrP> *__restrict__ merged) {
  typedef Fp<FrP> F;
  __shared__ F part[128];
  const uint32_t rep = groups[3 * blockIdx.x], first = groups[3 * blockIdx.x + 1], cnt = groups[3 * blockIdx.x + 2];
  F acc;
  F::set_zero(acc);
  for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x) {
    F s = segment_sums[first + k];
    F::add(acc, acc, s);
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t w = blockDim.x / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) F::add(part[threadIdx.x], part[threadIdx.x], part[threadIdx.x + w]);
    __syncthreads();
  }
  if (threadIdx.x == 0) merged[rep] = part[0];
}
template <class FrP>
static int msm_dedup_scalars(const void *d_scalars, size_t n, const MsmDedup &dd, MsmWorkspace &ws) {
  typedef Fp<FrP> F;
  cudaStream_t st = ws.prep_stream;
  B200_CHECK(ws.merged_scalars.reserve(n * sizeof(F)));
  B200_CUDA_CHECK(cudaMemcpyAsync(ws.merged_scalars.p, d_scalars, n * sizeof(F), cudaMemcpyDeviceToDevice, st));
  msm_dedup_segment_kernel<FrP><<<dd.nsegments, 128, 0, st>>>((const F *)d_scalars, dd.members.as<uint32_t>(),
                                                            dd.segments.as<uint32_t>(), ws.merged_scalars.as<F>(),
                                                            dd.segment_sums.as<F>());
  B200_CUDA_CHECK(cudaGetLastError());
  msm_dedup_group_kernel<FrP><<<dd.ngroups, 128, 0, st>>>(dd.groups.as<uint32_t>(), dd.segment_sums.as<F>(),
                                                        ws.merged_scalars.as<F>());
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch(2);
  return 0;
}

// merged == 0: bucket space is (window, |digit|), an entry is a point index into the n bases.
// merged == 1: all windows share one bucket set, an entry is j*n + i, an index into the table of pre-shifted bases
//              2^(start_j) * P_i (see msm_precompute_kernel).
__global__ void __launch_bounds__(256) msm_scatter_kernel(const int32_t *__restrict__ digits, uint32_t n, int W, int c,
                                                          int merged, const uint32_t *__restrict__ offsets,
                                                          uint32_t *__restrict__ cursor, uint32_t *__restrict__ entries) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)W * n) return;
  int32_t d = digits[idx];
  if (d == 0) return;
  uint32_t j = (uint32_t)(idx / n), i = (uint32_t)(idx % n);
  const uint32_t nb = 1u << (c - 1);
  uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
  size_t b = (merged ? (size_t)0 : (size_t)j * nb) + (mag - 1);
  uint32_t pos = offsets[b] + atomicAdd(&cursor[b], 1u);
  uint32_t e = merged ? (uint32_t)idx : i;
  entries[pos] = (e << 1) | (d < 0 ? 1u : 0u);
}

// re-index a producer's entry list for a consumer whose points are numbered with an offset (see MsmShare)
__global__ void __launch_bounds__(256) msm_share_entries_kernel(const uint32_t *__restrict__ src, size_t total, uint32_t n_src,
                                                                uint32_t n_dst, uint32_t shift, uint32_t *__restrict__ dst) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  const uint32_t e = src[k];
  uint32_t out = kSkipEntry;
  if (e != kSkipEntry) {
    const uint32_t idx = e >> 1, j = idx / n_src, s = idx - j * n_src;
    if (s >= shift && s - shift < n_dst) out = ((j * n_dst + (s - shift)) << 1) | (e & 1u);
  }
  dst[k] = out;
}
int msm_share_entries(const uint32_t *src_entries, size_t total, uint32_t n_src, uint32_t n_dst, uint32_t shift,
                      uint32_t *dst_entries, cudaStream_t st) {
  if (total == 0) return 0;
  msm_share_entries_kernel<<<grid_for(total, 256), 256, 0, st>>>(src_entries, total, n_src, n_dst, shift, dst_entries);
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}

// ---- folding of task sums for skewed inputs (a bucket with more than kFoldWidth task sums) ----------------------
__global__ void msm_fold_counts_kernel(const uint32_t *__restrict__ cnt_in, uint32_t nbuckets, uint32_t width,
                                       uint32_t *__restrict__ cnt_out) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nbuckets) cnt_out[b] = (cnt_in[b] + width - 1) / width;
}

// ---- tasks: a bucket's entry list is cut into pieces of at most T entries; one thread sums one piece -------------
__global__ void msm_ntasks_kernel(const uint32_t *__restrict__ counts, uint32_t nbuckets, uint32_t T,
                                  uint32_t *__restrict__ ntasks) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nbuckets) ntasks[b] = (counts[b] + T - 1) / T;
}
// one thread per bucket writes the descriptors of its tasks: owning bucket and length
__global__ void msm_task_fill_kernel(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ task_off,
                                     uint32_t nbuckets, uint32_t T, uint32_t *__restrict__ task_bucket,
                                     uint32_t *__restrict__ task_len) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  uint32_t cnt = counts[b], t = task_off[b];
  for (uint32_t done = 0; done < cnt; done += T, t++) {
    task_bucket[t] = b;
    task_len[t] = cnt - done < T ? cnt - done : T;
  }
}

__global__ void iota_kernel(uint32_t *v, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}


// ---------------------------------------------------------------------------------------------- host side
// Window width: minimise (bucket accumulation + bucket reduction) field multiplications. With pre-shifted bases
// (merged) the bucket reduction is paid once instead of once per window, which moves the optimum to wider windows.
static int choose_window(size_t n, bool merged) {
  if (g_forced_window >= 3 && g_forced_window <= 22) return g_forced_window;
  double best = 1e300;
  int best_c = 4;
  // (table mode: at least 8 bits, i.e. at most 95 windows - the table builder holds one Z per window on its stack)
  for (int c = merged ? 8 : 3; c <= (merged ? 22 : 20); c++) {
    double W = (754 + c - 1) / c;
    double nb = (double)(1u << (c - 1));
    double red = nb * 2.0 * 14.0 + nb / 32.0 * 30.0 * 13.0;
    // (round 2: weighting the merged bucket set's reduction 1.7x - its measured latency-bound cost - moves every choice one
    // bit down, c = 20 / 38 windows at 2^20; measured: G1 52.6 + 5.2 ms against 49.3 + 7.6 ms, G2 152.7 + 17.9 against
    // 146 + 26, step 430.1 against 428.3 ms - a wash, so the multiplication-count model stays)
    double cost = merged ? W * (double)n * 11.0 + red : W * ((double)n * 11.0 + red);
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

// window plan (see msm_digits_kernel): top window c-1 bits, `excess` low windows c-1 bits, the rest c bits
int msm_make_plan(size_t n, bool merged, MsmPlan &plan) {
  const int c = choose_window(n, merged);
  const int W = (754 + c - 1) / c;
  plan.c = c;
  plan.W = W;
  plan.merged = merged;
  plan.nb = 1u << (c - 1);
  plan.nbuckets = merged ? (size_t)plan.nb : (size_t)W * plan.nb;
  plan.windows.resize(W);
  int excess = W * c - 1 - 753;
  uint32_t start = 0;
  for (int j = 0; j < W; j++) {
    uint32_t width = (j == W - 1 || j < excess) ? (uint32_t)(c - 1) : (uint32_t)c;
    plan.windows[j] = start | (width << 16);
    start += width;
  }
  if (start != 753 || excess > W - 1) return set_error(-2, "msm: bad window plan c=%d", c);
  if ((size_t)W * n >= (1ull << 31)) return set_error(-2, "msm: W*n = %zu overflows 31-bit entry indices", (size_t)W * n);
  return 0;
}

static thread_local cudaStream_t g_input_stream = 0;
void msm_set_input_stream(cudaStream_t st) { g_input_stream = st; }
static thread_local bool g_high_priority = false;
void msm_thread_high_priority(bool on) { g_high_priority = on; }
static thread_local int g_slot = 0;
int msm_current_slot() { return g_slot; }
void msm_select_slot(int slot) { g_slot = ((slot % kMsmSlots) + kMsmSlots) % kMsmSlots; }
static MsmWorkspace *workspace_slots() {
  static thread_local MsmWorkspace ws[kMsmSlots];
  return ws;
}
MsmWorkspace &msm_workspace_slot(int slot) {
  MsmWorkspace &ws = workspace_slots()[slot];
  if (!ws.stream) {
    // Default: one stream per MSM for all of its kernels (worker threads of b200_prove_batch - the smaller proofs -
    // get high-priority streams so that they are not stuck behind the large proof's grids).
    // B200_SPLIT_STREAMS=1 (experiment, slower: profiles/r01_v3_summary.md): short kernels on a high-priority stream,
    // all accumulations of the thread on ONE low-priority stream.
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    static const bool split = getenv("B200_SPLIT_STREAMS") && getenv("B200_SPLIT_STREAMS")[0] == '1';
    if (split) {
      cudaStreamCreateWithPriority(&ws.stream, cudaStreamNonBlocking, greatest);
      static thread_local cudaStream_t acc = nullptr;
      if (!acc) cudaStreamCreateWithPriority(&acc, cudaStreamNonBlocking, g_high_priority ? (least + greatest) / 2 : least);
      ws.acc_stream = acc;
    } else {
      cudaStreamCreateWithPriority(&ws.stream, cudaStreamNonBlocking, g_high_priority ? greatest : least);
      ws.acc_stream = ws.stream;
    }
    cudaStreamCreateWithPriority(&ws.prep_stream, cudaStreamNonBlocking, greatest);
    cudaEventCreateWithFlags(&ws.aff_ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ws.acc_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ws.prep_done, cudaEventDisableTiming);
    ws.prepared = new MsmPlan();
  }
  return ws;
}
MsmWorkspace &msm_workspace() { return msm_workspace_slot(g_slot); }
MsmWorkspace::Staging *MsmWorkspace::next_staging(size_t bytes) {
  Staging &s = ring[ring_pos];
  ring_pos = (ring_pos + 1) & 3;
  if (!s.done) {
    cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    cudaEventCreate(&s.ta);
    cudaEventCreate(&s.t0);
    cudaEventCreate(&s.t1);
  }
  if (s.bytes < bytes) {
    if (s.pinned) cudaFreeHost(s.pinned);
    if (cudaMallocHost(&s.pinned, bytes) != cudaSuccess) {
      s.pinned = nullptr;
      s.bytes = 0;
      return nullptr;
    }
    s.bytes = bytes;
  }
  return &s;
}
void msm_release_workspace() {
  for (int s = 0; s < kMsmSlots; s++) {
  MsmWorkspace &ws = workspace_slots()[s];
  DevBuf *all[] = {&ws.digits, &ws.counts, &ws.offsets, &ws.cursor, &ws.entries, &ws.order,
                   &ws.counts_sorted, &ws.iota, &ws.cub_tmp, &ws.buckets, &ws.red_a, &ws.red_b, &ws.plan,
                   &ws.ntasks, &ws.task_off, &ws.task_bucket, &ws.task_len, &ws.task_len_sorted, &ws.partials,
                   &ws.scalar_out, &ws.fold_cnt, &ws.fold_off, &ws.fold_bucket, &ws.fold_partials,
                   &ws.aff_cnt, &ws.aff_off, &ws.aff_totals, &ws.aff_pts[0], &ws.aff_pts[1], &ws.aff_scratch,
                   &ws.aff_pairs, &ws.aff_oflag[0], &ws.aff_oflag[1], &ws.base_flags,
                   &ws.merged_scalars};
  for (DevBuf *b : all) b->release();
  }
}

int msm_prepare(int fr_tag, const void *d_scalars, size_t n, MsmPlan &plan, const MsmDedup *dedup) {
  if (n >= (1ull << 30)) return set_error(-2, "msm: n=%zu too large", n);
  if (plan.W == 0) B200_CHECK(msm_make_plan(n, false, plan));  // caller did not fix a plan: per-window buckets
  const int c = plan.c, W = plan.W;
  const int merged = plan.merged ? 1 : 0;
  const size_t nbuckets = plan.nbuckets;
  MsmWorkspace &ws = msm_workspace();
  cudaStream_t st = ws.prep_stream;
  {
    // this workspace's previous MSM may still be reading the arrays that are rebuilt here
    static thread_local cudaEvent_t prev = nullptr;
    if (!prev) B200_CUDA_CHECK(cudaEventCreateWithFlags(&prev, cudaEventDisableTiming));
    B200_CUDA_CHECK(cudaEventRecord(prev, ws.stream));
    B200_CUDA_CHECK(cudaStreamWaitEvent(st, prev, 0));
    if (ws.acc_stream != ws.stream) {
      B200_CUDA_CHECK(cudaEventRecord(prev, ws.acc_stream));
      B200_CUDA_CHECK(cudaStreamWaitEvent(st, prev, 0));
    }
  }
  {
    // inputs are produced on the default stream (copies, compute_H, generators): order this MSM after it
    static thread_local cudaEvent_t fence = nullptr;
    if (!fence) B200_CUDA_CHECK(cudaEventCreateWithFlags(&fence, cudaEventDisableTiming));
    B200_CUDA_CHECK(cudaEventRecord(fence, g_input_stream));
    B200_CUDA_CHECK(cudaStreamWaitEvent(st, fence, 0));
  }
  B200_CHECK(ws.digits.reserve((size_t)W * n * sizeof(int32_t)));
  B200_CHECK(ws.entries.reserve((size_t)W * n * sizeof(uint32_t)));
  B200_CHECK(ws.counts.reserve(nbuckets * sizeof(uint32_t)));
  B200_CHECK(ws.offsets.reserve(nbuckets * sizeof(uint32_t)));
  B200_CHECK(ws.cursor.reserve(nbuckets * sizeof(uint32_t)));
  B200_CHECK(ws.ntasks.reserve(nbuckets * sizeof(uint32_t)));
  B200_CHECK(ws.task_off.reserve(nbuckets * sizeof(uint32_t)));
  B200_CHECK(ws.plan.reserve(256 * sizeof(uint32_t)));
  B200_CHECK(ws.scalar_out.reserve(64));
  B200_CUDA_CHECK(cudaMemcpyAsync(ws.plan.p, plan.windows.data(), W * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  // diagnostic timers: persistent events, recorded asynchronously; nothing waits on them (msm_collect reads them after
  // the whole MSM has drained)
  for (int i = 0; i < 4; i++)
    if (!ws.tm_ev[i]) B200_CUDA_CHECK(cudaEventCreate(&ws.tm_ev[i]));

  // ---- digits + histogram (after folding the scalars of equal bases into one representative each)
  B200_CUDA_CHECK(cudaEventRecord(ws.tm_ev[0], st));
  if (dedup && dedup->merged) {
    if (fr_tag == 0) B200_CHECK(msm_dedup_scalars<PrimeA>(d_scalars, n, *dedup, ws));
    else B200_CHECK(msm_dedup_scalars<PrimeB>(d_scalars, n, *dedup, ws));
    d_scalars = ws.merged_scalars.p;
  }
  B200_CUDA_CHECK(cudaMemsetAsync(ws.counts.p, 0, nbuckets * sizeof(uint32_t), st));
  B200_CUDA_CHECK(cudaMemsetAsync(ws.cursor.p, 0, nbuckets * sizeof(uint32_t), st));
  if (fr_tag == 0)
    msm_digits_kernel<PrimeA><<<grid_for(n, 128), 128, 0, st>>>((const Fp<PrimeA> *)d_scalars, (uint32_t)n, c, W, merged,
                                                         ws.plan.as<uint32_t>(), ws.digits.as<int32_t>(),
                                                         ws.counts.as<uint32_t>());
  else
    msm_digits_kernel<PrimeB><<<grid_for(n, 128), 128, 0, st>>>((const Fp<PrimeB> *)d_scalars, (uint32_t)n, c, W, merged,
                                                         ws.plan.as<uint32_t>(), ws.digits.as<int32_t>(),
                                                         ws.counts.as<uint32_t>());
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  B200_CUDA_CHECK(cudaEventRecord(ws.tm_ev[1], st));

  // ---- counting sort of the entries by bucket: scan + scatter
  B200_CUDA_CHECK(cudaEventRecord(ws.tm_ev[2], st));
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ws.counts.as<uint32_t>(), ws.offsets.as<uint32_t>(), (int)nbuckets);
  B200_CHECK(ws.cub_tmp.reserve(tmp_bytes));
  size_t tb = ws.cub_tmp.bytes;
  B200_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ws.cub_tmp.p, tb, ws.counts.as<uint32_t>(), ws.offsets.as<uint32_t>(),
                                                (int)nbuckets, st));
  {
    size_t total = (size_t)W * n;
    msm_scatter_kernel<<<grid_for(total, 256), 256, 0, st>>>(ws.digits.as<int32_t>(), (uint32_t)n, W, c, merged,
                                                      ws.offsets.as<uint32_t>(), ws.cursor.as<uint32_t>(),
                                                      ws.entries.as<uint32_t>());
    B200_CUDA_CHECK(cudaGetLastError());
    note_launch();
  }
  // ---- tasks: pieces of <= T entries, visited longest-first (equal loop lengths inside a warp). T keeps a few
  // hundred thousand threads in flight even when there are few buckets (small n, or one merged bucket set), and it
  // bounds the serial chain of any single thread when the scalar distribution is skewed.
  {
    size_t entries_total = (size_t)W * n;
    uint32_t T = 64;
    while (T > 8 && entries_total / T < 260000) T >>= 1;
    plan.task_len = T;
    msm_ntasks_kernel<<<grid_for(nbuckets, 256), 256, 0, st>>>(ws.counts.as<uint32_t>(), (uint32_t)nbuckets, T,
                                                       ws.ntasks.as<uint32_t>());
    note_launch();
    tb = ws.cub_tmp.bytes;
    B200_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ws.cub_tmp.p, tb, ws.ntasks.as<uint32_t>(),
                                                  ws.task_off.as<uint32_t>(), (int)nbuckets, st));
    uint32_t last[3];
    {
      size_t need = 0;
      cub::DeviceReduce::Max(nullptr, need, ws.ntasks.as<uint32_t>(), ws.scalar_out.as<uint32_t>(), (int)nbuckets, st);
      B200_CHECK(ws.cub_tmp.reserve(need));
      tb = ws.cub_tmp.bytes;
      B200_CUDA_CHECK(cub::DeviceReduce::Max(ws.cub_tmp.p, tb, ws.ntasks.as<uint32_t>(), ws.scalar_out.as<uint32_t>(),
                                             (int)nbuckets, st));
    }
    B200_CUDA_CHECK(cudaMemcpyAsync(&last[0], ws.task_off.as<uint32_t>() + (nbuckets - 1), 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(&last[1], ws.ntasks.as<uint32_t>() + (nbuckets - 1), 4, cudaMemcpyDeviceToHost, st));
    {
      size_t need = 0;
      cub::DeviceReduce::Max(nullptr, need, ws.counts.as<uint32_t>(), ws.scalar_out.as<uint32_t>() + 1, (int)nbuckets, st);
      B200_CHECK(ws.cub_tmp.reserve(need));
      tb = ws.cub_tmp.bytes;
      B200_CUDA_CHECK(cub::DeviceReduce::Max(ws.cub_tmp.p, tb, ws.counts.as<uint32_t>(), ws.scalar_out.as<uint32_t>() + 1,
                                             (int)nbuckets, st));
    }
    uint32_t maxcount = 0;
    B200_CUDA_CHECK(cudaMemcpyAsync(&last[2], ws.scalar_out.as<uint32_t>(), 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(&maxcount, ws.scalar_out.as<uint32_t>() + 1, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    plan.max_count = maxcount;
    const size_t ntasks = (size_t)last[0] + last[1];
    plan.ntasks = ntasks;
    plan.max_tasks_per_bucket = last[2];
    const size_t cap = ntasks ? ntasks : 1;
    B200_CHECK(ws.task_bucket.reserve(cap * sizeof(uint32_t)));
    B200_CHECK(ws.task_len.reserve(cap * sizeof(uint32_t)));
    B200_CHECK(ws.task_len_sorted.reserve(cap * sizeof(uint32_t)));
    B200_CHECK(ws.order.reserve(cap * sizeof(uint32_t)));
    B200_CHECK(ws.iota.reserve(cap * sizeof(uint32_t)));
    if (ntasks) {
      msm_task_fill_kernel<<<grid_for(nbuckets, 256), 256, 0, st>>>(ws.counts.as<uint32_t>(), ws.task_off.as<uint32_t>(),
                                                            (uint32_t)nbuckets, T, ws.task_bucket.as<uint32_t>(),
                                                            ws.task_len.as<uint32_t>());
      note_launch();
      iota_kernel<<<grid_for(ntasks, 256), 256, 0, st>>>(ws.iota.as<uint32_t>(), (uint32_t)ntasks);
      note_launch();
      int end_bit = 1;
      while ((1u << end_bit) <= T) end_bit++;
      size_t need = 0;
      cub::DeviceRadixSort::SortPairsDescending(nullptr, need, ws.task_len.as<uint32_t>(), ws.task_len_sorted.as<uint32_t>(),
                                                ws.iota.as<uint32_t>(), ws.order.as<uint32_t>(), (int)ntasks, 0, end_bit);
      B200_CHECK(ws.cub_tmp.reserve(need));
      tb = ws.cub_tmp.bytes;
      B200_CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(ws.cub_tmp.p, tb, ws.task_len.as<uint32_t>(),
                                                                ws.task_len_sorted.as<uint32_t>(), ws.iota.as<uint32_t>(),
                                                                ws.order.as<uint32_t>(), (int)ntasks, 0, end_bit, st));
    }
  }
  B200_CUDA_CHECK(cudaEventRecord(ws.tm_ev[3], st));
  *ws.prepared = plan;
  g_last_plan[0] = plan.c;
  g_last_plan[1] = plan.W;
  g_last_plan[2] = (int)plan.task_len;
  B200_CUDA_CHECK(cudaEventRecord(ws.prep_done, st));
  return 0;
}

// bookkeeping of one fold level (group-independent): cnt_out = ceil(cnt_in / width), off_out = exclusive scan,
// returns the total number of groups (synchronises the stream; only reached for skewed scalar distributions)
int msm_fold_level(const uint32_t *cnt_in, uint32_t nbuckets, uint32_t width, uint32_t *cnt_out, uint32_t *off_out,
                   size_t &total_out) {
  MsmWorkspace &ws = msm_workspace();
  cudaStream_t st = ws.stream;
  msm_fold_counts_kernel<<<grid_for(nbuckets, 256), 256, 0, st>>>(cnt_in, nbuckets, width, cnt_out);
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, cnt_out, off_out, (int)nbuckets, st);
  B200_CHECK(ws.cub_tmp.reserve(need));
  size_t tb = ws.cub_tmp.bytes;
  B200_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ws.cub_tmp.p, tb, cnt_out, off_out, (int)nbuckets, st));
  uint32_t last[2];
  B200_CUDA_CHECK(cudaMemcpyAsync(&last[0], off_out + (nbuckets - 1), 4, cudaMemcpyDeviceToHost, st));
  B200_CUDA_CHECK(cudaMemcpyAsync(&last[1], cnt_out + (nbuckets - 1), 4, cudaMemcpyDeviceToHost, st));
  B200_CUDA_CHECK(cudaStreamSynchronize(st));
  total_out = (size_t)last[0] + last[1];
  return 0;
}

// ---- batch-affine bookkeeping kernels ---------------------------------------------------------------------------
// base_is_O[i] = 1 when base i is the point at infinity (y == 0 on the wire, serialization.hpp:87-89)
__global__ void msm_base_flags_kernel(const unsigned char *__restrict__ points, uint32_t n, uint32_t point_bytes,
                                      uint8_t *__restrict__ flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 *y = reinterpret_cast<const uint4 *>(points + (size_t)i * point_bytes + point_bytes / 2);
  uint32_t acc = 0;
  for (uint32_t k = 0; k < point_bytes / 32; k++) {
    const uint4 v = y[k];
    acc |= v.x | v.y | v.z | v.w;
  }
  flags[i] = acc == 0 ? 1 : 0;
}
// Operand pairs of one round: output j of the round (bucket b = the last one with off_out[b] <= j, position i = j -
// off_out[b]) adds inputs 2i and 2i+1 of list b (the second is missing for an odd leftover). Round 1 (entries !=
// nullptr): an operand is the counting sort's entry (table index << 1 | negate), bit 31 set when that base is O;
// later rounds: the index into the previous round's output << 1. One thread per output: a binary search instead of a
// walk along the bucket, so that a bucket holding millions of entries (skewed scalars) costs nothing special.
__global__ void __launch_bounds__(256) msm_affine_pairs_kernel(const uint32_t *__restrict__ cnt_in, const uint32_t *__restrict__ off_in,
                                                               const uint32_t *__restrict__ off_out, uint32_t nbuckets,
                                                               uint32_t total_out, const uint32_t *__restrict__ entries,
                                                               const uint8_t *__restrict__ base_is_O, uint32_t n_bases,
                                                               uint2 *__restrict__ pairs) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total_out) return;
  uint32_t lo = 0, hi = nbuckets;  // empty buckets share the offset of the next non-empty one: take the LAST b with off <= j
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (off_out[mid] <= j) lo = mid;
    else hi = mid;
  }
  const uint32_t b = lo, i = j - off_out[b];
  const uint32_t c = cnt_in[b], in = off_in[b] + 2 * i;
  const bool second = 2 * i + 1 < c;
  uint2 pr;
  if (entries) {
    // a dropped entry of a shared list (kSkipEntry) becomes an operand that is O
    auto operand = [&](uint32_t e) -> uint32_t {
      if (e == kSkipEntry) return 0x80000000u;
      return base_is_O[(e >> 1) % n_bases] ? (e | 0x80000000u) : e;
    };
    pr.x = operand(entries[in]);
    pr.y = second ? operand(entries[in + 1]) : 0xffffffffu;
  } else {
    pr.x = in << 1;
    pr.y = second ? (in + 1) << 1 : 0xffffffffu;
  }
  pairs[j] = pr;
}

// Batch-affine accumulation bookkeeping: level 0 = the bucket counts/offsets of the counting sort; level r+1 halves every
// list (ceil). All levels are computed up front so that the round kernels can be enqueued without host round trips.
int msm_affine_levels(const uint32_t *counts, const uint32_t *offsets, uint32_t nbuckets, uint32_t max_count,
                      std::vector<size_t> &totals) {
  MsmWorkspace &ws = msm_workspace();
  cudaStream_t st = ws.prep_stream;
  int rounds = 0;
  for (uint32_t c = max_count; c > 1; c = (c + 1) / 2) rounds++;
  totals.assign(rounds + 1, 0);
  B200_CHECK(ws.aff_cnt.reserve((size_t)(rounds + 1) * nbuckets * sizeof(uint32_t)));
  B200_CHECK(ws.aff_off.reserve((size_t)(rounds + 1) * nbuckets * sizeof(uint32_t)));
  B200_CHECK(ws.aff_totals.reserve((size_t)(rounds + 1) * 2 * sizeof(uint32_t)));
  uint32_t *cnt = ws.aff_cnt.as<uint32_t>(), *off = ws.aff_off.as<uint32_t>();
  B200_CUDA_CHECK(cudaMemcpyAsync(cnt, counts, (size_t)nbuckets * 4, cudaMemcpyDeviceToDevice, st));
  B200_CUDA_CHECK(cudaMemcpyAsync(off, offsets, (size_t)nbuckets * 4, cudaMemcpyDeviceToDevice, st));
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, cnt, off, (int)nbuckets, st);
  B200_CHECK(ws.cub_tmp.reserve(need));
  for (int r = 1; r <= rounds; r++) {
    uint32_t *ci = cnt + (size_t)(r - 1) * nbuckets, *co = cnt + (size_t)r * nbuckets, *oo = off + (size_t)r * nbuckets;
    msm_fold_counts_kernel<<<grid_for(nbuckets, 256), 256, 0, st>>>(ci, nbuckets, 2, co);
    B200_CUDA_CHECK(cudaGetLastError());
    note_launch();
    size_t tb = ws.cub_tmp.bytes;
    B200_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ws.cub_tmp.p, tb, co, oo, (int)nbuckets, st));
  }
  std::vector<uint32_t> last(2 * (rounds + 1));
  for (int r = 0; r <= rounds; r++) {
    B200_CUDA_CHECK(cudaMemcpyAsync(&last[2 * r], off + (size_t)r * nbuckets + (nbuckets - 1), 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(&last[2 * r + 1], cnt + (size_t)r * nbuckets + (nbuckets - 1), 4, cudaMemcpyDeviceToHost, st));
  }
  B200_CUDA_CHECK(cudaStreamSynchronize(st));
  for (int r = 0; r <= rounds; r++) totals[r] = (size_t)last[2 * r] + last[2 * r + 1];
  return 0;
}

int msm_base_flags(const void *d_points, size_t n, size_t point_bytes, DevBuf &flags, cudaStream_t st) {
  B200_CHECK(flags.reserve(n ? n : 1));
  if (n == 0) return 0;
  msm_base_flags_kernel<<<grid_for(n, 256), 256, 0, st>>>((const unsigned char *)d_points, (uint32_t)n, (uint32_t)point_bytes,
                                                          flags.as<uint8_t>());
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}

// operand pairs of every round, one after the other in ws.aff_pairs (round r starts at pair_off[r]); `levels_ws` owns the
// level arrays (aff_cnt / aff_off) and `entries` is the counting sort's list the first round reads
int msm_affine_pairs(MsmWorkspace &ws, const uint32_t *cnt, const uint32_t *off, uint32_t nbuckets,
                     const std::vector<size_t> &totals, const uint32_t *entries, const uint8_t *base_is_O, size_t n_bases,
                     std::vector<size_t> &pair_off) {
  cudaStream_t st = ws.prep_stream;
  const int rounds = (int)totals.size() - 1;
  pair_off.assign(rounds + 2, 0);
  for (int r = 1; r <= rounds; r++) pair_off[r + 1] = pair_off[r] + totals[r];
  B200_CHECK(ws.aff_pairs.reserve((pair_off[rounds + 1] ? pair_off[rounds + 1] : 1) * sizeof(uint2)));
  for (int r = 1; r <= rounds; r++) {
    if (totals[r] == 0) continue;
    msm_affine_pairs_kernel<<<grid_for(totals[r], 256), 256, 0, st>>>(
        cnt + (size_t)(r - 1) * nbuckets, off + (size_t)(r - 1) * nbuckets, off + (size_t)r * nbuckets, nbuckets,
        (uint32_t)totals[r], r == 1 ? entries : nullptr, base_is_O, (uint32_t)n_bases, ws.aff_pairs.as<uint2>() + pair_off[r]);
    B200_CUDA_CHECK(cudaGetLastError());
    note_launch();
  }
  return 0;
}

int msm_run_mnt4g1(const void *, const void *, size_t, void *);
int msm_run_mnt4g2(const void *, const void *, size_t, void *);
int msm_run_mnt6g1(const void *, const void *, size_t, void *);
int msm_run_mnt6g2(const void *, const void *, size_t, void *);

int msm_run_deferred_mnt4g1(const void *, const void *, size_t, void *, MsmTail &);
int msm_run_deferred_mnt4g2(const void *, const void *, size_t, void *, MsmTail &);
int msm_run_deferred_mnt6g1(const void *, const void *, size_t, void *, MsmTail &);
int msm_run_deferred_mnt6g2(const void *, const void *, size_t, void *, MsmTail &);

int msm_dispatch_deferred(int curve, int group, const void *d_scalars, const void *d_points, size_t n, void *h_out,
                          MsmTail &tail) {
  if (curve == 0 && group == 1) return msm_run_deferred_mnt4g1(d_scalars, d_points, n, h_out, tail);
  if (curve == 0 && group == 2) return msm_run_deferred_mnt4g2(d_scalars, d_points, n, h_out, tail);
  if (curve == 1 && group == 1) return msm_run_deferred_mnt6g1(d_scalars, d_points, n, h_out, tail);
  if (curve == 1 && group == 2) return msm_run_deferred_mnt6g2(d_scalars, d_points, n, h_out, tail);
  return set_error(-1, "msm: bad curve/group %d/%d", curve, group);
}

#define B200_DECL_G(name)                                                                                   \
  int msm_precompute_##name(const void *, size_t, MsmPlan &, DevBuf &);                                     \
  int msm_run_table_deferred_##name(const void *, const void *, size_t, const MsmPlan &, void *, MsmTail &, MsmShare, const MsmDedup *);
B200_DECL_G(mnt4g1) B200_DECL_G(mnt4g2) B200_DECL_G(mnt6g1) B200_DECL_G(mnt6g2)
#undef B200_DECL_G

int msm_precompute_dispatch(int curve, int group, const void *d_points, size_t n, MsmPlan &plan, DevBuf &table) {
  if (curve == 0 && group == 1) return msm_precompute_mnt4g1(d_points, n, plan, table);
  if (curve == 0 && group == 2) return msm_precompute_mnt4g2(d_points, n, plan, table);
  if (curve == 1 && group == 1) return msm_precompute_mnt6g1(d_points, n, plan, table);
  if (curve == 1 && group == 2) return msm_precompute_mnt6g2(d_points, n, plan, table);
  return set_error(-1, "msm: bad curve/group %d/%d", curve, group);
}
int msm_table_dispatch_deferred(int curve, int group, const void *d_scalars, const void *d_table, size_t n,
                                const MsmPlan &plan, void *h_out, MsmTail &tail, MsmShare share,
                                const MsmDedup *dedup) {
  if (curve == 0 && group == 1) return msm_run_table_deferred_mnt4g1(d_scalars, d_table, n, plan, h_out, tail, share, dedup);
  if (curve == 0 && group == 2) return msm_run_table_deferred_mnt4g2(d_scalars, d_table, n, plan, h_out, tail, share, dedup);
  if (curve == 1 && group == 1) return msm_run_table_deferred_mnt6g1(d_scalars, d_table, n, plan, h_out, tail, share, dedup);
  if (curve == 1 && group == 2) return msm_run_table_deferred_mnt6g2(d_scalars, d_table, n, plan, h_out, tail, share, dedup);
  return set_error(-1, "msm: bad curve/group %d/%d", curve, group);
}

int msm_dispatch(int curve, int group, const void *d_scalars, const void *d_points, size_t n, void *h_out) {
  if (curve == 0 && group == 1) return msm_run_mnt4g1(d_scalars, d_points, n, h_out);
  if (curve == 0 && group == 2) return msm_run_mnt4g2(d_scalars, d_points, n, h_out);
  if (curve == 1 && group == 1) return msm_run_mnt6g1(d_scalars, d_points, n, h_out);
  if (curve == 1 && group == 2) return msm_run_mnt6g2(d_scalars, d_points, n, h_out);
  return set_error(-1, "msm: bad curve/group %d/%d", curve, group);
}

}  // namespace b200
