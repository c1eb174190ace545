// Group-dependent half of the MSM: bucket accumulation, bucket reduction, window combine. Instantiated once per
// (curve, group) in msm_g_*.cu. See msm.cu for the overall schedule.
#pragma once
#include <chrono>
#include <functional>
#include <memory>
#include "common.cuh"
#include "curve.cuh"
#include "msm.h"
#include "msm_internal.h"

namespace b200 {

// One thread sums one task (a piece of <= T entries of one bucket's list) by mixed addition.
template <class G>
__global__ void __launch_bounds__(128, G::F::kDegree == 1 ? 4 : (G::F::kDegree == 2 ? 4 : 1)) msm_accumulate_kernel(const Affine<typename G::F> *__restrict__ points,
                                                             const uint32_t *__restrict__ entries,
                                                             const uint32_t *__restrict__ offsets,
                                                             const uint32_t *__restrict__ task_off,
                                                             const uint32_t *__restrict__ task_bucket,
                                                             const uint32_t *__restrict__ task_len_sorted,
                                                             const uint32_t *__restrict__ order, uint32_t ntasks,
                                                             uint32_t T, Proj<typename G::F> *__restrict__ partials) {
  typedef typename G::F F;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntasks) return;
  const uint32_t t = order[i];
  const uint32_t b = task_bucket[t];
  const uint32_t len = task_len_sorted[i];
  const uint32_t start = offsets[b] + (t - task_off[b]) * T;
  XYZZ<F> acc;
  xyzz_set_zero(acc);
  for (uint32_t k = 0; k < len; k++) {
    uint32_t e = entries[start + k];
    Affine<F> q = points[e >> 1];
    if (affine_is_zero(q)) continue;
    if (e & 1) F::neg(q.y, q.y);
    xyzz_madd<G>(acc, q);
  }
  Proj<F> out;
  xyzz_to_proj(out, acc);
  partials[t] = out;
}

// One fold level: the sums of one bucket are added in groups of `width`; thread i looks at sum i of the input list and
// works only if it is the first of its group. Keeps the per-bucket combine short when millions of scalars are equal.
template <class G>
__global__ void __launch_bounds__(128) msm_fold_kernel(const Proj<typename G::F> *__restrict__ in,
                                                       const uint32_t *__restrict__ bucket_in,
                                                       const uint32_t *__restrict__ off_in,
                                                       const uint32_t *__restrict__ cnt_in, uint32_t total_in,
                                                       uint32_t width, const uint32_t *__restrict__ off_out,
                                                       Proj<typename G::F> *__restrict__ out,
                                                       uint32_t *__restrict__ bucket_out) {
  typedef typename G::F F;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_in) return;
  const uint32_t b = bucket_in[i];
  const uint32_t s = i - off_in[b];
  if (s % width != 0) return;
  const uint32_t cnt = cnt_in[b];
  const uint32_t len = cnt - s < width ? cnt - s : width;
  Proj<F> acc = in[i];
  for (uint32_t k = 1; k < len; k++) {
    Proj<F> cur = in[i + k];
    proj_add<G>(acc, acc, cur);
  }
  const uint32_t o = off_out[b] + s / width;
  out[o] = acc;
  bucket_out[o] = b;
}

// bucket value = sum of the partial sums of its tasks (usually one: then this is a copy; none: O)
template <class G>
__global__ void __launch_bounds__(128) msm_combine_kernel(const Proj<typename G::F> *__restrict__ partials,
                                                          const uint32_t *__restrict__ task_off,
                                                          const uint32_t *__restrict__ ntasks, uint32_t nbuckets,
                                                          Proj<typename G::F> *__restrict__ buckets) {
  typedef typename G::F F;
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  const uint32_t cnt = ntasks[b], off = task_off[b];
  Proj<F> acc;
  proj_set_zero(acc);
  for (uint32_t s = 0; s < cnt; s++) {
    Proj<F> cur = partials[off + s];
    proj_add<G>(acc, acc, cur);
  }
  buckets[b] = acc;
}

// One thread reduces K consecutive buckets of one window: sum_{v in (lo, lo+K]} v * B_v  (bucket value v = index+1)
template <class G>
__global__ void __launch_bounds__(128) msm_reduce_kernel(const Proj<typename G::F> *__restrict__ buckets, int W,
                                                         uint32_t nb, uint32_t K,
                                                         Proj<typename G::F> *__restrict__ out) {
  typedef typename G::F F;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nchunks = nb / K;
  if (t >= (uint32_t)W * nchunks) return;
  uint32_t j = t / nchunks, q = t % nchunks;
  uint32_t lo = q * K;
  const Proj<F> *B = buckets + (size_t)j * nb;
  Proj<F> run, sum;
  proj_set_zero(run);
  proj_set_zero(sum);
  for (uint32_t k = K; k-- > 0;) {
    Proj<F> cur = B[lo + k];
    proj_add<G>(run, run, cur);
    proj_add<G>(sum, sum, run);
  }
  if (lo != 0) {
    Proj<F> scaled;
    proj_scalar_mul<G>(scaled, run, &lo, 1);
    proj_add<G>(sum, sum, scaled);
  }
  out[t] = sum;
}

// out[j][t] = sum_{r<R} in[j][t*R + r]
template <class G>
__global__ void __launch_bounds__(128) msm_sum_kernel(const Proj<typename G::F> *__restrict__ in, int W, uint32_t per_in,
                                                      uint32_t R, Proj<typename G::F> *__restrict__ out) {
  typedef typename G::F F;
  uint32_t per_out = (per_in + R - 1) / R;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint32_t)W * per_out) return;
  uint32_t j = t / per_out, q = t % per_out;
  Proj<F> acc;
  proj_set_zero(acc);
  for (uint32_t r = 0; r < R; r++) {
    uint32_t idx = q * R + r;
    if (idx >= per_in) break;
    Proj<F> cur = in[(size_t)j * per_in + idx];
    proj_add<G>(acc, acc, cur);
  }
  out[(size_t)j * per_out + q] = acc;
}


// GPU half of one MSM. Returns after the accumulation has finished and the (latency-bound) bucket reduction has been
// ENQUEUED on the workspace's stream; the window sums arrive asynchronously in `stage` (wait on stage->done).
// share_slot >= 0: reuse the scalar-side preparation (digits, sort, tasks) that workspace `share_slot` made for the SAME
// scalars and window plan instead of repeating it (this MSM then only waits for that slot's prep_done event).
template <class G>
int msm_gpu_phase(const void *d_scalars, const void *d_points, size_t n, MsmPlan &plan, MsmWorkspace::Staging *&stage,
                  int share_slot = -1) {
  typedef typename G::F F;
  typedef typename G::ScalarPrime FrP;
  MsmWorkspace &ws = msm_workspace();
  cudaStream_t st = ws.stream;
  MsmWorkspace *prep_ws = &ws;
  if (share_slot >= 0) {
    prep_ws = &msm_workspace_slot(share_slot);
    plan = *prep_ws->prepared;
    B200_CUDA_CHECK(cudaStreamWaitEvent(st, prep_ws->prep_done, 0));
  } else {
    B200_CHECK(msm_prepare(FrP::kTag == 'A' ? 0 : 1, d_scalars, n, plan));
  }
  MsmWorkspace &pw = *prep_ws;  // owner of entries / offsets / task arrays
  const int W = plan.merged ? 1 : plan.W;  // number of independent bucket sets to reduce
  const uint32_t nb = plan.nb;
  const size_t nbuckets = plan.nbuckets;
  B200_CHECK(ws.buckets.reserve(nbuckets * sizeof(Proj<F>)));
  stage = ws.next_staging((size_t)W * sizeof(Proj<F>));
  if (!stage) return set_error(-5, "msm: pinned staging allocation failed");

  // ---- bucket accumulation: one thread per task, then per-bucket combine of the task sums. Nothing below waits
  // on the host: the caller may already prepare the next MSM on the other stream.
  B200_CUDA_CHECK(cudaEventRecord(stage->ta, st));
  B200_CHECK(ws.partials.reserve((plan.ntasks ? plan.ntasks : 1) * sizeof(Proj<F>)));
  if (plan.ntasks) {
    msm_accumulate_kernel<G><<<grid_for(plan.ntasks, 128), 128, 0, st>>>(
        (const Affine<F> *)d_points, pw.entries.as<uint32_t>(), pw.offsets.as<uint32_t>(), pw.task_off.as<uint32_t>(),
        pw.task_bucket.as<uint32_t>(), pw.task_len_sorted.as<uint32_t>(), pw.order.as<uint32_t>(),
        (uint32_t)plan.ntasks, plan.task_len, ws.partials.as<Proj<F>>());
    B200_CUDA_CHECK(cudaGetLastError());
    note_launch();
  }
  // skewed scalars: fold the task sums of heavy buckets in parallel until no bucket has more than kFoldWidth of them
  const Proj<F> *sums = ws.partials.as<Proj<F>>();
  const uint32_t *sum_bucket = pw.task_bucket.as<uint32_t>(), *sum_off = pw.task_off.as<uint32_t>(),
                 *sum_cnt = pw.ntasks.as<uint32_t>();
  {
    size_t total = plan.ntasks;
    uint32_t maxc = plan.max_tasks_per_bucket;
    int half = 0;
    if (maxc > kFoldWidth) {
      const size_t cap = total / kFoldWidth + nbuckets + 1;  // upper bound on the first level's output
      B200_CHECK(ws.fold_cnt.reserve(2 * nbuckets * sizeof(uint32_t)));
      B200_CHECK(ws.fold_off.reserve(2 * nbuckets * sizeof(uint32_t)));
      B200_CHECK(ws.fold_bucket.reserve(2 * cap * sizeof(uint32_t)));
      B200_CHECK(ws.fold_partials.reserve(2 * cap * sizeof(Proj<F>)));
      while (maxc > kFoldWidth) {
        uint32_t *cnt_out = ws.fold_cnt.as<uint32_t>() + half * nbuckets, *off_out = ws.fold_off.as<uint32_t>() + half * nbuckets;
        uint32_t *bucket_out = ws.fold_bucket.as<uint32_t>() + half * cap;
        Proj<F> *out = ws.fold_partials.as<Proj<F>>() + half * cap;
        size_t total_out = 0;
        B200_CHECK(msm_fold_level(sum_cnt, (uint32_t)nbuckets, kFoldWidth, cnt_out, off_out, total_out));
        msm_fold_kernel<G><<<grid_for(total, 128), 128, 0, st>>>(sums, sum_bucket, sum_off, sum_cnt, (uint32_t)total,
                                                                 kFoldWidth, off_out, out, bucket_out);
        B200_CUDA_CHECK(cudaGetLastError());
        note_launch();
        sums = out;
        sum_bucket = bucket_out;
        sum_off = off_out;
        sum_cnt = cnt_out;
        total = total_out;
        maxc = (maxc + kFoldWidth - 1) / kFoldWidth;
        half ^= 1;
      }
    }
  }
  msm_combine_kernel<G><<<grid_for(nbuckets, 128), 128, 0, st>>>(sums, sum_off, sum_cnt, (uint32_t)nbuckets,
                                                                 ws.buckets.as<Proj<F>>());
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();

  // ---- bucket reduction (enqueued, not awaited): chunks of K buckets, then tree sum per bucket set
  B200_CUDA_CHECK(cudaEventRecord(stage->t0, st));
  // chunk length: long chunks amortise the lo*sum fix-up, short ones keep enough threads in flight (>= ~32 K)
  uint32_t K = 32;
  while (K > 2 && (size_t)W * (nb / K) < 32768) K >>= 1;
  if (K > nb) K = nb;
  uint32_t per = nb / K;
  B200_CHECK(ws.red_a.reserve((size_t)W * per * sizeof(Proj<F>)));
  B200_CHECK(ws.red_b.reserve((size_t)W * ((per + 1) / 2) * sizeof(Proj<F>) + 16));
  msm_reduce_kernel<G><<<grid_for((size_t)W * per, 128), 128, 0, st>>>(ws.buckets.as<Proj<F>>(), W, nb, K,
                                                                       ws.red_a.as<Proj<F>>());
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  Proj<F> *cur = ws.red_a.as<Proj<F>>(), *nxt = ws.red_b.as<Proj<F>>();
  while (per > 1) {
    // radix of the tree sum: 8 while a level still fills the GPU, 2 below that (every level then costs one
    // point addition of latency instead of eight)
    uint32_t R = ((size_t)W * per > 262144) ? 8 : 2;
    uint32_t per_out = (per + R - 1) / R;
    msm_sum_kernel<G><<<grid_for((size_t)W * per_out, 128), 128, 0, st>>>(cur, W, per, R, nxt);
    B200_CUDA_CHECK(cudaGetLastError());
    note_launch();
    Proj<F> *t = cur;
    cur = nxt;
    nxt = t;
    per = per_out;
  }
  B200_CUDA_CHECK(cudaMemcpyAsync(stage->pinned, cur, (size_t)W * sizeof(Proj<F>), cudaMemcpyDeviceToHost, st));
  B200_CUDA_CHECK(cudaEventRecord(stage->t1, st));
  B200_CUDA_CHECK(cudaEventRecord(stage->done, st));
  for (int i = 0; i < 2; i++) g_msm_phase_total[F::kDegree == 1 ? 0 : 1][i] += g_msm_phase_ms[i];
  return 0;
}

// wait for the window sums of an MSM issued by msm_gpu_phase and fetch them
template <class G>
int msm_collect(const MsmPlan &plan, MsmWorkspace::Staging *stage, std::vector<Proj<typename G::F>> &win) {
  typedef typename G::F F;
  B200_CUDA_CHECK(cudaEventSynchronize(stage->done));
  const int W = plan.merged ? 1 : plan.W;
  win.resize(W);
  memcpy(win.data(), stage->pinned, (size_t)W * sizeof(Proj<F>));
  float ms = 0;
  cudaEventElapsedTime(&ms, stage->ta, stage->t0);
  g_msm_phase_ms[2] = ms;
  g_msm_phase_total[F::kDegree == 1 ? 0 : 1][2] += ms;
  cudaEventElapsedTime(&ms, stage->t0, stage->t1);
  g_msm_phase_ms[3] = ms;
  g_msm_phase_total[F::kDegree == 1 ? 0 : 1][3] += ms;
  return 0;
}

// Host half: result = sum_j 2^(start_j) * S_j (Horner, most significant window first) - 753 serial doublings.
template <class G>
double msm_host_phase(const MsmPlan &plan, const std::vector<Proj<typename G::F>> &win, void *h_out) {
  typedef typename G::F F;
  Proj<F> result;
  proj_set_zero(result);
  auto t0 = std::chrono::steady_clock::now();
  if (plan.merged) result = win[0];  // pre-shifted bases: the single bucket set already carries the 2^start_j weights
  for (int j = plan.merged ? -1 : plan.W - 1; j >= 0; j--) {
    if (!proj_is_zero(result))
      for (uint32_t k = 0; k < (plan.windows[j] >> 16); k++) proj_dbl<G>(result, result);
    proj_add<G>(result, result, win[j]);
  }
  memcpy(h_out, &result, sizeof(result));
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

template <class G>
int msm_run(const void *d_scalars, const void *d_points, size_t n, void *h_out) {
  typedef typename G::F F;
  if (n == 0) {
    Proj<F> zero;
    proj_set_zero(zero);
    memcpy(h_out, &zero, sizeof(zero));
    return 0;
  }
  MsmPlan plan;
  MsmWorkspace::Staging *stage = nullptr;
  std::vector<Proj<F>> win;
  B200_CHECK(msm_gpu_phase<G>(d_scalars, d_points, n, plan, stage));
  B200_CHECK(msm_collect<G>(plan, stage, win));
  g_msm_phase_ms[4] = msm_host_phase<G>(plan, win, h_out);
  g_msm_phase_total[F::kDegree == 1 ? 0 : 1][4] += g_msm_phase_ms[4];
  return 0;
}

// Same sum, but the wait for the bucket reduction and the serial host tail are returned as a closure, so that the
// caller can run them on another thread while the next MSM already occupies the GPU (b200_prove does this for its five
// MSMs, alternating the two workspaces/streams).
template <class G>
int msm_run_deferred(const void *d_scalars, const void *d_points, size_t n, void *h_out, std::function<void()> &tail) {
  typedef typename G::F F;
  if (n == 0) {
    Proj<F> zero;
    proj_set_zero(zero);
    memcpy(h_out, &zero, sizeof(zero));
    tail = []() {};
    return 0;
  }
  auto plan = std::make_shared<MsmPlan>();
  MsmWorkspace::Staging *stage = nullptr;
  B200_CHECK(msm_gpu_phase<G>(d_scalars, d_points, n, *plan, stage));
  tail = [plan, stage, h_out]() {
    std::vector<Proj<F>> win;
    if (msm_collect<G>(*plan, stage, win) == 0) g_msm_phase_ms[4] = msm_host_phase<G>(*plan, win, h_out);
  };
  return 0;
}

// ---- pre-shifted bases ------------------------------------------------------------------------------------------
// table[j*n + i] = 2^(start_j) * P_i in affine wire format, for the W windows of `plan`. One thread per base: W-1 runs
// of doublings in projective coordinates, then ONE field inversion for all W points (Montgomery's trick).
// With this table every window's digits land in one shared bucket set (weights are baked into the bases), so the
// bucket reduction runs once instead of W times and the serial 753-doubling window combine disappears; the saved
// work also moves the optimal window width up (fewer windows). The table depends only on the proving key: it is
// built when the key is loaded, like the reference parses and stores the key before its timer starts (main.cpp:200-203).
template <class G, int MAXW>
__global__ void __launch_bounds__(128) msm_precompute_kernel(const Affine<typename G::F> *__restrict__ points, uint32_t n,
                                                             int W, const uint32_t *__restrict__ plan,
                                                             Affine<typename G::F> *__restrict__ table) {
  typedef typename G::F F;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> a = points[i];
  if (affine_is_zero(a)) {
    for (int j = 0; j < W; j++) table[(size_t)j * n + i] = a;
    return;
  }
  F zs[MAXW], prefix[MAXW];
  Proj<F> cur;
  proj_from_affine(cur, a);
  for (int j = 0; j < W; j++) {
    // stash X, Y in the output slot, keep Z and the running product of Z's
    Affine<F> xy;
    xy.x = cur.X;
    xy.y = cur.Y;
    table[(size_t)j * n + i] = xy;
    zs[j] = cur.Z;
    if (j == 0) prefix[0] = cur.Z;
    else F::mul(prefix[j], prefix[j - 1], cur.Z);
    if (j + 1 < W) {
      uint32_t width = plan[j] >> 16;
      for (uint32_t k = 0; k < width; k++) proj_dbl<G>(cur, cur);
    }
  }
  F inv;
  F::inv(inv, prefix[W - 1]);
  for (int j = W - 1; j >= 0; j--) {
    F zinv;
    if (j > 0) F::mul(zinv, inv, prefix[j - 1]);
    else zinv = inv;
    F::mul(inv, inv, zs[j]);
    Affine<F> xy = table[(size_t)j * n + i];
    F::mul(xy.x, xy.x, zinv);
    F::mul(xy.y, xy.y, zinv);
    table[(size_t)j * n + i] = xy;
  }
}

template <class G>
int msm_precompute(const void *d_points, size_t n, MsmPlan &plan, DevBuf &table) {
  typedef typename G::F F;
  B200_CHECK(msm_make_plan(n, true, plan));
  if (plan.W > 96) return set_error(-2, "msm_precompute: %d windows exceed the kernel's limit", plan.W);
  B200_CHECK(table.alloc((size_t)plan.W * n * sizeof(Affine<F>)));
  DevBuf dplan;
  B200_CHECK(dplan.alloc(plan.W * sizeof(uint32_t)));
  B200_CUDA_CHECK(cudaMemcpy(dplan.p, plan.windows.data(), plan.W * sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (plan.W <= 48)
    msm_precompute_kernel<G, 48><<<grid_for(n, 128), 128>>>((const Affine<F> *)d_points, (uint32_t)n, plan.W,
                                                            dplan.as<uint32_t>(), table.as<Affine<F>>());
  else
    msm_precompute_kernel<G, 96><<<grid_for(n, 128), 128>>>((const Affine<F> *)d_points, (uint32_t)n, plan.W,
                                                            dplan.as<uint32_t>(), table.as<Affine<F>>());
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  B200_CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
}

// MSM over a table built by msm_precompute (plan must be the table's plan).
template <class G>
int msm_run_table_deferred(const void *d_scalars, const void *d_table, size_t n, const MsmPlan &table_plan, void *h_out,
                           std::function<void()> &tail, int share_slot) {
  typedef typename G::F F;
  if (n == 0) {
    Proj<F> zero;
    proj_set_zero(zero);
    memcpy(h_out, &zero, sizeof(zero));
    tail = []() {};
    return 0;
  }
  auto plan = std::make_shared<MsmPlan>(table_plan);
  MsmWorkspace::Staging *stage = nullptr;
  B200_CHECK(msm_gpu_phase<G>(d_scalars, d_table, n, *plan, stage, share_slot));
  tail = [plan, stage, h_out]() {
    std::vector<Proj<F>> win;
    if (msm_collect<G>(*plan, stage, win) == 0) g_msm_phase_ms[4] = msm_host_phase<G>(*plan, win, h_out);
  };
  return 0;
}

}  // namespace b200
