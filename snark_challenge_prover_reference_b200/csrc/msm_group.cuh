// Group-dependent half of the MSM: bucket accumulation, bucket reduction, window combine. Instantiated once per
// (curve, group) in msm_g_*.cu. See msm.cu for the overall schedule.
#pragma once
#include <chrono>
#include <functional>
#include <memory>
#include "common.cuh"
#include "coop.cuh"
#include "curve.cuh"
#include "msm.h"
#include "msm_internal.h"

namespace b200 {

// L2 prefetch of a whole object: one prefetch per 128-byte line (objects are 64-byte aligned multiples of 64 bytes, or
// 32-byte aligned field elements). A per-thread instruction: cp.async.bulk.prefetch would be cheaper per byte, but it
// takes warp-uniform operands, and with 32 different addresses the compiler serialises it over the lanes (measured:
// 9.7 % of the round kernel's instructions).
template <class T>
__device__ __forceinline__ void prefetch_l2(const T *p) {
  const char *c = reinterpret_cast<const char *>(p);
#pragma unroll
  for (int off = 0; off < (int)sizeof(T); off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
  if (sizeof(T) % 128 != 0 && sizeof(T) % 64 != 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + sizeof(T) - 1));
}
// streaming (evict-first) copies: the points, the prefix products and the round outputs pass through the caches once;
// the L1 / L2 capacity is needed for the threads' stack frames (the operands of every field operation live there).
// Without the hint the 40 GB that stream through one round evict the 70 MB of frames from the L2, and every field
// operation then waits on DRAM (measured: L2 hit rate 46 %, 23 % of the warps' time in long-scoreboard stalls).
template <class T>
__device__ __forceinline__ void load_streaming(T &dst, const T *src) {
  static_assert(sizeof(T) % 16 == 0, "16-byte granules");
  const uint4 *s = reinterpret_cast<const uint4 *>(src);
  uint4 *d = reinterpret_cast<uint4 *>(&dst);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(T) / 16); k++) d[k] = __ldcs(s + k);
}
template <class T>
__device__ __forceinline__ void store_streaming(T *dst, const T &src) {
  static_assert(sizeof(T) % 16 == 0, "16-byte granules");
  const uint4 *s = reinterpret_cast<const uint4 *>(&src);
  uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(T) / 16); k++) __stcs(d + k, s[k]);
}

// One thread sums one task (a piece of <= T entries of one bucket's list) by mixed addition.
// (G2 over Fq3: 2 blocks per SM cost nothing - the kernel needs 255 registers either way - and double the warps that
// hide each other's latency; round 1 ran it at 1.)
#ifndef B200_ACC_BLOCKS_FQ3
#define B200_ACC_BLOCKS_FQ3 2
#endif
template <class G>
__global__ void __launch_bounds__(128, G::F::kDegree == 3 ? B200_ACC_BLOCKS_FQ3 : 4) msm_accumulate_kernel(const Affine<typename G::F> *__restrict__ points,
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
    if (e == kSkipEntry) continue;  // dropped when a shared entry list was re-indexed for this query (MsmShare)
    Affine<F> q;
    load_streaming(q, points + (e >> 1));  // table gathers pass through once: keep the L2 for the stack frames
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

// ---- batch-affine bucket accumulation (b200_msm_set_batch_affine / B200_BATCH_AFFINE) ------------------------------
// Every bucket's list is summed as a balanced tree: a round adds adjacent pairs of every list (an odd leftover is
// carried over), ceil(log2(longest list)) rounds in all. All additions of a round are independent, so they are done
// in AFFINE coordinates with Montgomery's simultaneous-inversion trick: a thread takes M outputs of the round,
// multiplies their denominators together, inverts ONCE (batched binary gcd, Fp::inv_bingcd: ~75 K ALU instructions,
// no multiplier work) and unwinds. Cost per addition: 3 multiplications for the shared inversion + 3 for
// (lambda, x3, y3) = 5M + 1S instead of the 8M + 2S of the XYZZ mixed addition (and 9M + 2S of the reference's
// mixed_add, mnt4753_g1.cpp:265-313). Special cases are classified per pair: O + Q, P + O, P + P (tangent:
// denominator 2y, numerator 3x^2 + a), P + (-P) = O. Points are kept in wire format ((0,0) = O) plus one O-flag byte
// per point, so that the common path never has to look at a y coordinate to classify.
//
// Schedule (round 2 of the build): thread t of S takes outputs t, t + S, t + 2S, ... - consecutive threads work on
// consecutive outputs, whose operands are consecutive in memory from the second round on (coalesced), and whose
// prefix products pre[j] are written / read back coalesced. The operand pairs of every round come from a small
// bookkeeping kernel (msm_affine_pairs_kernel), so the round kernel carries no bucket-walking state. While output k is
// being computed, the operands of the thread's next output are prefetched into L2 (cp.async.bulk.prefetch): first-round
// operands are gathers from a 7 GB table, and four dependent DRAM latencies per addition were what kept the round-1
// kernel of the previous build at 57 % multiplier-pipe utilisation (profiles/r01_v3_summary.md 2).
#ifndef B200_AFF_BLOCKS
#define B200_AFF_BLOCKS 3
#endif
#ifndef B200_AFF_MIN_M
#define B200_AFF_MIN_M 32
#endif
constexpr uint32_t kAffNone = 0xffffffffu;
// A round's outputs are cut into up to three REGIONS handled by consecutive block ranges of one launch. Inside a
// region thread t takes outputs first + t, first + t + S, ... The first region is one full wave of resident threads
// with most of the outputs (long per-thread batches: the shared inversion is amortised over ~200 additions); the
// following, much shorter regions are scheduled as the first one's blocks retire and fill the tail that a single
// wave of equal batches leaves on a 148-SM chip (measured: 13.4 of 16 warps resident on average, 16 % of the pipe lost).
struct AffRegions {
  uint32_t first[3], count[3], stride[3], block0[3];  // block0: first block of the region
  int n;
};
__device__ __forceinline__ bool aff_my_region(const AffRegions &rg, uint32_t &first, uint32_t &count, uint32_t &S, uint32_t &t) {
  int r = 0;
  if (rg.n > 1 && blockIdx.x >= rg.block0[1]) r = 1;
  if (rg.n > 2 && blockIdx.x >= rg.block0[2]) r = 2;
  first = rg.first[r];
  count = rg.count[r];
  S = rg.stride[r];
  t = (blockIdx.x - rg.block0[r]) * blockDim.x + threadIdx.x;
  return t < S && t < count;
}

// One operand of an addition: index of the stored point (table entry or previous round's output), `neg` says the
// operand is its negative (first round: negative digit), `inf` that it is O.
struct AffOperand {
  uint32_t idx, neg, inf;
};
__device__ __forceinline__ AffOperand aff_operand(const uint8_t *oflag_in, uint32_t v) {
  AffOperand o;
  o.idx = (v & 0x7fffffffu) >> 1;
  o.neg = v & 1u;
  o.inf = oflag_in ? oflag_in[o.idx] : (v >> 31);
  return o;
}
// kind of one output: 0 copy first operand, 1 copy second, 2 result O, 3 chord, 4 tangent; den = the denominator of
// lambda for kinds 3 and 4. p1 / p2 are local copies of the operands (x only is enough unless the x's are equal: `ys`
// says whether the y coordinates have been loaded, and they are fetched here when needed).
// Every case is arranged so that no y coordinate ever has to be negated up front:
//   same flags     : lambda = (y2 - y1)/(x2 - x1), y3 = lambda (x1 - x3) - y1   - the sum of the stored points
//   different flags: lambda = (y1 + y2)/(x2 - x1), y3 = lambda (x3 - x1) - y1   - P1 - P2 of the stored points
// and the result is negated (y3 := y1 - ...) when the FIRST operand carried the flag.
template <class F>
__device__ __forceinline__ int aff_classify(const Affine<F> *src, const AffOperand &a, const AffOperand &b, bool has_second,
                                            Affine<F> &p1, Affine<F> &p2, bool ys, F &den) {
  if (!has_second) return 0;
  if (a.inf) return 1;
  if (b.inf) return 0;
  F::sub(den, p2.x, p1.x);
  if (!F::is_zero(den)) return 3;
  if (!ys) {
    load_streaming(p1.y, &src[a.idx].y);
    load_streaming(p2.y, &src[b.idx].y);
  }
  const bool same_point = F::eq(p1.y, p2.y) == (a.neg == b.neg);
  if (!same_point) return 2;
  F::dbl(den, p1.y);
  return 4;
}

template <class G>
__global__ void __launch_bounds__(128, B200_AFF_BLOCKS) msm_affine_round_kernel(
    const Affine<typename G::F> *__restrict__ src, const uint8_t *__restrict__ oflag_in, const uint2 *__restrict__ pairs,
    AffRegions rg, Affine<typename G::F> *__restrict__ pts_out, uint8_t *__restrict__ oflag_out,
    typename G::F *__restrict__ pre) {
  typedef typename G::F F;
  uint32_t first, total_out, S, t;
  if (!aff_my_region(rg, first, total_out, S, t)) return;
  pairs += first;
  pts_out += first;
  oflag_out += first;
  pre += first;
  // the field elements on the stack; inv holds the running product first; p1 / p2 are the operands of the current output
  F inv, den, lam, num;
  Affine<F> p1, p2;
  F::set_one(inv);
  // ---- forward: running product of the denominators
  uint32_t j = t;
  uint2 nx = pairs[j];
  for (;;) {
    const uint2 pr = nx;
    const uint32_t jn = j + S;
    const bool more = jn < total_out && jn > j;
    if (more) {
      nx = pairs[jn];
      if (nx.y != kAffNone) {
        prefetch_l2(&src[(nx.x & 0x7fffffffu) >> 1].x);
        prefetch_l2(&src[(nx.y & 0x7fffffffu) >> 1].x);
      }
    }
    const bool has2 = pr.y != kAffNone;
    const AffOperand q1 = aff_operand(oflag_in, pr.x);
    const AffOperand q2 = has2 ? aff_operand(oflag_in, pr.y) : q1;
    if (has2 && !q1.inf && !q2.inf) {
      load_streaming(p1.x, &src[q1.idx].x);
      load_streaming(p2.x, &src[q2.idx].x);
    }
    const int kind = aff_classify(src, q1, q2, has2, p1, p2, false, den);
    if (kind >= 3) F::mul(inv, inv, den);
    store_streaming(pre + j, inv);
    if (!more) break;
    j = jn;
  }
  F::inv(inv, inv);
  // ---- backward: unwind the product, finish every addition (j is the thread's last output)
  nx = pairs[j];
  for (;;) {
    const uint2 pr = nx;
    const bool more = j >= S;
    if (more) {
      nx = pairs[j - S];
      prefetch_l2(src + ((nx.x & 0x7fffffffu) >> 1));
      if (nx.y != kAffNone) prefetch_l2(src + ((nx.y & 0x7fffffffu) >> 1));
      if (j >= 2 * S) prefetch_l2(pre + (j - 2 * S));
    }
    const bool has2 = pr.y != kAffNone;
    const AffOperand q1 = aff_operand(oflag_in, pr.x);
    const AffOperand q2 = has2 ? aff_operand(oflag_in, pr.y) : q1;
    // all operand loads of this output are issued together (one memory latency), then everything is local
    load_streaming(p1, src + q1.idx);
    if (has2) load_streaming(p2, src + q2.idx);
    if (more) load_streaming(lam, pre + (j - S));
    const int kind = aff_classify(src, q1, q2, has2, p1, p2, true, den);
    Affine<F> *out = pts_out + j;
    if (kind <= 1) {
      const AffOperand &q = kind == 0 ? q1 : q2;
      Affine<F> &pq = kind == 0 ? p1 : p2;
      if (q.neg && !q.inf) F::neg(pq.y, pq.y);
      store_streaming(out, pq);
      oflag_out[j] = (uint8_t)q.inf;
    } else if (kind == 2) {
      F::set_zero(p1.x);
      F::set_zero(p1.y);
      store_streaming(out, p1);
      oflag_out[j] = 1;
    } else {
      const bool same = kind == 4 || q1.neg == q2.neg;
      if (kind == 3) {
        if (same) F::sub(num, p2.y, p1.y);
        else F::add(num, p1.y, p2.y);
      } else {
        F::sqr(num, p1.x);
        F::add(p2.y, num, num);
        F::add(num, p2.y, num);
        F::set_one(p2.y);
        G::mul_by_a(p2.y, p2.y);
        F::add(num, num, p2.y);  // 3 x^2 + a   (p2.y is free: P2 == P1 here)
      }
      // lambda = num * (inv * prefix) ; then drop this denominator from the running inverse
      // (NOTE: a variant that kept the per-output inverse in its own temporary - dinv = inv * prefix; lambda = num * dinv -
      // was miscompiled by nvcc 12.9 for the MNT6753-G1 instantiation: the PTX passed the SAME stack slot for dinv and
      // num. This ordering needs no such temporary; the tests run every MSM case in both accumulation modes.)
      F::mul(num, num, inv);
      if (more) F::mul(lam, num, lam);  // lam held pre[j - S]
      else lam = num;
      F::mul(inv, inv, den);
      F::sqr(den, lam);                 // den is free from here on: x3 is built in it
      F::sub(den, den, p1.x);
      F::sub(den, den, p2.x);
      if (same) F::sub(num, p1.x, den);
      else F::sub(num, den, p1.x);
      F::mul(num, lam, num);
      if (q1.neg) F::sub(p1.y, p1.y, num);
      else F::sub(p1.y, num, p1.y);
      p1.x = den;
      store_streaming(out, p1);
      oflag_out[j] = 0;
    }
    if (!more) break;
    j -= S;
  }
}

// ---- the same round for groups over the BASE field (G1), with the thread's three live field elements in SHARED
// memory. In the generic kernel above every field operation fetches its operands from the stack frame; 384-512 threads
// x 1.8 KB do not fit the L1, so a quarter of the warps' time went into long-scoreboard stalls in front of every
// multiplication (profiles/r02_summary.md). Three temporaries per thread - the running inverse and two scratch values -
// are all a G1 addition needs when the operands are read in place from global memory (streaming loads inside the
// add / sub routines, prefetched into L2 one output ahead) and x3 is stored as soon as it exists:
//   B = x2 - x1 [den]            A = y2 -/+ y1 ; A *= inv ; A *= pre[j-S] [lambda] ; inv *= B
//   B = A^2 - x1 - x2 [x3] -> out.x ;  B = x1 - B ; B *= A ; A = B - y1 -> out.y
// Layout: slot s of thread t at word ((s * 128 + t) * 28): a 112-byte stride makes the 128-bit accesses of a
// quarter-warp hit 8 different bank groups. 3 x 128 x 112 B = 43 KB per block, 4 blocks per SM.
#ifndef B200_AFF_G1_BLOCKS
#define B200_AFF_G1_BLOCKS 4
#endif
constexpr int kAffSlotStride = 112;  // bytes
constexpr size_t kAffG1Smem = 3 * 128 * kAffSlotStride;

template <class G>
__global__ void __launch_bounds__(128, B200_AFF_G1_BLOCKS) msm_affine_round_g1_kernel(
    const Affine<typename G::F> *__restrict__ src, const uint8_t *__restrict__ oflag_in, const uint2 *__restrict__ pairs,
    AffRegions rg, Affine<typename G::F> *__restrict__ pts_out, uint8_t *__restrict__ oflag_out,
    typename G::F *__restrict__ pre) {
  typedef typename G::F F;
  static_assert(F::kDegree == 1, "base-field groups only");
  extern __shared__ uint4 aff_smem[];
  uint32_t first, total_out, S, t;
  if (!aff_my_region(rg, first, total_out, S, t)) return;
  pairs += first;
  pts_out += first;
  oflag_out += first;
  pre += first;
  char *sm = reinterpret_cast<char *>(aff_smem);
  F &inv = *reinterpret_cast<F *>(sm + (0 * 128 + threadIdx.x) * kAffSlotStride);
  F &A = *reinterpret_cast<F *>(sm + (1 * 128 + threadIdx.x) * kAffSlotStride);
  F &B = *reinterpret_cast<F *>(sm + (2 * 128 + threadIdx.x) * kAffSlotStride);
  F::set_one(inv);
  // classification of one output into B (= den); 0 copy first, 1 copy second, 2 O, 3 chord, 4 tangent
  auto classify = [&](const AffOperand &q1, const AffOperand &q2, bool has2) -> int {
    if (!has2) return 0;
    if (q1.inf) return 1;
    if (q2.inf) return 0;
    F::template sub_g<true, true>(B, src[q2.idx].x, src[q1.idx].x);
    if (!F::is_zero(B)) return 3;
    const bool same_point = F::eq(src[q1.idx].y, src[q2.idx].y) == (q1.neg == q2.neg);
    if (!same_point) return 2;
    F::template add_g<true, true>(B, src[q1.idx].y, src[q1.idx].y);
    return 4;
  };
  // ---- forward: running product of the denominators
  uint32_t j = t;
  uint2 nx = pairs[j];
  for (;;) {
    const uint2 pr = nx;
    const uint32_t jn = j + S;
    const bool more = jn < total_out && jn > j;
    if (more) {
      nx = pairs[jn];
      if (nx.y != kAffNone) {
        prefetch_l2(&src[(nx.x & 0x7fffffffu) >> 1].x);
        prefetch_l2(&src[(nx.y & 0x7fffffffu) >> 1].x);
      }
    }
    const bool has2 = pr.y != kAffNone;
    const AffOperand q1 = aff_operand(oflag_in, pr.x);
    const AffOperand q2 = has2 ? aff_operand(oflag_in, pr.y) : q1;
    const int kind = classify(q1, q2, has2);
    if (kind >= 3) F::mul(inv, inv, B);
    store_streaming(pre + j, inv);
    if (!more) break;
    j = jn;
  }
  F::inv(inv, inv);
  // ---- backward
  nx = pairs[j];
  for (;;) {
    const uint2 pr = nx;
    const bool more = j >= S;
    if (more) {
      nx = pairs[j - S];
      prefetch_l2(src + ((nx.x & 0x7fffffffu) >> 1));
      if (nx.y != kAffNone) prefetch_l2(src + ((nx.y & 0x7fffffffu) >> 1));
      if (j >= 2 * S) prefetch_l2(pre + (j - 2 * S));
    }
    const bool has2 = pr.y != kAffNone;
    const AffOperand q1 = aff_operand(oflag_in, pr.x);
    const AffOperand q2 = has2 ? aff_operand(oflag_in, pr.y) : q1;
    const Affine<F> &p1 = src[q1.idx], &p2 = src[q2.idx];
    const int kind = classify(q1, q2, has2);
    Affine<F> *out = pts_out + j;
    if (kind <= 1) {
      const AffOperand &q = kind == 0 ? q1 : q2;
      const Affine<F> &pq = kind == 0 ? p1 : p2;
      load_streaming(A, &pq.x);
      store_streaming(&out->x, A);
      load_streaming(A, &pq.y);
      if (q.neg && !q.inf) F::neg(A, A);
      store_streaming(&out->y, A);
      oflag_out[j] = (uint8_t)q.inf;
    } else if (kind == 2) {
      F::set_zero(A);
      store_streaming(&out->x, A);
      store_streaming(&out->y, A);
      oflag_out[j] = 1;
    } else {
      const bool same = kind == 4 || q1.neg == q2.neg;
      if (kind == 3) {
        if (same) F::template sub_g<true, true>(A, p2.y, p1.y);
        else F::template add_g<true, true>(A, p1.y, p2.y);
      } else {
        // 3 x^2 + a ; inv must survive, B holds den = 2y: build the numerator in A with the output slot as scratch
        F &T = out->x;  // (global scratch: the tangent case is rare - a duplicated base meeting itself)
        load_streaming(A, &p1.x);
        F::sqr(A, A);
        F::add(T, A, A);
        F::add(A, T, A);
        F::set_one(T);
        G::mul_by_a(T, T);
        F::add(A, A, T);
      }
      F::mul(A, A, inv);                       // num * inv
      if (more) F::mul_bg(A, A, pre[j - S]);   // lambda
      F::mul(inv, inv, B);                     // drop this denominator from the running inverse
      F::sqr(B, A);
      F::template sub_g<false, true>(B, B, p1.x);
      F::template sub_g<false, true>(B, B, p2.x);  // x3
      store_streaming(&out->x, B);
      if (same) F::template sub_g<true, false>(B, p1.x, B);
      else F::template sub_g<false, true>(B, B, p1.x);
      F::mul(B, A, B);
      if (q1.neg) F::template sub_g<true, false>(A, p1.y, B);
      else F::template sub_g<false, true>(A, B, p1.y);
      store_streaming(&out->y, A);
      oflag_out[j] = 0;
    }
    if (!more) break;
    j -= S;
  }
}

// bucket[b] = the single remaining point of list b (or O), converted to the projective form the reduction uses
template <class G>
__global__ void __launch_bounds__(128) msm_affine_finish_kernel(const Affine<typename G::F> *__restrict__ src,
                                                                const uint8_t *__restrict__ oflag_in,
                                                                const uint32_t *__restrict__ entries,
                                                                const uint8_t *__restrict__ base_is_O, uint32_t n_bases,
                                                                const uint32_t *__restrict__ off_in,
                                                                const uint32_t *__restrict__ cnt_in, uint32_t nbuckets,
                                                                Proj<typename G::F> *__restrict__ buckets) {
  typedef typename G::F F;
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  Proj<F> out;
  proj_set_zero(out);
  if (cnt_in[b] != 0) {
    // no round ran at all (every list has <= 1 entry): the point still is a first-round entry
    uint32_t idx = off_in[b], neg = 0, inf;
    if (entries) {
      const uint32_t e = entries[idx];
      idx = e >> 1;
      neg = e & 1u;
      inf = e == kSkipEntry ? 1u : base_is_O[idx % n_bases];
    } else {
      inf = oflag_in[idx];
    }
    if (!inf) {
      Affine<F> p = src[idx];
      if (neg) F::neg(p.y, p.y);
      proj_from_affine(out, p);
    }
  }
  buckets[b] = out;
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

// ---- bucket reduction without per-chunk scalar multiplications (one bucket set: pre-shifted bases) --------------------
// sum_v v * B_v over nb = per * K buckets (bucket i has value i + 1). A thread (coop.cuh: a lane group) takes K consecutive
// buckets:  run_q = sum_k B[qK + k],  sum_q = sum_k (k + 1) B[qK + k]  (2K additions by running sums), and the total is
//   sum_q sum_q  +  K * sum_q q * run_q.
// msm_reduce_kernel computes q * run_q per chunk by double-and-add: ~28 more point operations per chunk. Here the second
// sum is taken bit by bit of q:  sum_q q * run_q = sum_b 2^b O_b,  O_b = sum of run_q over the q with bit b set, and the
// O_b fall out of ONE pairwise tree over the runs: level b adds neighbours, X_{b+1}[i] = X_b[2i] + X_b[2i+1]; the odd
// elements X_b[2i+1] are exactly the partial sums over the q with bit b set, so they open a new row that the following
// levels halve like every other row. Rows = {X, S (the sum_q), O_0, O_1, ...}: a step halves all rows and opens one; about
// 3 additions per chunk in total, log2(per) steps of one addition each. The host finishes
//   S + K * (O_0 + 2 (O_1 + 2 (...)))   (log2(nb) doublings, msm_host_phase).
#ifndef B200_RED_BLOCKS_G2
// resident blocks per SM asked of the G2 instantiations; measured at 2^20 buckets: ptxas' own choice (255 registers)
// 20.5 ms, 3 blocks 18.9 ms, 4 blocks 20.5 ms
#define B200_RED_BLOCKS_G2 3
#endif
template <class G>
__global__ void __launch_bounds__(128, G::F::kDegree == 1 ? 4 : B200_RED_BLOCKS_G2) msm_reduce_rows_kernel(const Proj<typename G::F> *__restrict__ buckets, uint32_t per,
                                                              uint32_t K, Proj<typename G::F> *__restrict__ rows) {
  typedef typename G::F F;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per) return;
  const Proj<F> *B = buckets + (size_t)t * K;
  Proj<F> run, sum;
  proj_set_zero(run);
  proj_set_zero(sum);
  for (uint32_t k = K; k-- > 0;) {
    Proj<F> cur = B[k];
    proj_add<G>(run, run, cur);
    proj_add<G>(sum, sum, run);
  }
  rows[t] = run;                 // row 0: X_0
  rows[(size_t)per + t] = sum;   // row 1: S
}
// `rows` rows of `len` points -> rows + 1 rows of len / 2 points
template <class G>
__global__ void __launch_bounds__(128, G::F::kDegree == 1 ? 4 : B200_RED_BLOCKS_G2) msm_planes_step_kernel(const Proj<typename G::F> *__restrict__ in, uint32_t rows,
                                                              uint32_t len, Proj<typename G::F> *__restrict__ out) {
  typedef typename G::F F;
  const uint32_t half = len / 2;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (rows + 1) * half) return;
  const uint32_t r = t / half, i = t % half;
  if (r == rows) {  // the new row: the odd elements of X
    out[t] = in[2 * i + 1];
    return;
  }
  Proj<F> a = in[(size_t)r * len + 2 * i], b = in[(size_t)r * len + 2 * i + 1];
  proj_add<G>(a, a, b);
  out[t] = a;
}

// Bucket accumulation, default mode: one thread per task (XYZZ mixed additions), parallel fold of heavy buckets,
// per-bucket combine of the task sums. `pw` owns the entry lists / task arrays (this MSM's workspace or the one it
// shares its scalar preparation with). Nothing here waits on the host.
template <class G>
int msm_accumulate_xyzz(const void *d_points, const MsmPlan &plan, MsmWorkspace &ws, MsmWorkspace &pw,
                        const uint32_t *entries) {
  typedef typename G::F F;
  cudaStream_t st = ws.stream;
  const size_t nbuckets = plan.nbuckets;
  B200_CHECK(ws.partials.reserve((plan.ntasks ? plan.ntasks : 1) * sizeof(Proj<F>)));
  // the long kernel goes to the low-priority stream, fenced on both sides: after the entry lists / task arrays are
  // ready (pw.prep_done; for an own preparation also everything queued before it on `st`), before fold / combine
  B200_CUDA_CHECK(cudaStreamWaitEvent(ws.acc_stream, pw.prep_done, 0));
  if (plan.ntasks) {
    msm_accumulate_kernel<G><<<grid_for(plan.ntasks, 128), 128, 0, ws.acc_stream>>>(
        (const Affine<F> *)d_points, entries, pw.offsets.as<uint32_t>(), pw.task_off.as<uint32_t>(),
        pw.task_bucket.as<uint32_t>(), pw.task_len_sorted.as<uint32_t>(), pw.order.as<uint32_t>(),
        (uint32_t)plan.ntasks, plan.task_len, ws.partials.as<Proj<F>>());
    B200_CUDA_CHECK(cudaGetLastError());
    note_launch();
  }
  B200_CUDA_CHECK(cudaEventRecord(ws.acc_done, ws.acc_stream));
  B200_CUDA_CHECK(cudaStreamWaitEvent(st, ws.acc_done, 0));
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
  return 0;
}

// which round kernel a group uses: base-field groups take the shared-memory one (-DB200_AFF_G1_GENERIC: the generic one)
template <class G, int DEG = G::F::kDegree>
struct AffineRound {
  typedef typename G::F F;
  typedef void (*Kernel)(const Affine<F> *, const uint8_t *, const uint2 *, AffRegions, Affine<F> *, uint8_t *, F *);
  static constexpr size_t kSmem = 0;
  static Kernel kernel() { return msm_affine_round_kernel<G>; }
};
#if !defined(B200_AFF_G1_GENERIC)
template <class G>
struct AffineRound<G, 1> {
  typedef typename G::F F;
  typedef void (*Kernel)(const Affine<F> *, const uint8_t *, const uint2 *, AffRegions, Affine<F> *, uint8_t *, F *);
  static constexpr size_t kSmem = kAffG1Smem;
  static Kernel kernel() { return msm_affine_round_g1_kernel<G>; }
};
#endif

// Bucket accumulation by rounds of batched affine additions (see msm_affine_round_kernel). n_bases = the MSM's n:
// the first n entries of d_points are the bases themselves (window 0 of a table, or the plain query).
template <class G>
int msm_accumulate_batch_affine(const void *d_points, size_t n_bases, const MsmPlan &plan, MsmWorkspace &ws, MsmWorkspace &pw,
                                const uint32_t *entries) {
  typedef typename G::F F;
  cudaStream_t st = ws.stream;
  const size_t nbuckets = plan.nbuckets;
  std::vector<size_t> totals, pair_off;
  // bookkeeping (levels, operand pairs, O flags of the bases) on the high-priority preparation stream
  B200_CUDA_CHECK(cudaStreamWaitEvent(ws.prep_stream, pw.prep_done, 0));
  {
    // the previous MSM of this workspace may still read the pair / flag arrays
    B200_CUDA_CHECK(cudaEventRecord(ws.aff_ready, st));
    B200_CUDA_CHECK(cudaStreamWaitEvent(ws.prep_stream, ws.aff_ready, 0));
  }
  B200_CHECK(msm_base_flags(d_points, n_bases, sizeof(Affine<F>), ws.base_flags, ws.prep_stream));
  B200_CHECK(msm_affine_levels(pw.counts.as<uint32_t>(), pw.offsets.as<uint32_t>(), (uint32_t)nbuckets, plan.max_count, totals));
  const int rounds = (int)totals.size() - 1;
  const uint32_t *cnt = ws.aff_cnt.as<uint32_t>(), *off = ws.aff_off.as<uint32_t>();
  B200_CHECK(msm_affine_pairs(ws, cnt, off, (uint32_t)nbuckets, totals, entries, ws.base_flags.as<uint8_t>(), n_bases, pair_off));
  B200_CUDA_CHECK(cudaEventRecord(ws.aff_ready, ws.prep_stream));
  B200_CUDA_CHECK(cudaStreamWaitEvent(st, ws.aff_ready, 0));
  static int wave = 0;  // resident threads of one full wave of the round kernel
  if (!wave) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (AffineRound<G>::kSmem)
      B200_CUDA_CHECK(cudaFuncSetAttribute(AffineRound<G>::kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)AffineRound<G>::kSmem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, AffineRound<G>::kernel(), 128, AffineRound<G>::kSmem);
    wave = (per_sm > 0 ? per_sm : 1) * sms * 128;
  }
  const Affine<F> *src = (const Affine<F> *)d_points;
  const uint8_t *oflag_in = nullptr;
  int done = 0;
  for (int r = 1; r <= rounds; r++) {
    const size_t total_out = totals[r];
    if (total_out == 0) break;
    // regions (see AffRegions): a round with >= 64 outputs per resident thread is cut 80 % / 20 %, each part one wave;
    // smaller rounds are one region with at least B200_AFF_MIN_M outputs per thread (the shared inversion costs as
    // many instructions as ~15 additions)
    AffRegions rg;
    unsigned blocks = 0;
    auto add_region = [&](int r, size_t first, size_t count, size_t S) {
      S = (S + 127) / 128 * 128;
      rg.first[r] = (uint32_t)first;
      rg.count[r] = (uint32_t)count;
      rg.stride[r] = (uint32_t)S;
      rg.block0[r] = blocks;
      blocks += (unsigned)(S / 128);
      rg.n = r + 1;
    };
    rg.n = 0;
    if (msm_affine_split_tail() && total_out / (size_t)wave >= 64) {
      const size_t nA = total_out / 5 * 4;
      add_region(0, 0, nA, (size_t)wave);
      add_region(1, nA, total_out - nA, (size_t)wave);
    } else {
      size_t S = (size_t)wave;
      if (total_out / S < B200_AFF_MIN_M) S = (total_out + B200_AFF_MIN_M - 1) / B200_AFF_MIN_M;
      add_region(0, 0, total_out, S);
    }
    DevBuf &outbuf = ws.aff_pts[r & 1], &oflag = ws.aff_oflag[r & 1];
    B200_CHECK(outbuf.reserve(total_out * sizeof(Affine<F>)));
    B200_CHECK(oflag.reserve(total_out));
    B200_CHECK(ws.aff_scratch.reserve(total_out * sizeof(F)));
    AffineRound<G>::kernel()<<<blocks, 128, AffineRound<G>::kSmem, st>>>(src, oflag_in, ws.aff_pairs.as<uint2>() + pair_off[r], rg,
                                                                     outbuf.as<Affine<F>>(), oflag.as<uint8_t>(),
                                                                     ws.aff_scratch.as<F>());
    B200_CUDA_CHECK(cudaGetLastError());
    note_launch();
    src = outbuf.as<Affine<F>>();
    oflag_in = oflag.as<uint8_t>();
    done = r;
  }
  msm_affine_finish_kernel<G><<<grid_for(nbuckets, 128), 128, 0, st>>>(
      src, oflag_in, done == 0 ? entries : nullptr, ws.base_flags.as<uint8_t>(), (uint32_t)n_bases,
      off + (size_t)done * nbuckets, cnt + (size_t)done * nbuckets, (uint32_t)nbuckets, ws.buckets.as<Proj<F>>());
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}

// GPU half of one MSM. Returns after the accumulation has finished and the (latency-bound) bucket reduction has been
// ENQUEUED on the workspace's stream; the window sums arrive asynchronously in `stage` (wait on stage->done).
// share_slot >= 0: reuse the scalar-side preparation (digits, sort, tasks) that workspace `share_slot` made for the SAME
// scalars and window plan instead of repeating it (this MSM then only waits for that slot's prep_done event).
template <class G>
int msm_gpu_phase(const void *d_scalars, const void *d_points, size_t n, MsmPlan &plan, MsmWorkspace::Staging *&stage,
                  MsmShare share = MsmShare(), const MsmDedup *dedup = nullptr) {
  const int share_slot = share.slot;
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
    B200_CHECK(msm_prepare(FrP::kTag == 'A' ? 0 : 1, d_scalars, n, plan, dedup));  // on ws.prep_stream
    B200_CUDA_CHECK(cudaStreamWaitEvent(st, ws.prep_done, 0));
  }
  MsmWorkspace &pw = *prep_ws;  // owner of entries / offsets / task arrays
  const uint32_t *entries = pw.entries.as<uint32_t>();
  if (share_slot >= 0 && share.n_src) {
    // same buckets, other point numbering: this MSM's own copy of the producer's entry list, re-indexed
    const size_t total = (size_t)plan.W * share.n_src;
    B200_CHECK(ws.entries.reserve(total * sizeof(uint32_t)));
    B200_CHECK(msm_share_entries(entries, total, share.n_src, (uint32_t)n, share.shift, ws.entries.as<uint32_t>(), st));
    if (ws.acc_stream != st) B200_CUDA_CHECK(cudaStreamSynchronize(st));  // (split-stream experiment only)
    entries = ws.entries.as<uint32_t>();
  }
  const int W = plan.merged ? 1 : plan.W;  // number of independent bucket sets to reduce
  const uint32_t nb = plan.nb;
  const size_t nbuckets = plan.nbuckets;
  B200_CHECK(ws.buckets.reserve(nbuckets * sizeof(Proj<F>)));
  stage = ws.next_staging((size_t)(W > 1 ? W : 32) * sizeof(Proj<F>));  // one bucket set: up to 1 + log2(nb) points
  if (!stage) return set_error(-5, "msm: pinned staging allocation failed");
  stage->slot = msm_current_slot();
  for (int i = 0; i < 4; i++) stage->prep_ev[i] = share_slot >= 0 ? nullptr : ws.tm_ev[i];

  // (batch-affine operand words keep bit 31 for the O flag: entry indices must stay below 2^30)
  const int accum_mode = msm_accum_mode();
  const bool affine = (accum_mode == 1 || (accum_mode == 2 && msm_affine_wins(F::kDegree, (size_t)plan.W * n))) &&
                      (size_t)plan.W * n < ((size_t)1 << 30);
  if (affine) {
    B200_CUDA_CHECK(cudaEventRecord(stage->ta, st));
    B200_CHECK(msm_accumulate_batch_affine<G>(d_points, n, plan, ws, pw, entries));
  } else {
    B200_CUDA_CHECK(cudaStreamWaitEvent(ws.acc_stream, pw.prep_done, 0));
    B200_CUDA_CHECK(cudaEventRecord(stage->ta, ws.acc_stream));
    B200_CHECK(msm_accumulate_xyzz<G>(d_points, plan, ws, pw, entries));
  }

  // ---- bucket reduction (enqueued, not awaited): chunks of K buckets, then tree sum per bucket set
  B200_CUDA_CHECK(cudaEventRecord(stage->t0, st));
  Proj<F> *cur = nullptr;
  size_t nout = (size_t)W;  // points copied to the host
  static const bool planes_env = getenv("B200_REDUCE_PLANES") ? atoi(getenv("B200_REDUCE_PLANES")) != 0 : true;
  if (W == 1 && planes_env) {
    // one bucket set: chunk sums + bit-plane tree (msm_reduce_rows_kernel), thread per chunk or lane group per chunk
    const bool coop = msm_use_coop(nb, F::kDegree);
    typedef CoopTables<G> CT;
    constexpr int LG = CT::kLanes, groups_per_block = 128 / LG;
    const size_t smem_red = groups_per_block * coop_group_bytes<G>(3), smem_step = groups_per_block * coop_group_bytes<G>(2);
    static uint32_t resident[2] = {0, 0};  // chunks in flight at once: threads / lane groups of one wave
    if (!resident[coop]) {
      int per_sm = 0, dev = 0, sms = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (coop) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(msm_reduce_rows_coop_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_red));
        B200_CUDA_CHECK(cudaFuncSetAttribute(msm_planes_step_coop_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_step));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msm_reduce_rows_coop_kernel<G>, 128, smem_red);
      } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msm_reduce_rows_kernel<G>, 128, 0);
      }
      resident[coop] = (uint32_t)((per_sm > 0 ? per_sm : 1) * sms * (coop ? groups_per_block : 128));
    }
    // chunk length: the dependent chain is waves * 2K additions + log2(per) tree levels; ties go to the longer chunk
    // ((2K + 3) / K additions per bucket). (A second estimate that also charges the multiplier time of the extra
    // additions of short chunks picked K = 16 instead of 2 for 2^20 G1 buckets - 5.70 against 5.94 ms - but longer chunks
    // for the small bucket sets too: MNT6753 proof 34.1 -> 37.1 ms, a 1/7 share of MNT4753 81.7 -> 82.6 ms. Not kept.)
    uint32_t K = coop ? 4 : 2, best = 0xffffffffu;
    for (uint32_t k = K; k <= 64 && k <= nb; k <<= 1) {
      const uint32_t chunks = nb / k, waves = (chunks + resident[coop] - 1) / resident[coop];
      uint32_t lg = 0;
      while ((1u << lg) < chunks) lg++;
      const uint32_t est = waves * 2 * k + lg;
      if (est <= best) {
        best = est;
        K = k;
      }
    }
    if (K > nb) K = nb;
    const uint32_t per = nb / K;
    B200_CHECK(ws.red_a.reserve((size_t)2 * per * sizeof(Proj<F>)));
    B200_CHECK(ws.red_b.reserve(((size_t)3 * per / 2 + 2) * sizeof(Proj<F>)));
    cur = ws.red_a.as<Proj<F>>();
    Proj<F> *nxt = ws.red_b.as<Proj<F>>();
    if (coop)
      msm_reduce_rows_coop_kernel<G><<<grid_for((size_t)per * LG, 128), 128, smem_red, st>>>(ws.buckets.as<Proj<F>>(), per, K, cur);
    else
      msm_reduce_rows_kernel<G><<<grid_for(per, 128), 128, 0, st>>>(ws.buckets.as<Proj<F>>(), per, K, cur);
    B200_CUDA_CHECK(cudaGetLastError());
    note_launch();
    uint32_t rows = 2, len = per;
    while (len > 1) {
      const size_t outs = (size_t)(rows + 1) * (len / 2);
      if (coop)
        msm_planes_step_coop_kernel<G><<<grid_for(outs * LG, 128), 128, smem_step, st>>>(cur, rows, len, nxt);
      else
        msm_planes_step_kernel<G><<<grid_for(outs, 128), 128, 0, st>>>(cur, rows, len, nxt);
      B200_CUDA_CHECK(cudaGetLastError());
      note_launch();
      Proj<F> *t = cur;
      cur = nxt;
      nxt = t;
      rows++;
      len /= 2;
    }
    plan.red_planes = (int)rows - 2;
    plan.red_K = K;
    nout = (size_t)rows - 1;  // S, O_0 .. O_{planes-1}
    cur += 1;                 // (skip X: the sum of all buckets carries weight 0)
  } else if (msm_use_coop((size_t)W * nb, F::kDegree)) {
    // lane-cooperative reduction (coop.cuh): a group of 8 / 16 / 32 lanes per chunk. The chunk length makes the chunks
    // about one wave of resident groups: short dependent chains, no second wave.
    typedef CoopTables<G> CT;
    constexpr int LG = CT::kLanes, groups_per_block = 128 / LG;
    const size_t smem_red = groups_per_block * coop_group_bytes<G>(4), smem_sum = groups_per_block * coop_group_bytes<G>(2);
    static int resident_groups = 0;
    if (!resident_groups) {
      int per_sm = 0, dev = 0, sms = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      B200_CUDA_CHECK(cudaFuncSetAttribute(msm_reduce_coop_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_red));
      B200_CUDA_CHECK(cudaFuncSetAttribute(msm_sum_coop_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sum));
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msm_reduce_coop_kernel<G>, 128, smem_red);
      resident_groups = (per_sm > 0 ? per_sm : 1) * sms * groups_per_block;
    }
    uint32_t K = 4;
    while (K < 64 && (size_t)W * (nb / K) > (size_t)resident_groups) K <<= 1;
    if (K > nb) K = nb;
    uint32_t per = nb / K;
    B200_CHECK(ws.red_a.reserve((size_t)W * per * sizeof(Proj<F>)));
    B200_CHECK(ws.red_b.reserve((size_t)W * ((per + 1) / 2) * sizeof(Proj<F>) + 16));
    msm_reduce_coop_kernel<G><<<grid_for((size_t)W * per * LG, 128), 128, smem_red, st>>>(ws.buckets.as<Proj<F>>(), W, nb, K,
                                                                                        ws.red_a.as<Proj<F>>());
    B200_CUDA_CHECK(cudaGetLastError());
    note_launch();
    cur = ws.red_a.as<Proj<F>>();
    Proj<F> *nxt = ws.red_b.as<Proj<F>>();
    while (per > 1) {  // pairwise: one cooperative addition of latency per level
      const uint32_t per_out = (per + 1) / 2;
      msm_sum_coop_kernel<G><<<grid_for((size_t)W * per_out * LG, 128), 128, smem_sum, st>>>(cur, W, per, 2, nxt);
      B200_CUDA_CHECK(cudaGetLastError());
      note_launch();
      Proj<F> *t = cur;
      cur = nxt;
      nxt = t;
      per = per_out;
    }
  } else {
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
  cur = ws.red_a.as<Proj<F>>();
  Proj<F> *nxt = ws.red_b.as<Proj<F>>();
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
  }
  B200_CUDA_CHECK(cudaMemcpyAsync(stage->pinned, cur, nout * sizeof(Proj<F>), cudaMemcpyDeviceToHost, st));
  B200_CUDA_CHECK(cudaEventRecord(stage->t1, st));
  B200_CUDA_CHECK(cudaEventRecord(stage->done, st));
  return 0;
}

// wait for the window sums of an MSM issued by msm_gpu_phase and fetch them
template <class G>
int msm_collect(const MsmPlan &plan, MsmWorkspace::Staging *stage, std::vector<Proj<typename G::F>> &win) {
  typedef typename G::F F;
  B200_CUDA_CHECK(cudaEventSynchronize(stage->done));
  const int W = plan.red_planes >= 0 ? 1 + plan.red_planes : plan.merged ? 1 : plan.W;
  win.resize(W);
  memcpy(win.data(), stage->pinned, (size_t)W * sizeof(Proj<F>));
  float ms = 0;
  if (stage->prep_ev[0]) {  // this MSM prepared its own digits / sort (events completed: they precede `done`)
    for (int i = 0; i < 2; i++) {
      cudaEventElapsedTime(&ms, stage->prep_ev[2 * i], stage->prep_ev[2 * i + 1]);
      g_msm_phase_ms[i] = ms;
      msm_stat_add(g_msm_phase_total[F::kDegree == 1 ? 0 : 1][i], ms);
    }
  }
  cudaEventElapsedTime(&ms, stage->ta, stage->t0);
  g_msm_phase_ms[2] = ms;
  msm_stat_add(g_msm_phase_total[F::kDegree == 1 ? 0 : 1][2], ms);
  cudaEventElapsedTime(&ms, stage->t0, stage->t1);
  msm_timeline_note(stage->slot, stage->ta, stage->t0, stage->t1);
  g_msm_phase_ms[3] = ms;
  msm_stat_add(g_msm_phase_total[F::kDegree == 1 ? 0 : 1][3], ms);
  return 0;
}

// Host half: result = sum_j 2^(start_j) * S_j (Horner, most significant window first) - 753 serial doublings.
template <class G>
double msm_host_phase(const MsmPlan &plan, const std::vector<Proj<typename G::F>> &win, void *h_out) {
  typedef typename G::F F;
  Proj<F> result;
  proj_set_zero(result);
  auto t0 = std::chrono::steady_clock::now();
  if (plan.red_planes >= 0) {
    // one bucket set reduced by bit planes: win = {S, O_0, .., O_{planes-1}}, result = S + K * sum_b 2^b O_b
    result = win[0];
    if (plan.red_planes > 0) {
      Proj<F> acc = win[plan.red_planes];
      for (int b = plan.red_planes - 2; b >= 0; b--) {
        proj_dbl<G>(acc, acc);
        proj_add<G>(acc, acc, win[1 + b]);
      }
      for (uint32_t k = plan.red_K; k > 1; k >>= 1) proj_dbl<G>(acc, acc);
      proj_add<G>(result, result, acc);
    }
  } else if (plan.merged) {
    result = win[0];  // pre-shifted bases: the single bucket set already carries the 2^start_j weights
  }
  for (int j = (plan.merged || plan.red_planes >= 0) ? -1 : plan.W - 1; j >= 0; j--) {
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
  msm_stat_add(g_msm_phase_total[F::kDegree == 1 ? 0 : 1][4], g_msm_phase_ms[4]);
  return 0;
}

// The host tail of a deferred MSM. It may run on a thread other than the issuing one: the device that was current
// at issue time is made current there, and a failure (e.g. an asynchronous kernel fault surfacing at the event wait)
// is returned with its message instead of being left in the tail thread's thread-local error slot.
template <class G>
MsmTail msm_make_tail(std::shared_ptr<MsmPlan> plan, MsmWorkspace::Staging *stage, void *h_out) {
  typedef typename G::F F;
  int dev = 0;
  cudaGetDevice(&dev);
  return [plan, stage, h_out, dev](std::string &err) -> int {
    cudaError_t e = cudaSetDevice(dev);
    int rc = e == cudaSuccess ? 0 : set_error(-100 - (int)e, "msm tail: cudaSetDevice(%d): %s", dev, cudaGetErrorString(e));
    std::vector<Proj<F>> win;
    if (rc == 0) rc = msm_collect<G>(*plan, stage, win);
    if (rc) {
      err = last_error();
      return rc;
    }
    g_msm_phase_ms[4] = msm_host_phase<G>(*plan, win, h_out);
    return 0;
  };
}

// Same sum, but the wait for the bucket reduction and the serial host tail are returned as a closure, so that the
// caller can run them on another thread while the next MSM already occupies the GPU (b200_prove does this for its five
// MSMs, alternating the two workspaces/streams).
template <class G>
int msm_run_deferred(const void *d_scalars, const void *d_points, size_t n, void *h_out, MsmTail &tail) {
  typedef typename G::F F;
  if (n == 0) {
    Proj<F> zero;
    proj_set_zero(zero);
    memcpy(h_out, &zero, sizeof(zero));
    tail = [](std::string &) { return 0; };
    return 0;
  }
  auto plan = std::make_shared<MsmPlan>();
  MsmWorkspace::Staging *stage = nullptr;
  B200_CHECK(msm_gpu_phase<G>(d_scalars, d_points, n, *plan, stage));
  tail = msm_make_tail<G>(plan, stage, h_out);
  return 0;
}

// ---- pre-shifted bases ------------------------------------------------------------------------------------------
// table[j*n + i] = 2^(start_j) * P_i in affine wire format, for the W windows of `plan`. One thread per base: W-1 runs
// of doublings in projective coordinates, then ONE field inversion for all W points (Montgomery's trick).
// With this table every window's digits land in one shared bucket set (weights are baked into the bases), so the
// bucket reduction runs once instead of W times and the serial 753-doubling window combine disappears; the saved
// work also moves the optimal window width up (fewer windows). The table depends only on the proving key: it is
// built when the key is loaded, like the reference parses and stores the key before its timer starts (main.cpp:200-203).
// (launch bounds: left alone ptxas takes 216-236 registers - 2 warps per scheduler, the multiplier ~60 % busy; every
// multiplication / squaring body is out of line and fits 128 registers without spills, the kernel itself spills < 1 KB)
#ifndef B200_PRE_BLOCKS_G2
#define B200_PRE_BLOCKS_G2 3
#endif
template <class G, int MAXW>
__global__ void __launch_bounds__(128, G::F::kDegree == 1 ? 4 : B200_PRE_BLOCKS_G2) msm_precompute_kernel(const Affine<typename G::F> *__restrict__ points, uint32_t n,
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
  // Jacobian coordinates (jac_dbl: 1M + 8S per doubling, squarings through the dedicated squaring): the table builder
  // is 753 doublings per base and nothing else
  F zs[MAXW], prefix[MAXW];
  Proj<F> cur;  // (X : Y : Z) read as Jacobian here: x = X/Z^2, y = Y/Z^3
  cur.X = a.x;
  cur.Y = a.y;
  F::set_one(cur.Z);
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
      for (uint32_t k = 0; k < width; k++) jac_dbl<G>(cur, cur);
    }
  }
  F inv;
  F::inv(inv, prefix[W - 1]);
  for (int j = W - 1; j >= 0; j--) {
    F zinv, zinv2;
    if (j > 0) F::mul(zinv, inv, prefix[j - 1]);
    else zinv = inv;
    F::mul(inv, inv, zs[j]);
    Affine<F> xy = table[(size_t)j * n + i];
    F::sqr(zinv2, zinv);
    F::mul(xy.x, xy.x, zinv2);      // X / Z^2
    F::mul(zinv2, zinv2, zinv);
    F::mul(xy.y, xy.y, zinv2);      // Y / Z^3
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
                           MsmTail &tail, MsmShare share, const MsmDedup *dedup) {
  typedef typename G::F F;
  if (n == 0) {
    Proj<F> zero;
    proj_set_zero(zero);
    memcpy(h_out, &zero, sizeof(zero));
    tail = [](std::string &) { return 0; };
    return 0;
  }
  auto plan = std::make_shared<MsmPlan>(table_plan);
  MsmWorkspace::Staging *stage = nullptr;
  B200_CHECK(msm_gpu_phase<G>(d_scalars, d_table, n, *plan, stage, share, dedup));
  tail = msm_make_tail<G>(plan, stage, h_out);
  return 0;
}

}  // namespace b200
