// Internal interface between msm.cu (scalar side: digits, counting sort - independent of the group) and the four
// per-group translation units msm_g_*.cu (bucket accumulation / reduction kernels, instantiated from msm_group.cuh).
// Split this way so the slow-to-compile group code builds in parallel.
#pragma once
#include <stdint.h>
#include <vector>
#include "common.cuh"

namespace b200 {

struct MsmWorkspace {
  DevBuf digits, counts, offsets, cursor, entries, order, counts_sorted, iota, cub_tmp, buckets, red_a, red_b, plan;
};
MsmWorkspace &msm_workspace();

struct MsmPlan {
  int c = 0, W = 0;
  uint32_t nb = 0;              // buckets per window = 2^(c-1)
  size_t nbuckets = 0;          // W * nb
  std::vector<uint32_t> windows;  // start_bit | width << 16
};

extern double g_msm_phase_ms[5];         // last call: digits, sort, accumulate, reduce, host tail
extern double g_msm_phase_total[2][5];   // accumulated, [0] G1 calls, [1] G2 calls

// Phase 1+2 (group independent): window plan, signed digits, histogram, counting sort of point indices by
// (window, bucket), bucket visiting order by descending size. fr_tag: 0 = modulus A, 1 = modulus B.
int msm_prepare(int fr_tag, const void *d_scalars, size_t n, MsmPlan &plan);

}  // namespace b200
