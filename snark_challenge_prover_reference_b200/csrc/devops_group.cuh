// Group-dependent test/bench kernels (element-wise group ops, synthetic base generator); instantiated per (curve, group)
// in devops_g_*.cu so that they compile in parallel.
#pragma once
#include "common.cuh"
#include "curve.cuh"

namespace b200 {

template <class G>
__global__ void __launch_bounds__(128) group_op_kernel(int op, const void *__restrict__ p, const void *__restrict__ q,
                                                       void *__restrict__ r, size_t n) {
  typedef typename G::F F;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Proj<F> a = ((const Proj<F> *)p)[i], c;
  if (op == 0) {
    Proj<F> b = ((const Proj<F> *)q)[i];
    proj_add<G>(c, a, b);
    ((Proj<F> *)r)[i] = c;
  } else if (op == 1) {
    proj_dbl<G>(c, a);
    ((Proj<F> *)r)[i] = c;
  } else if (op == 2) {
    Affine<F> b = ((const Affine<F> *)q)[i];
    if (!affine_is_zero(b)) proj_madd<G>(a, b);
    ((Proj<F> *)r)[i] = a;
  } else {
    Affine<F> o;
    proj_to_affine<G>(o, a);
    ((Affine<F> *)r)[i] = o;
  }
}

// ---- generators ------------------------------------------------------------------------------------------------
template <class G> struct GenOf;
template <> struct GenOf<Mnt4G1> {
  B200_HD static void get(Affine<Mnt4G1::F> &g) {
    for (int i = 0; i < kLimbs; i++) { g.x.l[i] = MNT4753Gen::g1x(i); g.y.l[i] = MNT4753Gen::g1y(i); }
  }
};
template <> struct GenOf<Mnt6G1> {
  B200_HD static void get(Affine<Mnt6G1::F> &g) {
    for (int i = 0; i < kLimbs; i++) { g.x.l[i] = MNT6753Gen::g1x(i); g.y.l[i] = MNT6753Gen::g1y(i); }
  }
};
template <> struct GenOf<Mnt4G2> {
  B200_HD static void get(Affine<Mnt4G2::F> &g) {
    for (int i = 0; i < kLimbs; i++) {
      g.x.c0.l[i] = MNT4753Gen::g2x0(i); g.x.c1.l[i] = MNT4753Gen::g2x1(i);
      g.y.c0.l[i] = MNT4753Gen::g2y0(i); g.y.c1.l[i] = MNT4753Gen::g2y1(i);
    }
  }
};
template <> struct GenOf<Mnt6G2> {
  B200_HD static void get(Affine<Mnt6G2::F> &g) {
    for (int i = 0; i < kLimbs; i++) {
      g.x.c0.l[i] = MNT6753Gen::g2x0(i); g.x.c1.l[i] = MNT6753Gen::g2x1(i); g.x.c2.l[i] = MNT6753Gen::g2x2(i);
      g.y.c0.l[i] = MNT6753Gen::g2y0(i); g.y.c1.l[i] = MNT6753Gen::g2y1(i); g.y.c2.l[i] = MNT6753Gen::g2y2(i);
    }
  }
};

// out[i] = (first + i) * G, affine wire format. One thread walks a run of L consecutive multiples.
template <class G>
__global__ void __launch_bounds__(128) gen_points_kernel(Affine<typename G::F> *__restrict__ out, size_t n,
                                                         unsigned long long first, uint32_t L) {
  typedef typename G::F F;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t start = t * L;
  if (start >= n) return;
  Affine<F> g;
  GenOf<G>::get(g);
  Proj<F> gp, cur;
  proj_from_affine(gp, g);
  unsigned long long k0 = first + start;
  uint32_t kw[2] = {(uint32_t)k0, (uint32_t)(k0 >> 32)};
  proj_scalar_mul<G>(cur, gp, kw, 2);
  for (uint32_t i = 0; i < L && start + i < n; i++) {
    Affine<F> a;
    proj_to_affine<G>(a, cur);
    out[start + i] = a;
    proj_madd<G>(cur, g);
  }
}

template <class G>
int group_op_t(int op, const void *p, const void *q, void *r, size_t n) {
  group_op_kernel<G><<<grid_for(n, 128), 128>>>(op, p, q, r, n);
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}
template <class G>
int gen_points_t(void *out, size_t n, uint64_t first) {
  const uint32_t L = 8;
  size_t threads = (n + L - 1) / L;
  gen_points_kernel<G><<<grid_for(threads, 128), 128>>>((Affine<typename G::F> *)out, n, first, L);
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace b200
