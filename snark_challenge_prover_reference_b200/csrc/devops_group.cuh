// Group-dependent test/bench kernels (element-wise group ops, synthetic base generator); instantiated per (curve, group)
// in devops_g_*.cu so that they compile in parallel.
#pragma once
#include <string.h>
#include <vector>
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

// ---- fixed-base batch exponentiation (key generation) -----------------------------------------------------------
// out[i] = scalars[i] * g for ONE base g and many scalars: what libff::batch_exp (multiexp.tcc:613-645) does with
// get_window_table (:547-583) and windowed_exp (:585-611), the inner loop of the reference's key generator
// (r1cs_gg_ppzksnark.tcc:289-342: five queries of ~2^20 points each, 10.5 minutes on 8 cores).
// Same decomposition: the scalar is cut into W windows of c bits, table[j][d-1] = d * 2^(jc) * g for d in [1, 2^c).
// B200 schedule: the table is built by msm-style runs (one thread walks L consecutive multiples of 2^(jc) g, then ONE
// inversion brings the run to affine form), kept in affine wire format so that every lookup feeds a mixed addition;
// one thread per scalar accumulates its W lookups in XYZZ coordinates; a last kernel converts runs of results to the
// affine wire format the key files use, again with one inversion per run.
template <class G>
__global__ void __launch_bounds__(128) batch_exp_table_kernel(const Affine<typename G::F> *__restrict__ window_bases,
                                                              int W, uint32_t per_window, uint32_t L,
                                                              Affine<typename G::F> *__restrict__ table) {
  typedef typename G::F F;
  constexpr int kMaxRun = 32;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t runs_per_window = (per_window + L - 1) / L;
  if (t >= (size_t)W * runs_per_window) return;
  const uint32_t j = (uint32_t)(t / runs_per_window), run = (uint32_t)(t % runs_per_window);
  const uint32_t first = run * L;  // table index of the run's first entry: multiple (first + 1)
  const uint32_t cnt = per_window - first < L ? per_window - first : L;
  const Affine<F> g = window_bases[j];
  Affine<F> *out = table + (size_t)j * per_window + first;
  if (affine_is_zero(g)) {
    for (uint32_t i = 0; i < cnt; i++) out[i] = g;
    return;
  }
  Proj<F> gp, cur;
  proj_from_affine(gp, g);
  uint32_t k = first + 1;
  proj_scalar_mul<G>(cur, gp, &k, 1);
  F zs[kMaxRun], prefix[kMaxRun];
  for (uint32_t i = 0; i < cnt; i++) {
    // a multiple of g can be O only if the order of g divides it: never for the prime-order groups used here, but a
    // zero Z must not enter the shared inversion
    const bool inf = proj_is_zero(cur);
    Affine<F> xy;
    xy.x = cur.X;
    xy.y = cur.Y;
    if (inf) {
      F::set_zero(xy.x);
      F::set_zero(xy.y);
      F::set_one(zs[i]);
    } else {
      zs[i] = cur.Z;
    }
    out[i] = xy;
    if (i == 0) prefix[0] = zs[0];
    else F::mul(prefix[i], prefix[i - 1], zs[i]);
    proj_madd<G>(cur, g);
  }
  F inv;
  F::inv(inv, prefix[cnt - 1]);
  for (uint32_t i = cnt; i-- > 0;) {
    F zi;
    if (i > 0) F::mul(zi, inv, prefix[i - 1]);
    else zi = inv;
    F::mul(inv, inv, zs[i]);
    Affine<F> xy = out[i];
    F::mul(xy.x, xy.x, zi);
    F::mul(xy.y, xy.y, zi);
    out[i] = xy;
  }
}

template <class G>
__global__ void __launch_bounds__(128, G::F::kDegree == 3 ? 2 : 4) batch_exp_kernel(
    const Fp<typename G::ScalarPrime> *__restrict__ scalars, size_t n, int c, int W, uint32_t per_window,
    const Affine<typename G::F> *__restrict__ table, Proj<typename G::F> *__restrict__ out) {
  typedef typename G::F F;
  typedef Fp<typename G::ScalarPrime> Fr;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = scalars[i];
  Fr::from_mont(s, s);  // as_bigint (multiexp.tcc:591)
  XYZZ<F> acc;
  xyzz_set_zero(acc);
  for (int j = 0; j < W; j++) {
    const uint32_t bitpos = (uint32_t)j * (uint32_t)c, word = bitpos >> 5, off = bitpos & 31;
    uint64_t two = word < (uint32_t)kLimbs ? s.l[word] : 0u;
    if (word + 1 < (uint32_t)kLimbs) two |= (uint64_t)s.l[word + 1] << 32;
    const uint32_t d = (uint32_t)(two >> off) & ((1u << c) - 1);
    if (d == 0) continue;
    const Affine<F> q = table[(size_t)j * per_window + (d - 1)];
    if (affine_is_zero(q)) continue;
    xyzz_madd<G>(acc, q);
  }
  Proj<F> r;
  xyzz_to_proj(r, acc);
  out[i] = r;
}

// runs of L projective points -> affine wire format with one inversion per run (O -> (0, 0))
template <class G>
__global__ void __launch_bounds__(128) proj_to_affine_runs_kernel(const Proj<typename G::F> *__restrict__ in, size_t n,
                                                                  uint32_t L, Affine<typename G::F> *__restrict__ out) {
  typedef typename G::F F;
  constexpr int kMaxRun = 32;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t first = t * L;
  if (first >= n) return;
  const uint32_t cnt = n - first < L ? (uint32_t)(n - first) : L;
  F prefix[kMaxRun];
  F one;
  F::set_one(one);
  for (uint32_t i = 0; i < cnt; i++) {
    const Proj<F> p = in[first + i];
    const F &z = proj_is_zero(p) ? one : p.Z;
    if (i == 0) prefix[0] = z;
    else F::mul(prefix[i], prefix[i - 1], z);
  }
  F inv;
  F::inv(inv, prefix[cnt - 1]);
  for (uint32_t i = cnt; i-- > 0;) {
    const Proj<F> p = in[first + i];
    Affine<F> xy;
    if (proj_is_zero(p)) {
      F::set_zero(xy.x);
      F::set_zero(xy.y);  // inv is unchanged: this point contributed a factor 1
    } else {
      F zi;
      if (i > 0) F::mul(zi, inv, prefix[i - 1]);
      else zi = inv;
      F::mul(inv, inv, p.Z);
      F::mul(xy.x, p.X, zi);
      F::mul(xy.y, p.Y, zi);
    }
    out[first + i] = xy;
  }
}

// window width: table cost W * 2^c additions against n * W lookups -> about log2(n) - 2, between 4 and 16 bits
static inline int batch_exp_window(size_t n) {
  int c = 4;
  while (c < 16 && ((size_t)1 << (c + 3)) <= n) c++;
  return c;
}

template <class G>
int batch_exp_t(const void *h_base_affine, const void *d_scalars, size_t n, void *d_out_affine, int window, double *ms3) {
  typedef typename G::F F;
  if (n == 0) return 0;
  const int c = window > 0 ? window : batch_exp_window(n);
  if (c < 2 || c > 20) return set_error(-1, "batch_exp: window %d not in [2, 20]", c);
  const int W = (753 + c - 1) / c;
  const uint32_t per_window = (1u << c) - 1;
  // window bases 2^(jc) * g on the host (753 doublings; the same formulas as the device)
  std::vector<Affine<F>> bases(W);
  {
    Affine<F> g;
    memcpy(&g, h_base_affine, sizeof(g));
    Proj<F> cur;
    proj_from_affine(cur, g);
    for (int j = 0; j < W; j++) {
      proj_to_affine<G>(bases[j], cur);
      if (j + 1 < W)
        for (int k = 0; k < c; k++) proj_dbl<G>(cur, cur);
    }
  }
  DevBuf d_bases, table, proj;
  B200_CHECK(d_bases.alloc((size_t)W * sizeof(Affine<F>)));
  B200_CHECK(table.alloc((size_t)W * per_window * sizeof(Affine<F>)));
  B200_CHECK(proj.alloc(n * sizeof(Proj<F>)));
  B200_CUDA_CHECK(cudaMemcpy(d_bases.p, bases.data(), (size_t)W * sizeof(Affine<F>), cudaMemcpyHostToDevice));
  Timer t0, t1, t2;
  const uint32_t L = per_window < 16 ? per_window : 16;
  const size_t table_threads = (size_t)W * ((per_window + L - 1) / L);
  t0.start();
  batch_exp_table_kernel<G><<<grid_for(table_threads, 128), 128>>>(d_bases.as<Affine<F>>(), W, per_window, L, table.as<Affine<F>>());
  B200_CUDA_CHECK(cudaGetLastError());
  t0.stop_async();
  t1.start();
  batch_exp_kernel<G><<<grid_for(n, 128), 128>>>((const Fp<typename G::ScalarPrime> *)d_scalars, n, c, W, per_window,
                                                 table.as<Affine<F>>(), proj.as<Proj<F>>());
  B200_CUDA_CHECK(cudaGetLastError());
  t1.stop_async();
  t2.start();
  const uint32_t R = 16;
  proj_to_affine_runs_kernel<G><<<grid_for((n + R - 1) / R, 128), 128>>>(proj.as<Proj<F>>(), n, R, (Affine<F> *)d_out_affine);
  B200_CUDA_CHECK(cudaGetLastError());
  t2.stop_async();
  note_launch(3);
  B200_CUDA_CHECK(cudaDeviceSynchronize());
  if (ms3) {
    ms3[0] = t0.elapsed();
    ms3[1] = t1.elapsed();
    ms3[2] = t2.elapsed();
  }
  return 0;
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
