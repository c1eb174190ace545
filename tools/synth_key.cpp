// synth_key - writes a SYNTHETIC proving key + witness in the reference's file formats at any size, fast.
//
//   synth_key <MNT4753|MNT6753> <log2 m> <params-out> <input-out> [seed]
//
// Bench / test infrastructure (not on the prover path): `./generate_parameters` at the challenge size takes ten minutes
// (BASELINE.md 2), and libsnark's ./main never validates that the files form a real key - it evaluates the 7 FFTs and
// 5 MSMs on whatever well-formed files it is given (main.cpp:200-268). So the SAME files feed the
// unmodified reference prover and the B200 prover, and their outputs are compared byte for byte.
//
// File formats (libsnark/serialization.hpp:22-111, generate_parameters.cpp:60-125):
//   params: d (u64) | m (u64) | A[m+1] G1 | B1[m+1] G1 | B2[m+1] G2 | L[m-1] G1 | H[d] G1     (affine, Montgomery, (0,0)=O)
//   input : w[m+1] | ca[d+1] | cb[d+1] | cc[d+1] | r                                             (Fr, Montgomery)
// Contents: base i of a query = (first + i) * G (valid curve points, all distinct), then the structure observed in
// real keys is imposed (SURVEY.md 8 pitfalls 1-2): A has m/2 copies of one point and O at index m; B1 / B2 have O at
// indices 0 and m and one duplicate pair. Scalars: uniformly random 752-bit values (every value < 2^752 < r is the
// Montgomery representation of some element), w[0] = Montgomery one. Deterministic in (curve, log2 m, seed).
// The same multiples of the generator are produced on the device by gen_points_kernel (devops_group.cuh); the GPU test
// suite checks the two against each other.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>
#include "curve.cuh"

using namespace b200;

template <class G> struct GenOf;
template <> struct GenOf<Mnt4G1> {
  static void get(Affine<Mnt4G1::F> &g) {
    for (int i = 0; i < kLimbs; i++) { g.x.l[i] = MNT4753Gen::g1x(i); g.y.l[i] = MNT4753Gen::g1y(i); }
  }
};
template <> struct GenOf<Mnt6G1> {
  static void get(Affine<Mnt6G1::F> &g) {
    for (int i = 0; i < kLimbs; i++) { g.x.l[i] = MNT6753Gen::g1x(i); g.y.l[i] = MNT6753Gen::g1y(i); }
  }
};
template <> struct GenOf<Mnt4G2> {
  static void get(Affine<Mnt4G2::F> &g) {
    for (int i = 0; i < kLimbs; i++) {
      g.x.c0.l[i] = MNT4753Gen::g2x0(i); g.x.c1.l[i] = MNT4753Gen::g2x1(i);
      g.y.c0.l[i] = MNT4753Gen::g2y0(i); g.y.c1.l[i] = MNT4753Gen::g2y1(i);
    }
  }
};
template <> struct GenOf<Mnt6G2> {
  static void get(Affine<Mnt6G2::F> &g) {
    for (int i = 0; i < kLimbs; i++) {
      g.x.c0.l[i] = MNT6753Gen::g2x0(i); g.x.c1.l[i] = MNT6753Gen::g2x1(i); g.x.c2.l[i] = MNT6753Gen::g2x2(i);
      g.y.c0.l[i] = MNT6753Gen::g2y0(i); g.y.c1.l[i] = MNT6753Gen::g2y1(i); g.y.c2.l[i] = MNT6753Gen::g2y2(i);
    }
  }
};

// out[i] = (first + i) * G for i in [lo, hi): one scalar multiplication, then a chain of mixed additions; the
// projective points of a run are brought to affine form with ONE inversion (Montgomery's trick).
template <class G>
static void gen_range(Affine<typename G::F> *out, size_t lo, size_t hi, uint64_t first) {
  typedef typename G::F F;
  Affine<F> g;
  GenOf<G>::get(g);
  Proj<F> gp, cur;
  proj_from_affine(gp, g);
  const uint64_t k0 = first + lo;
  uint32_t kw[2] = {(uint32_t)k0, (uint32_t)(k0 >> 32)};
  proj_scalar_mul<G>(cur, gp, kw, 2);
  const size_t kRun = 256;
  std::vector<Proj<F>> run(kRun);
  std::vector<F> prefix(kRun);
  for (size_t base = lo; base < hi; base += kRun) {
    const size_t cnt = hi - base < kRun ? hi - base : kRun;
    for (size_t i = 0; i < cnt; i++) {
      run[i] = cur;
      if (i == 0) prefix[0] = cur.Z;
      else F::mul(prefix[i], prefix[i - 1], cur.Z);
      proj_madd<G>(cur, g);
    }
    F inv;
    F::inv(inv, prefix[cnt - 1]);
    for (size_t i = cnt; i-- > 0;) {
      F zi;
      if (i > 0) F::mul(zi, inv, prefix[i - 1]);
      else zi = inv;
      F::mul(inv, inv, run[i].Z);
      F::mul(out[base + i].x, run[i].X, zi);
      F::mul(out[base + i].y, run[i].Y, zi);
    }
  }
}

template <class G>
static void gen_query(std::vector<unsigned char> &buf, size_t n, uint64_t first, unsigned nthreads) {
  typedef Affine<typename G::F> A;
  buf.resize(n * sizeof(A));
  A *out = (A *)buf.data();
  std::vector<std::thread> th;
  const size_t per = (n + nthreads - 1) / nthreads;
  for (unsigned t = 0; t < nthreads; t++) {
    const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
    if (lo >= hi) break;
    th.emplace_back([=] { gen_range<G>(out, lo, hi, first); });
  }
  for (auto &x : th) x.join();
}

static uint64_t splitmix64(uint64_t &s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

static void must_write(FILE *f, const void *p, size_t n, const char *path) {
  if (fwrite(p, 1, n, f) != n) {
    fprintf(stderr, "synth_key: short write to %s\n", path);
    exit(2);
  }
}

template <class G1, class G2, class FrP>
static int run(int k, const char *params_path, const char *input_path, uint64_t seed) {
  const size_t m = (size_t)1 << k, d = m - 1;
  unsigned nthreads = std::thread::hardware_concurrency();
  if (nthreads == 0) nthreads = 4;
  const size_t g1 = sizeof(Affine<typename G1::F>), g2 = sizeof(Affine<typename G2::F>);
  std::vector<unsigned char> A, B1, B2, L, H;
  gen_query<G1>(A, m + 1, 1000003, nthreads);
  gen_query<G1>(B1, m + 1, 2000003, nthreads);
  gen_query<G2>(B2, m + 1, 3000017, nthreads);
  gen_query<G1>(L, m - 1, 4000037, nthreads);
  gen_query<G1>(H, d, 5000011, nthreads);
  // structure of real keys (same as bench.py's device-side make_key)
  for (size_t i = 4; i + 1 < m; i += 2) memcpy(&A[i * g1], &A[2 * g1], g1);
  memcpy(&A[(m - 1) * g1], &A[2 * g1], g1);
  memset(&A[m * g1], 0, g1);
  memset(&B1[0], 0, g1);
  memset(&B1[m * g1], 0, g1);
  memcpy(&B1[(m - 2) * g1], &B1[(m - 3) * g1], g1);
  memset(&B2[0], 0, g2);
  memset(&B2[m * g2], 0, g2);
  memcpy(&B2[(m - 2) * g2], &B2[(m - 3) * g2], g2);
  FILE *f = fopen(params_path, "wb");
  if (!f) {
    perror(params_path);
    return 2;
  }
  uint64_t hdr[2] = {(uint64_t)d, (uint64_t)m};
  must_write(f, hdr, 16, params_path);
  must_write(f, A.data(), A.size(), params_path);
  must_write(f, B1.data(), B1.size(), params_path);
  must_write(f, B2.data(), B2.size(), params_path);
  must_write(f, L.data(), L.size(), params_path);
  must_write(f, H.data(), H.size(), params_path);
  fclose(f);
  // input image
  const size_t n = (m + 1) + 3 * (d + 1) + 1;
  std::vector<unsigned char> img(n * 96);
  uint64_t s = seed * 0x2545f4914f6cdd1dull + (uint64_t)k;
  uint64_t *w64 = (uint64_t *)img.data();
  for (size_t i = 0; i < n; i++) {
    for (int j = 0; j < 12; j++) w64[i * 12 + j] = splitmix64(s);
    w64[i * 12 + 11] &= 0x0000ffffffffffffull;  // < 2^752
  }
  Fp<FrP> one;
  Fp<FrP>::set_one(one);
  memcpy(img.data(), &one, 96);
  f = fopen(input_path, "wb");
  if (!f) {
    perror(input_path);
    return 2;
  }
  must_write(f, img.data(), img.size(), input_path);
  fclose(f);
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: synth_key <MNT4753|MNT6753> <log2 m> <params-out> <input-out> [seed]\n");
    return 1;
  }
  const std::string curve = argv[1];
  const int k = atoi(argv[2]);
  const uint64_t seed = argc > 5 ? strtoull(argv[5], nullptr, 0) : 77;
  if (k < 2 || k > 24) {
    fprintf(stderr, "synth_key: log2 m out of range\n");
    return 1;
  }
  if (curve == "MNT4753") return run<Mnt4G1, Mnt4G2, PrimeA>(k, argv[3], argv[4], seed);
  if (curve == "MNT6753") {
    if (k > 15) {
      fprintf(stderr, "synth_key: MNT6753's scalar field has 2-adicity 15\n");
      return 1;
    }
    return run<Mnt6G1, Mnt6G2, PrimeB>(k, argv[3], argv[4], seed);
  }
  fprintf(stderr, "synth_key: unknown curve %s\n", curve.c_str());
  return 1;
}
