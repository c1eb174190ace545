// Radix-2 evaluation domain over Fr (MNT4753: 2-adicity 30, MNT6753: 2-adicity 15) on one B200.
//
// Replaces libfqfft::basic_radix2_domain (depends/libfqfft/libfqfft/evaluation_domain/domains/basic_radix2_domain.tcc:
// FFT 62-68, iFFT 70-82, cosetFFT 84-89, icosetFFT 91-96, divide_by_Z_on_coset 125-134) and its workers
// (_basic_serial_radix2_FFT basic_radix2_domain_aux.tcc:167-202, _multiply_by_coset :321-330) behind
// B::domain_* (libsnark/prover_reference_functions.cpp:222-245). Same transform: natural order in and out,
// out[j] = sum_i a[i] * omega^(i*j), omega = libff's 2^k-th root (field_utils.tcc:40-89).
//
// Schedule: the k butterfly stages (decimation in time after a bit reversal) are cut into ceil(k/8) passes. A pass
// stages a tile of 2^r elements (r <= 8) in shared memory, limb-major with a +1 pad so that both the element-major
// global accesses (96 contiguous bytes per element) and the butterfly accesses are bank-conflict free, runs r
// stages there, and writes the tile back. The bit reversal is folded into the first pass's gather, the coset
// multiplication (a[i] * g^i) into its load, and the 1/m and g^-i factors into the last pass's store, so a transform
// of any kind reads and writes the vector exactly ceil(k/8) times. Twiddles omega^i (i < m/2) are tabulated once per
// domain; the early passes touch only a handful of them (L1/L2 resident).
#include "common.cuh"
#include "field.cuh"
#include "ntt.h"

namespace b200 {

// stream of every launch below for the calling host thread (default: the legacy default stream, which is what the
// public domain_* entry points promise; b200_prove* put compute_H on a high-priority stream of their own)
static thread_local cudaStream_t g_ntt_stream = 0;
void ntt_set_stream(cudaStream_t st) { g_ntt_stream = st; }

// ---------------------------------------------------------------------------------------------- kernels
template <class P>
__global__ void __launch_bounds__(128) powers_kernel(Fp<P> *__restrict__ out, size_t n, const Fp<P> *__restrict__ base_p,
                                                     const Fp<P> *__restrict__ scale_p, uint32_t L) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t start = t * L;
  if (start >= n) return;
  uint32_t e[2] = {(uint32_t)start, (uint32_t)(start >> 32)};
  const Fp<P> base = *base_p, scale = *scale_p;
  Fp<P> cur;
  Fp<P>::pow_words(cur, base, e, 2);
  Fp<P>::mul(cur, cur, scale);
  for (uint32_t i = 0; i < L && start + i < n; i++) {
    out[start + i] = cur;
    Fp<P>::mul(cur, cur, base);
  }
}

template <class P>
__global__ void __launch_bounds__(128) ntt_pass_kernel(const Fp<P> *src, Fp<P> *dst, int k,
                                                       int s0, int r, const Fp<P> *__restrict__ tw, int bitrev_in,
                                                       const Fp<P> *__restrict__ pre_tab,
                                                       const Fp<P> *__restrict__ post_tab,
                                                       const Fp<P> *__restrict__ post_const) {
  extern __shared__ uint32_t sm[];
  const uint32_t T = 1u << r, stride = T + 1;
  const uint32_t u = threadIdx.x;
  const size_t blk = blockIdx.x;
  const size_t lo = blk & (((size_t)1 << s0) - 1), hi = blk >> s0;
  const size_t base = (hi << (s0 + r)) | lo;
#pragma unroll 1
  for (int e = 0; e < 2; e++) {
    uint32_t t = u + e * (T >> 1);
    size_t idx = base | ((size_t)t << s0);
    size_t sidx = bitrev_in ? (size_t)(__brevll((unsigned long long)idx) >> (64 - k)) : idx;
    Fp<P> v = src[sidx];
    if (pre_tab) Fp<P>::mul(v, v, pre_tab[sidx]);
#pragma unroll
    for (int i = 0; i < kLimbs; i++) sm[i * stride + t] = v.l[i];
  }
  __syncthreads();
#pragma unroll 1
  for (int q = 1; q <= r; q++) {
    const uint32_t h = 1u << (q - 1);
    const uint32_t t0 = ((u >> (q - 1)) << q) | (u & (h - 1));
    const uint32_t t1 = t0 + h;
    const size_t j = ((size_t)(t0 & (h - 1)) << s0) | lo;
    const int s = s0 + q;
    Fp<P> a, b, x;
#pragma unroll
    for (int i = 0; i < kLimbs; i++) {
      a.l[i] = sm[i * stride + t0];
      b.l[i] = sm[i * stride + t1];
    }
    Fp<P>::mul(b, b, tw[j << (k - s)]);
    Fp<P>::add(x, a, b);
    Fp<P>::sub(b, a, b);
#pragma unroll
    for (int i = 0; i < kLimbs; i++) {
      sm[i * stride + t0] = x.l[i];
      sm[i * stride + t1] = b.l[i];
    }
    __syncthreads();
  }
#pragma unroll 1
  for (int e = 0; e < 2; e++) {
    uint32_t t = u + e * (T >> 1);
    size_t idx = base | ((size_t)t << s0);
    Fp<P> v;
#pragma unroll
    for (int i = 0; i < kLimbs; i++) v.l[i] = sm[i * stride + t];
    if (post_tab)
      Fp<P>::mul(v, v, post_tab[idx]);
    else if (post_const)
      Fp<P>::mul(v, v, *post_const);
    dst[idx] = v;
  }
}

// pseudo-random field elements (< 2^752, i.e. valid Montgomery representations) for the warm-up proof of a freshly
// loaded key: SplitMix64 of the limb index
__global__ void __launch_bounds__(256) fr_fill_kernel(uint64_t *__restrict__ out, size_t n_limbs, uint64_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_limbs) return;
  uint64_t z = seed + (i + 1) * 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z ^= z >> 31;
  if (i % 12 == 11) z &= 0x0000ffffffffffffull;
  out[i] = z;
}
int fr_fill_pseudo_random(void *d_out, size_t n_elements, uint64_t seed) {
  const size_t limbs = n_elements * 12;
  if (limbs == 0) return 0;
  fr_fill_kernel<<<grid_for(limbs, 256), 256, 0, g_ntt_stream>>>((uint64_t *)d_out, limbs, seed);
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}

template <class P>
__global__ void __launch_bounds__(128) fr_muleq_kernel(Fp<P> *__restrict__ a, const Fp<P> *__restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<P> x = a[i];
  Fp<P>::mul(x, x, b[i]);
  a[i] = x;
}
template <class P>
__global__ void __launch_bounds__(128) fr_subeq_kernel(Fp<P> *__restrict__ a, const Fp<P> *__restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<P> x = a[i], y = b[i];
  Fp<P>::sub(x, x, y);
  a[i] = x;
}
template <class P>
__global__ void __launch_bounds__(128) fr_scale_kernel(Fp<P> *__restrict__ a, const Fp<P> *__restrict__ c, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<P> x = a[i];
  Fp<P>::mul(x, x, *c);
  a[i] = x;
}

// ---------------------------------------------------------------------------------------------- host side
struct Domain {
  int curve;
  size_t m;
  int k;
  DevBuf tw_fwd, tw_inv, coset, icoset_scaled, scratch, consts;  // consts: [0] = 1/m, [1] = 1/Z(g)
};

template <class P>
static void host_pow_u64(Fp<P> &r, const Fp<P> &a, uint64_t e) {
  uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
  Fp<P>::pow_words(r, a, w, 2);
}
template <class P>
static void load_const(Fp<P> &r, uint32_t (*f)(int)) {
  for (int i = 0; i < kLimbs; i++) r.l[i] = f(i);
}

template <class P>
static int fill_powers(DevBuf &buf, size_t n, const Fp<P> &base, const Fp<P> &scale) {
  B200_CHECK(buf.alloc(n * sizeof(Fp<P>)));
  DevBuf args;
  B200_CHECK(args.alloc(2 * sizeof(Fp<P>)));
  Fp<P> hargs[2] = {base, scale};
  B200_CUDA_CHECK(cudaMemcpy(args.p, hargs, sizeof(hargs), cudaMemcpyHostToDevice));
  const uint32_t L = 64;
  size_t threads = (n + L - 1) / L;
  powers_kernel<P><<<grid_for(threads, 128), 128>>>(buf.as<Fp<P>>(), n, args.as<Fp<P>>(), args.as<Fp<P>>() + 1, L);
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  B200_CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
}

template <class P>
static int domain_build(Domain *d) {
  typedef Fp<P> F;
  const int k = d->k;
  const size_t m = d->m;
  F root, omega, omega_inv, g, g_inv, one, minv, zinv, t;
  load_const<P>(root, &P::root);
  load_const<P>(g, &P::gen);
  load_const<P>(g_inv, &P::gen_inv);
  F::set_one(one);
  omega = root;  // omega = root^(2^(s-k)): get_root_of_unity, field_utils.tcc:40-89
  for (int i = k; i < P::kTwoAdicity; i++) F::sqr(omega, omega);
  F::inv(omega_inv, omega);
  F mint;
  F::set_zero(mint);
  mint.l[0] = (uint32_t)m;
  mint.l[1] = (uint32_t)((uint64_t)m >> 32);
  F::to_mont(mint, mint);
  F::inv(minv, mint);  // sconst = 1/m, basic_radix2_domain.tcc:77
  host_pow_u64<P>(t, g, (uint64_t)m);
  F::sub(t, t, one);
  F::inv(zinv, t);  // Z(g)^-1 = (g^m - 1)^-1, basic_radix2_domain.tcc:111-114,128
  size_t half = m / 2 ? m / 2 : 1;
  B200_CHECK(fill_powers<P>(d->tw_fwd, half, omega, one));
  B200_CHECK(fill_powers<P>(d->tw_inv, half, omega_inv, one));
  B200_CHECK(fill_powers<P>(d->coset, m, g, one));                // g^i
  B200_CHECK(fill_powers<P>(d->icoset_scaled, m, g_inv, minv));   // g^-i / m
  B200_CHECK(d->scratch.alloc(m * sizeof(F)));
  B200_CHECK(d->consts.alloc(2 * sizeof(F)));
  F hc[2] = {minv, zinv};
  B200_CUDA_CHECK(cudaMemcpy(d->consts.p, hc, sizeof(hc), cudaMemcpyHostToDevice));
  B200_CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
}

int domain_create(int curve, size_t m, Domain **out) {
  int k = 0;
  while (((size_t)1 << k) < m) k++;
  int s = curve == 0 ? PrimeA::kTwoAdicity : PrimeB::kTwoAdicity;
  // basic_radix2_domain constructor: m must be a power of two, 1 < m <= 2^s (basic_radix2_domain.tcc:25-60);
  // libfqfft would fall back to other domain families for other sizes - the prover never needs them
  // (d+1 is a power of two, generate_parameters.cpp:35).
  if (m < 2 || ((size_t)1 << k) != m || k > s)
    return set_error(-3, "domain size %zu is not a power of two in [2, 2^%d]", m, s);
  Domain *d = new Domain();
  d->curve = curve;
  d->m = m;
  d->k = k;
  int rc = curve == 0 ? domain_build<PrimeA>(d) : domain_build<PrimeB>(d);
  if (rc) {
    delete d;
    return rc;
  }
  *out = d;
  return 0;
}
void domain_destroy(Domain *d) { delete d; }
int domain_table(const Domain *d, int which, void *h_out, size_t count) {
  const DevBuf *t[6] = {&d->tw_fwd, &d->tw_inv, &d->coset, &d->icoset_scaled, &d->consts, &d->scratch};
  if (which < 0 || which > 5) return set_error(-1, "bad table");
  if (count * 96 > t[which]->bytes) return set_error(-1, "table has fewer entries");
  B200_CUDA_CHECK(cudaMemcpy(h_out, t[which]->p, count * 96, cudaMemcpyDeviceToHost));
  return 0;
}
size_t domain_size(const Domain *d) { return d->m; }

enum { kPlain = 0, kInverse = 1, kCoset = 2, kInverseCoset = 3 };

template <class P>
static int transform(Domain *d, void *d_a, int kind) {
  typedef Fp<P> F;
  const int k = d->k;
  const int npass = (k + 7) / 8;
  const int rbase = k / npass, rem = k % npass;
  const F *tw = (kind == kInverse || kind == kInverseCoset) ? d->tw_inv.as<F>() : d->tw_fwd.as<F>();
  const F *pre = kind == kCoset ? d->coset.as<F>() : nullptr;
  const F *post_tab = kind == kInverseCoset ? d->icoset_scaled.as<F>() : nullptr;
  const F *post_const = kind == kInverse ? d->consts.as<F>() : nullptr;
  F *a = (F *)d_a, *scr = d->scratch.as<F>();
  int s0 = 0;
  for (int p = 0; p < npass; p++) {
    int r = rbase + (p < rem ? 1 : 0);
    const F *src = p == 0 ? a : scr;
    F *dst = (p == npass - 1 && npass > 1) ? a : scr;
    bool last = p == npass - 1;
    size_t smem = (size_t)kLimbs * ((1u << r) + 1) * sizeof(uint32_t);
    unsigned threads = 1u << (r - 1);
    size_t blocks = (size_t)1 << (k - r);
    ntt_pass_kernel<P><<<(unsigned)blocks, threads, smem, g_ntt_stream>>>(src, dst, k, s0, r, tw, p == 0 ? 1 : 0, p == 0 ? pre : nullptr,
                                                            last ? post_tab : nullptr, last ? post_const : nullptr);
    B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
    s0 += r;
  }
  if (npass == 1) B200_CUDA_CHECK(cudaMemcpyAsync(a, scr, d->m * sizeof(F), cudaMemcpyDeviceToDevice, g_ntt_stream));
  return 0;
}

int domain_transform(Domain *d, void *d_a, int kind) {
  return d->curve == 0 ? transform<PrimeA>(d, d_a, kind) : transform<PrimeB>(d, d_a, kind);
}

template <class P>
static int scale_by(Domain *d, void *d_a, int which) {
  fr_scale_kernel<P><<<grid_for(d->m, 128), 128, 0, g_ntt_stream>>>((Fp<P> *)d_a, d->consts.as<Fp<P>>() + which, d->m);
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}
int domain_divide_by_z(Domain *d, void *d_a) {
  return d->curve == 0 ? scale_by<PrimeA>(d, d_a, 1) : scale_by<PrimeB>(d, d_a, 1);
}

int fr_muleq(int curve, void *d_a, const void *d_b, size_t n) {
  if (n == 0) return 0;
  if (curve == 0)
    fr_muleq_kernel<PrimeA><<<grid_for(n, 128), 128, 0, g_ntt_stream>>>((Fp<PrimeA> *)d_a, (const Fp<PrimeA> *)d_b, n);
  else
    fr_muleq_kernel<PrimeB><<<grid_for(n, 128), 128, 0, g_ntt_stream>>>((Fp<PrimeB> *)d_a, (const Fp<PrimeB> *)d_b, n);
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}
int fr_subeq(int curve, void *d_a, const void *d_b, size_t n) {
  if (n == 0) return 0;
  if (curve == 0)
    fr_subeq_kernel<PrimeA><<<grid_for(n, 128), 128, 0, g_ntt_stream>>>((Fp<PrimeA> *)d_a, (const Fp<PrimeA> *)d_b, n);
  else
    fr_subeq_kernel<PrimeB><<<grid_for(n, 128), 128, 0, g_ntt_stream>>>((Fp<PrimeB> *)d_a, (const Fp<PrimeB> *)d_b, n);
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}

// libsnark/main.cpp:104-163 == cuda_prover_piecewise.cu:18-53
int compute_h(Domain *d, void *d_ca, void *d_cb, void *d_cc, void *d_out) {
  const size_t m = d->m;
  B200_CHECK(domain_transform(d, d_ca, kInverse));
  B200_CHECK(domain_transform(d, d_cb, kInverse));
  B200_CHECK(domain_transform(d, d_ca, kCoset));
  B200_CHECK(domain_transform(d, d_cb, kCoset));
  B200_CHECK(fr_muleq(d->curve, d_ca, d_cb, m));
  B200_CHECK(domain_transform(d, d_cc, kInverse));
  B200_CHECK(domain_transform(d, d_cc, kCoset));
  B200_CHECK(fr_subeq(d->curve, d_ca, d_cc, m));
  B200_CHECK(domain_divide_by_z(d, d_ca));
  B200_CHECK(domain_transform(d, d_ca, kInverseCoset));
  B200_CUDA_CHECK(cudaMemcpyAsync(d_out, d_ca, m * 96, cudaMemcpyDeviceToDevice, g_ntt_stream));
  B200_CUDA_CHECK(cudaMemsetAsync((char *)d_out + m * 96, 0, 96, g_ntt_stream));
  return 0;
}

}  // namespace b200
