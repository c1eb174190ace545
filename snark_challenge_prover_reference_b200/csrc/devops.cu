// Element-wise test/bench hooks: they apply, one thread per element, exactly the device primitives the MSM and NTT
// kernels are built from (field.cuh / curve.cuh), so tests/ can compare them with the CPU oracle; plus the synthetic
// base generator used by bench.py and the IMAD-pipe roofline microbenchmark.
#include "common.cuh"
#include "curve.cuh"
#include "devops.h"

namespace b200 {

template <class F>
__global__ void __launch_bounds__(128) field_op_kernel(int op, const F *__restrict__ a, const F *__restrict__ b,
                                                       F *__restrict__ r, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = a[i], y, z;
  if (b) y = b[i];
  else F::set_zero(y);
  switch (op) {
    case 0: F::add(z, x, y); break;
    case 1: F::sub(z, x, y); break;
    case 2: F::mul(z, x, y); break;
    case 3: F::sqr(z, x); break;
    default: F::inv(z, x); break;  // op 4 (tower fields: Fp2::inv / Fp3::inv; the base field goes through fp_conv_kernel)
  }
  r[i] = z;
}

template <class P>
__global__ void __launch_bounds__(128) fp_conv_kernel(int op, const Fp<P> *__restrict__ a, Fp<P> *__restrict__ r, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<P> x = a[i], z;
  if (op == 4) Fp<P>::from_mont(z, x);
  else if (op == 5) Fp<P>::to_mont(z, x);
  else Fp<P>::inv(z, x);
  r[i] = z;
}

// ---- IMAD roofline microbenchmarks -----------------------------------------------------------------------------
// (1) independent mad.wide.u32 chains (8 per thread): the raw IMAD.WIDE issue rate of the chip.
__global__ void __launch_bounds__(256) imad_wide_kernel(unsigned long long *out, uint32_t b, int iters) {
  unsigned long long acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = (unsigned long long)(threadIdx.x + 1) * (k + 3) + blockIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uint32_t lo = (uint32_t)acc[k];
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(lo), "r"(b));
    }
  }
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= acc[k];
  if (s == 0x1234567ull) out[0] = s;
}
// (2) the pattern the Montgomery multiplier uses: two 12-long carry chains of (mad.lo.cc, madc.hi.cc) pairs.
__global__ void __launch_bounds__(256) imad_carry_kernel(uint32_t *out, uint32_t b, int iters) {
  uint32_t x[24], y[24], a[24];
#pragma unroll
  for (int k = 0; k < 24; k++) {
    x[k] = threadIdx.x * 7 + k;
    y[k] = blockIdx.x * 3 + k;
    a[k] = 0x9e3779b9u * (k + 1) + threadIdx.x;
  }
  for (int it = 0; it < iters; it++) {
    uint32_t m = x[0] + b;
    asm volatile(
        "mad.lo.cc.u32 %0, %24, %25, %0;\n\tmadc.hi.cc.u32 %1, %24, %25, %1;\n\t"
        "madc.lo.cc.u32 %2, %24, %26, %2;\n\tmadc.hi.cc.u32 %3, %24, %26, %3;\n\t"
        "madc.lo.cc.u32 %4, %24, %27, %4;\n\tmadc.hi.cc.u32 %5, %24, %27, %5;\n\t"
        "madc.lo.cc.u32 %6, %24, %28, %6;\n\tmadc.hi.cc.u32 %7, %24, %28, %7;\n\t"
        "madc.lo.cc.u32 %8, %24, %29, %8;\n\tmadc.hi.cc.u32 %9, %24, %29, %9;\n\t"
        "madc.lo.cc.u32 %10, %24, %30, %10;\n\tmadc.hi.cc.u32 %11, %24, %30, %11;\n\t"
        "madc.lo.cc.u32 %12, %24, %31, %12;\n\tmadc.hi.cc.u32 %13, %24, %31, %13;\n\t"
        "madc.lo.cc.u32 %14, %24, %32, %14;\n\tmadc.hi.cc.u32 %15, %24, %32, %15;\n\t"
        "madc.lo.cc.u32 %16, %24, %33, %16;\n\tmadc.hi.cc.u32 %17, %24, %33, %17;\n\t"
        "madc.lo.cc.u32 %18, %24, %34, %18;\n\tmadc.hi.cc.u32 %19, %24, %34, %19;\n\t"
        "madc.lo.cc.u32 %20, %24, %35, %20;\n\tmadc.hi.cc.u32 %21, %24, %35, %21;\n\t"
        "madc.lo.cc.u32 %22, %24, %36, %22;\n\tmadc.hi.u32 %23, %24, %36, %23;\n\t"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(x[8]),
          "+r"(x[9]), "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15]), "+r"(x[16]),
          "+r"(x[17]), "+r"(x[18]), "+r"(x[19]), "+r"(x[20]), "+r"(x[21]), "+r"(x[22]), "+r"(x[23])
        : "r"(m), "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[8]), "r"(a[10]), "r"(a[12]), "r"(a[14]),
          "r"(a[16]), "r"(a[18]), "r"(a[20]), "r"(a[22]));
    asm volatile(
        "mad.lo.cc.u32 %0, %24, %25, %0;\n\tmadc.hi.cc.u32 %1, %24, %25, %1;\n\t"
        "madc.lo.cc.u32 %2, %24, %26, %2;\n\tmadc.hi.cc.u32 %3, %24, %26, %3;\n\t"
        "madc.lo.cc.u32 %4, %24, %27, %4;\n\tmadc.hi.cc.u32 %5, %24, %27, %5;\n\t"
        "madc.lo.cc.u32 %6, %24, %28, %6;\n\tmadc.hi.cc.u32 %7, %24, %28, %7;\n\t"
        "madc.lo.cc.u32 %8, %24, %29, %8;\n\tmadc.hi.cc.u32 %9, %24, %29, %9;\n\t"
        "madc.lo.cc.u32 %10, %24, %30, %10;\n\tmadc.hi.cc.u32 %11, %24, %30, %11;\n\t"
        "madc.lo.cc.u32 %12, %24, %31, %12;\n\tmadc.hi.cc.u32 %13, %24, %31, %13;\n\t"
        "madc.lo.cc.u32 %14, %24, %32, %14;\n\tmadc.hi.cc.u32 %15, %24, %32, %15;\n\t"
        "madc.lo.cc.u32 %16, %24, %33, %16;\n\tmadc.hi.cc.u32 %17, %24, %33, %17;\n\t"
        "madc.lo.cc.u32 %18, %24, %34, %18;\n\tmadc.hi.cc.u32 %19, %24, %34, %19;\n\t"
        "madc.lo.cc.u32 %20, %24, %35, %20;\n\tmadc.hi.cc.u32 %21, %24, %35, %21;\n\t"
        "madc.lo.cc.u32 %22, %24, %36, %22;\n\tmadc.hi.u32 %23, %24, %36, %23;\n\t"
        : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7]), "+r"(y[8]),
          "+r"(y[9]), "+r"(y[10]), "+r"(y[11]), "+r"(y[12]), "+r"(y[13]), "+r"(y[14]), "+r"(y[15]), "+r"(y[16]),
          "+r"(y[17]), "+r"(y[18]), "+r"(y[19]), "+r"(y[20]), "+r"(y[21]), "+r"(y[22]), "+r"(y[23])
        : "r"(m), "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(a[9]), "r"(a[11]), "r"(a[13]), "r"(a[15]),
          "r"(a[17]), "r"(a[19]), "r"(a[21]), "r"(a[23]));
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 24; k++) s ^= x[k] ^ y[k];
  if (s == 0x1234567u) out[0] = s;
}

// (3) the production instruction sequence itself: the generated Montgomery multiply (fp_ptx_gen.cuh: 1152 wide MACs in
// ~1400 instructions) on register-resident operands, two dependent multiplications per iteration, at the occupancy of
// the accumulation kernels (16 warps per SM). No memory traffic: what the multiplier can do when nothing else stalls.
__global__ void __launch_bounds__(128, 4) imad_mont_kernel(uint32_t *out, int iters) {
  uint32_t x[kLimbs], y[kLimbs], z[kLimbs];
#pragma unroll
  for (int k = 0; k < kLimbs; k++) {
    x[k] = threadIdx.x * 7 + k;
    y[k] = blockIdx.x * 3 + k + 1;
  }
  x[kLimbs - 1] &= 0xffu;
  y[kLimbs - 1] &= 0xffu;
  for (int it = 0; it < iters; it++) {
    fp_mul_ptx_B(z, x, y);
    fp_mul_ptx_B(x, z, y);
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kLimbs; k++) s ^= x[k];
  if (s == 0x1234567u) out[0] = s;
}

int imad_peak(double *mac32_per_s2, double *ms2) {
  DevBuf buf;
  B200_CHECK(buf.alloc(64));
  cudaDeviceProp prop;
  int dev = 0;
  B200_CUDA_CHECK(cudaGetDevice(&dev));
  B200_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  Timer tm;
  // nominal: 32 IMAD.WIDE per clock per SM (half the 32-bit IMAD rate) at the maximum SM clock
  int khz = 0;
  B200_CUDA_CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  mac32_per_s2[2] = 32.0 * prop.multiProcessorCount * (double)khz * 1e3;
  ms2[2] = khz / 1e3;
  for (int variant = 0; variant < 3; variant++) {
    // >= 50 ms per run: a 3 ms run ends before the SM clock has ramped up and under-reads the pipe by ~10 %
    int iters = variant == 0 ? (1 << 18) : (variant == 1 ? (1 << 16) : 3000);
    const int mont_blocks = prop.multiProcessorCount * 4, mont_threads = 128;
    double macs = variant == 0 ? (double)blocks * threads * 8.0 * iters
                  : variant == 1 ? (double)blocks * threads * 24.0 * iters
                                 : (double)mont_blocks * mont_threads * 2.0 * 1152.0 * iters;
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      tm.start();
      if (variant == 0)
        imad_wide_kernel<<<blocks, threads>>>(buf.as<unsigned long long>(), 0x7fffffffu, iters);
      else if (variant == 1)
        imad_carry_kernel<<<blocks, threads>>>(buf.as<uint32_t>(), 0x7fffffffu, iters);
      else
        imad_mont_kernel<<<mont_blocks, mont_threads>>>(buf.as<uint32_t>(), iters);
      float ms = tm.stop();
      if (rep > 0 && ms < best) best = ms;
    }
    B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
    const int slot = variant < 2 ? variant : 3;  // [2] holds the nominal figure
    mac32_per_s2[slot] = macs / (best * 1e-3);
    ms2[slot] = best;
  }
  return 0;
}

// ---- dispatch ------------------------------------------------------------------------------------------------------
template <class P>
static int fp_op_t(int op, const void *a, const void *b, void *r, size_t n) {
  if (op <= 3)
    field_op_kernel<Fp<P>><<<grid_for(n, 128), 128>>>(op, (const Fp<P> *)a, (const Fp<P> *)b, (Fp<P> *)r, n);
  else
    fp_conv_kernel<P><<<grid_for(n, 128), 128>>>(op, (const Fp<P> *)a, (Fp<P> *)r, n);
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}
int dev_fp_op(int tag, int op, const void *a, const void *b, void *r, size_t n) {
  if (n == 0) return 0;
  if (op < 0 || op > 6) return set_error(-1, "dev_fp_op: bad op %d", op);
  return tag == 0 ? fp_op_t<PrimeA>(op, a, b, r, n) : fp_op_t<PrimeB>(op, a, b, r, n);
}
int dev_fqe_op(int curve, int op, const void *a, const void *b, void *r, size_t n) {
  if (n == 0) return 0;
  if (op < 0 || op > 4) return set_error(-1, "dev_fqe_op: bad op %d", op);
  if (curve == 0) {
    typedef Mnt4G2::F F;
    field_op_kernel<F><<<grid_for(n, 128), 128>>>(op, (const F *)a, (const F *)b, (F *)r, n);
  } else {
    typedef Mnt6G2::F F;
    field_op_kernel<F><<<grid_for(n, 128), 128>>>(op, (const F *)a, (const F *)b, (F *)r, n);
  }
  B200_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return 0;
}
int group_op_mnt4g1(int, const void *, const void *, void *, size_t);
int group_op_mnt4g2(int, const void *, const void *, void *, size_t);
int group_op_mnt6g1(int, const void *, const void *, void *, size_t);
int group_op_mnt6g2(int, const void *, const void *, void *, size_t);
int gen_points_mnt4g1(void *, size_t, uint64_t);
int gen_points_mnt4g2(void *, size_t, uint64_t);
int gen_points_mnt6g1(void *, size_t, uint64_t);
int gen_points_mnt6g2(void *, size_t, uint64_t);
int dev_group_op(int curve, int group, int op, const void *p, const void *q, void *r, size_t n) {
  if (n == 0) return 0;
  if (op < 0 || op > 3) return set_error(-1, "dev_group_op: bad op %d", op);
  if (curve == 0 && group == 1) return group_op_mnt4g1(op, p, q, r, n);
  if (curve == 0 && group == 2) return group_op_mnt4g2(op, p, q, r, n);
  if (curve == 1 && group == 1) return group_op_mnt6g1(op, p, q, r, n);
  if (curve == 1 && group == 2) return group_op_mnt6g2(op, p, q, r, n);
  return set_error(-1, "dev_group_op: bad curve/group");
}
int batch_exp_mnt4g1(const void *, const void *, size_t, void *, int, double *);
int batch_exp_mnt4g2(const void *, const void *, size_t, void *, int, double *);
int batch_exp_mnt6g1(const void *, const void *, size_t, void *, int, double *);
int batch_exp_mnt6g2(const void *, const void *, size_t, void *, int, double *);
int batch_exp(int curve, int group, const void *h_base, const void *d_scalars, size_t n, void *d_out, int window, double *ms3) {
  if (curve == 0 && group == 1) return batch_exp_mnt4g1(h_base, d_scalars, n, d_out, window, ms3);
  if (curve == 0 && group == 2) return batch_exp_mnt4g2(h_base, d_scalars, n, d_out, window, ms3);
  if (curve == 1 && group == 1) return batch_exp_mnt6g1(h_base, d_scalars, n, d_out, window, ms3);
  if (curve == 1 && group == 2) return batch_exp_mnt6g2(h_base, d_scalars, n, d_out, window, ms3);
  return set_error(-1, "batch_exp: bad curve/group");
}
int gen_points(int curve, int group, void *out, size_t n, uint64_t first) {
  if (n == 0) return 0;
  if (curve == 0 && group == 1) return gen_points_mnt4g1(out, n, first);
  if (curve == 0 && group == 2) return gen_points_mnt4g2(out, n, first);
  if (curve == 1 && group == 1) return gen_points_mnt6g1(out, n, first);
  if (curve == 1 && group == 2) return gen_points_mnt6g2(out, n, first);
  return set_error(-1, "gen_points: bad curve/group");
}

}  // namespace b200
