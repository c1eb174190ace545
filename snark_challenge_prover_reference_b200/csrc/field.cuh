// 753-bit prime-field arithmetic (Fq / Fr of MNT4753 and MNT6753) and the Fq2 / Fq3 towers used by G2.
//
// One field element = 24 little-endian u32 limbs = 96 bytes, Montgomery form x*2^768 mod p, always fully reduced,
// i.e. bit-compatible with the reference's on-disk / in-memory encoding (libsnark/serialization.hpp:22-32,
// depends/libff/libff/algebra/fields/fp.tcc:161-186).
//
// The same formulas compile for host and device:
//   * device: one thread owns one element; the multiply is the generated IMAD.WIDE chain in fp_ptx_gen.cuh, kept
//     out-of-line (one copy of the ~1.3k-instruction body per modulus) and fed from memory-resident operands.
//   * host  : portable 12 x u64 CIOS with unsigned __int128 - used for the O(1) serial tails of the prover
//     (window combine, r*Bt1, to-affine inversion) and by the CPU-side unit tests of the shared formulas.
#pragma once
#include <stdint.h>
#include <string.h>
#if defined(__x86_64__) && !defined(__CUDA_ARCH__)
#include <immintrin.h>
#endif
#include "constants_gen.h"

#if defined(__CUDACC__)
#pragma nv_diag_suppress 1675  // "#pragma GCC unroll" in the host-only arithmetic is meant for the host compiler
#include "fp_ptx_gen.cuh"
#define B200_DEV __device__
#define B200_INLINE __forceinline__
// Tower-field operations are kept out of line on the device: a G2 mixed addition inlines to >300 KB of straight-line
// add/sub chains otherwise, which thrashes the instruction cache (ncu: no_instruction was the top stall) and gives
// every inlined temporary its own stack slot (6.5 KB / thread of local memory spilling to DRAM).
#define B200_NOINLINE __noinline__
#else
#define B200_DEV
#define B200_INLINE inline
#define B200_NOINLINE
#endif

namespace b200 {

constexpr int kLimbs = 24;

// ------------------------------------------------------------------------------------------------ base field
template <class P>
struct alignas(16) Fp {
  uint32_t l[kLimbs];
  typedef P Prime;
  static constexpr int kDegree = 1;

  B200_HD static void set_zero(Fp &r) {
#pragma unroll
    for (int i = 0; i < kLimbs; i++) r.l[i] = 0;
  }
  B200_HD static void set_one(Fp &r) {
#pragma unroll
    for (int i = 0; i < kLimbs; i++) r.l[i] = P::one(i);
  }
  B200_HD static bool is_zero(const Fp &a) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kLimbs; i++) acc |= a.l[i];
    return acc == 0;
  }
  B200_HD static bool eq(const Fp &a, const Fp &b) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kLimbs; i++) acc |= (a.l[i] ^ b.l[i]);
    return acc == 0;
  }

  // ---------------------------------------------------------------- host implementations (u64 limbs)
#if !defined(__CUDA_ARCH__)
  static inline uint64_t p64(int i) { return (uint64_t)P::p(2 * i) | ((uint64_t)P::p(2 * i + 1) << 32); }
  static inline uint64_t ld64(const Fp &a, int i) { return (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32); }
  static inline void st64(Fp &r, int i, uint64_t v) {
    r.l[2 * i] = (uint32_t)v;
    r.l[2 * i + 1] = (uint32_t)(v >> 32);
  }
  static inline void host_cond_sub_p(uint64_t *t) {  // t in [0, 2p) -> [0, p)
    uint64_t u[12];
    unsigned __int128 brw = 0;
#pragma GCC unroll 12
    for (int i = 0; i < 12; i++) {
      unsigned __int128 d = (unsigned __int128)t[i] - p64(i) - (uint64_t)brw;
      u[i] = (uint64_t)d;
      brw = (d >> 64) & 1;
    }
    if (!brw)
      for (int i = 0; i < 12; i++) t[i] = u[i];
  }
  static inline void host_add(Fp &r, const Fp &a, const Fp &b) {
    uint64_t t[12];
    unsigned __int128 c = 0;
#pragma GCC unroll 12
    for (int i = 0; i < 12; i++) {
      c += (unsigned __int128)ld64(a, i) + ld64(b, i);
      t[i] = (uint64_t)c;
      c >>= 64;
    }
    host_cond_sub_p(t);
    for (int i = 0; i < 12; i++) st64(r, i, t[i]);
  }
  static inline void host_sub(Fp &r, const Fp &a, const Fp &b) {
    uint64_t t[12];
    unsigned __int128 brw = 0;
#pragma GCC unroll 12
    for (int i = 0; i < 12; i++) {
      unsigned __int128 d = (unsigned __int128)ld64(a, i) - ld64(b, i) - (uint64_t)brw;
      t[i] = (uint64_t)d;
      brw = (d >> 64) & 1;
    }
    if (brw) {
      unsigned __int128 c = 0;
  #pragma GCC unroll 12
    for (int i = 0; i < 12; i++) {
        c += (unsigned __int128)t[i] + p64(i);
        t[i] = (uint64_t)c;
        c >>= 64;
      }
    }
    for (int i = 0; i < 12; i++) st64(r, i, t[i]);
  }
#if defined(__x86_64__) && defined(__BMI2__) && defined(__ADX__)
  // x86-64 fast path: t[0..13] += a[0..11] * b with one mulx per limb; the product halves are chained through CF (adcx:
  // high half of limb j-1 into the low half of limb j), the accumulator through OF (adox) - two independent carry chains,
  // fully unrolled. (The intrinsics version of this row was compiled by gcc into loops over stack arrays: 830 ns per
  // multiplication against ~140 ns for this one; the serial host tail r * Bt1 is ~11 K multiplications.)
  static inline __attribute__((always_inline)) void host_mac_row(unsigned long long *t, const unsigned long long *a,
                                                                 unsigned long long b) {
    unsigned long long lo, h0, h1;
#define B200_MAC_STEP(OFF, HPREV, HNEXT)                                                                               \
  "mulx " #OFF "(%[a]), %[lo], %[" #HNEXT "]\n\t"                                                                      \
  "adcx %[" #HPREV "], %[lo]\n\t"                                                                                      \
  "adox " #OFF "(%[t]), %[lo]\n\t"                                                                                     \
  "movq %[lo], " #OFF "(%[t])\n\t"
    asm volatile(
        "xorl %%eax, %%eax\n\t"  // rax = 0, CF = OF = 0
        "mulx 0(%[a]), %[lo], %[h0]\n\t"
        "adox 0(%[t]), %[lo]\n\t"
        "movq %[lo], 0(%[t])\n\t"
        B200_MAC_STEP(8, h0, h1) B200_MAC_STEP(16, h1, h0) B200_MAC_STEP(24, h0, h1) B200_MAC_STEP(32, h1, h0)
        B200_MAC_STEP(40, h0, h1) B200_MAC_STEP(48, h1, h0) B200_MAC_STEP(56, h0, h1) B200_MAC_STEP(64, h1, h0)
        B200_MAC_STEP(72, h0, h1) B200_MAC_STEP(80, h1, h0) B200_MAC_STEP(88, h0, h1)
        "adcx %%rax, %[h1]\n\t"      // last high half + CF (cannot overflow: the high half of a product is <= 2^64 - 2)
        "adox 96(%[t]), %[h1]\n\t"   // + t[12] + OF
        "movq %[h1], 96(%[t])\n\t"
        "adox 104(%[t]), %%rax\n\t"  // t[13] += OF
        "movq %%rax, 104(%[t])\n\t"
        : [lo] "=&r"(lo), [h0] "=&r"(h0), [h1] "=&r"(h1)
        : [t] "r"(t), [a] "r"(a), "d"(b)
        : "rax", "cc", "memory");
#undef B200_MAC_STEP
  }
  struct HostConsts {
    unsigned long long p[12], inv;
    HostConsts() {
      for (int i = 0; i < 12; i++) p[i] = p64(i);
      unsigned long long x = 1;  // p[0]^-1 mod 2^64 by Newton steps
      for (int i = 0; i < 6; i++) x *= 2 - p[0] * x;
      inv = 0 - x;
    }
  };
  static inline void host_mul(Fp &r, const Fp &a, const Fp &b) {
    static const HostConsts K;
    // row i works on the 14-word window t[i .. i+13]: no shifting between rows; the result is t[12 .. 23]
    unsigned long long t[26], av[12], bv[12];
    memcpy(av, a.l, 96);
    memcpy(bv, b.l, 96);
    for (int i = 0; i < 26; i++) t[i] = 0;
#pragma GCC unroll 12
    for (int i = 0; i < 12; i++) {
      host_mac_row(t + i, av, bv[i]);
      host_mac_row(t + i, K.p, t[i] * K.inv);
    }
    uint64_t tt[12];
    for (int i = 0; i < 12; i++) tt[i] = t[12 + i];
    host_cond_sub_p(tt);
    memcpy(r.l, tt, 96);
  }
#else
  static inline void host_mul(Fp &r, const Fp &a, const Fp &b) {
    // CIOS Montgomery, 12 x u64, inv64 derived from the 32-bit constant by one Newton step.
    static const uint64_t inv64 = []() {
      uint64_t p0 = p64(0), x = 1;  // x = p0^-1 mod 2^64
      for (int i = 0; i < 6; i++) x *= 2 - p0 * x;
      return (uint64_t)(0 - x);
    }();
    uint64_t t[14];
    for (int i = 0; i < 14; i++) t[i] = 0;
    uint64_t av[12], bv[12], pv[12];
    for (int i = 0; i < 12; i++) {
      av[i] = ld64(a, i);
      bv[i] = ld64(b, i);
      pv[i] = p64(i);
    }
    for (int i = 0; i < 12; i++) {
      unsigned __int128 c = 0;
      for (int j = 0; j < 12; j++) {
        c += (unsigned __int128)av[j] * bv[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[12];
      t[12] = (uint64_t)c;
      t[13] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * inv64;
      c = (unsigned __int128)m * pv[0] + t[0];
      c >>= 64;
      for (int j = 1; j < 12; j++) {
        c += (unsigned __int128)m * pv[j] + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[12];
      t[11] = (uint64_t)c;
      t[12] = t[13] + (uint64_t)(c >> 64);
    }
    host_cond_sub_p(t);
    for (int i = 0; i < 12; i++) st64(r, i, t[i]);
  }
#endif
#endif

  // ---------------------------------------------------------------- dispatch
#if defined(__CUDACC__)
  // out-of-line device add / sub (operands by pointer, like the multiply): callers then hold no limbs in registers
  // across field operations, which keeps the big point-arithmetic kernels at ~128 registers (16 warps/SM).
  // every Fp object is 16-byte aligned (alignas(16), 96-byte stride in arrays): move operands as 6 x 128 bits
  static __device__ __forceinline__ void load_limbs(uint32_t (&x)[kLimbs], const uint32_t *a) {
    const uint4 *a4 = reinterpret_cast<const uint4 *>(a);
#pragma unroll
    for (int k = 0; k < kLimbs / 4; k++) {
      const uint4 v = a4[k];
      x[4 * k] = v.x;
      x[4 * k + 1] = v.y;
      x[4 * k + 2] = v.z;
      x[4 * k + 3] = v.w;
    }
  }
  static __device__ __forceinline__ void store_limbs(uint32_t *r, const uint32_t (&z)[kLimbs]) {
    uint4 *r4 = reinterpret_cast<uint4 *>(r);
#pragma unroll
    for (int k = 0; k < kLimbs / 4; k++) r4[k] = make_uint4(z[4 * k], z[4 * k + 1], z[4 * k + 2], z[4 * k + 3]);
  }
  // streaming variant: the limbs come from GLOBAL memory with the evict-first hint (data that passes through once)
  static __device__ __forceinline__ void load_limbs_cs(uint32_t (&x)[kLimbs], const uint32_t *a) {
    const uint4 *a4 = reinterpret_cast<const uint4 *>(a);
#pragma unroll
    for (int k = 0; k < kLimbs / 4; k++) {
      const uint4 v = __ldcs(a4 + k);
      x[4 * k] = v.x;
      x[4 * k + 1] = v.y;
      x[4 * k + 2] = v.z;
      x[4 * k + 3] = v.w;
    }
  }
  // r = a +/- b with operands read in place; GA / GB: that operand lives in global memory and is streamed
  template <bool SUB, bool GA, bool GB>
  static __device__ __noinline__ void addsub_dev_g(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t x[kLimbs], y[kLimbs], z[kLimbs];
    if (GA) load_limbs_cs(x, a);
    else load_limbs(x, a);
    if (GB) load_limbs_cs(y, b);
    else load_limbs(y, b);
    if (SUB) {
      if (P::kTag == 'A') fp_sub_ptx_A(z, x, y);
      else fp_sub_ptx_B(z, x, y);
    } else {
      if (P::kTag == 'A') fp_add_ptx_A(z, x, y);
      else fp_add_ptx_B(z, x, y);
    }
    store_limbs(r, z);
  }
  template <bool GA, bool GB>
  static __device__ __forceinline__ void sub_g(Fp &r, const Fp &a, const Fp &b) { addsub_dev_g<true, GA, GB>(r.l, a.l, b.l); }
  template <bool GA, bool GB>
  static __device__ __forceinline__ void add_g(Fp &r, const Fp &a, const Fp &b) { addsub_dev_g<false, GA, GB>(r.l, a.l, b.l); }
  static __device__ __noinline__ void add_dev(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t x[kLimbs], y[kLimbs], z[kLimbs];
    load_limbs(x, a);
    load_limbs(y, b);
    if (P::kTag == 'A')
      fp_add_ptx_A(z, x, y);
    else
      fp_add_ptx_B(z, x, y);
    store_limbs(r, z);
  }
  static __device__ __noinline__ void sub_dev(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t x[kLimbs], y[kLimbs], z[kLimbs];
    load_limbs(x, a);
    load_limbs(y, b);
    if (P::kTag == 'A')
      fp_sub_ptx_A(z, x, y);
    else
      fp_sub_ptx_B(z, x, y);
    store_limbs(r, z);
  }
#endif
  B200_HD static B200_INLINE void add(Fp &r, const Fp &a, const Fp &b) {
#if defined(__CUDA_ARCH__)
#if defined(B200_FP_ADD_INLINE)
    if (P::kTag == 'A')
      fp_add_ptx_A(r.l, a.l, b.l);
    else
      fp_add_ptx_B(r.l, a.l, b.l);
#else
    add_dev(r.l, a.l, b.l);
#endif
#else
    host_add(r, a, b);
#endif
  }
  B200_HD static B200_INLINE void sub(Fp &r, const Fp &a, const Fp &b) {
#if defined(__CUDA_ARCH__)
#if defined(B200_FP_ADD_INLINE)
    if (P::kTag == 'A')
      fp_sub_ptx_A(r.l, a.l, b.l);
    else
      fp_sub_ptx_B(r.l, a.l, b.l);
#else
    sub_dev(r.l, a.l, b.l);
#endif
#else
    host_sub(r, a, b);
#endif
  }
  B200_HD static B200_INLINE void dbl(Fp &r, const Fp &a) { add(r, a, a); }
  B200_HD static B200_INLINE void neg(Fp &r, const Fp &a) {
    Fp z;
    set_zero(z);
    sub(r, z, a);
  }
#if defined(__CUDACC__)
  // out-of-line device multiply: operands come from (local/shared/global) memory, limbs live in registers only
  // inside the body. One copy per modulus per module keeps the instruction footprint inside the 32 KB L1.5 I-cache.
  // b_streamed != 0: b lives in global memory and is read with the evict-first hint (a run-time flag, not a second
  // copy of the body)
  static __device__ __noinline__ void mul_dev(uint32_t *r, const uint32_t *a, const uint32_t *b, int b_streamed = 0) {
    uint32_t x[kLimbs], y[kLimbs], z[kLimbs];
    load_limbs(x, a);
    if (b_streamed) load_limbs_cs(y, b);
    else load_limbs(y, b);
    // Interleaved CIOS, 1152 IMAD.WIDE. (A one-level Karatsuba product with half-width reduction was generated and
    // measured in round 1 - 12 % fewer multiplier instructions, but longer carry chains and +16 registers drop the
    // fmaheavy pipe from 94 % to 84 % busy - and removed in round 2; tools/gen_fp_ptx.py --experimental still emits it.)
    if (P::kTag == 'A')
      fp_mul_ptx_A(z, x, y);
    else
      fp_mul_ptx_B(z, x, y);
    store_limbs(r, z);
  }
#endif
  B200_HD static B200_INLINE void mul(Fp &r, const Fp &a, const Fp &b) {
#if defined(__CUDA_ARCH__)
    mul_dev(r.l, a.l, b.l);
#else
    host_mul(r, a, b);
#endif
  }
#if defined(__CUDACC__)
  // r = a * b with b streamed from global memory
  static __device__ __forceinline__ void mul_bg(Fp &r, const Fp &a, const Fp &b) { mul_dev(r.l, a.l, b.l, 1); }
#endif
  // sqr: a multiplication. A dedicated 876-MAC squaring exists (sqr_fast) but does not pay in the bucket accumulation
  // (round 2: G1 49.5 -> 49.0 ms, G2 160.7 -> 162.0 ms - squarings are 2 of 10 multiplications there and a second
  // 22 KB body competes for the instruction cache); the base-table builder, whose Jacobian doublings are 8 squarings +
  // 1 multiplication, calls sqr_fast.
  B200_HD static B200_INLINE void sqr(Fp &r, const Fp &a) { mul(r, a, a); }
#if defined(__CUDACC__)
  static __device__ __noinline__ void sqr_dev(uint32_t *r, const uint32_t *a) {
    uint32_t x[kLimbs], z[kLimbs];
    load_limbs(x, a);
    if (P::kTag == 'A')
      fp_sqr_ptx_A(z, x);
    else
      fp_sqr_ptx_B(z, x);
    store_limbs(r, z);
  }
#endif
  B200_HD static B200_INLINE void sqr_fast(Fp &r, const Fp &a) {
#if defined(__CUDA_ARCH__)
    sqr_dev(r.l, a.l);
#else
    host_mul(r, a, a);
#endif
  }

  // out-of-line add / sub / small-constant multiply for the tower fields (keeps Fq2/Fq3 code a short list of calls:
  // with the ~100-instruction carry chains inlined ~40 times the G2 kernels no longer fit the instruction cache)
  B200_HD static B200_NOINLINE void add_ni(Fp &r, const Fp &a, const Fp &b) { add(r, a, b); }
  B200_HD static B200_NOINLINE void sub_ni(Fp &r, const Fp &a, const Fp &b) { sub(r, a, b); }
  B200_HD static B200_NOINLINE void neg_ni(Fp &r, const Fp &a) { neg(r, a); }
  template <unsigned K>
  B200_HD static B200_NOINLINE void mul_small_ni(Fp &r, const Fp &a) { mul_small<K>(r, a); }

  // r = k*a for a small compile-time constant k (double-and-add; used for curve a, non-residues 13 / 11)
  template <unsigned K>
  B200_HD static B200_INLINE void mul_small(Fp &r, const Fp &a) {
    static_assert(K >= 1 && K < 256, "small constant");
    Fp acc = a, base = a;
    int top = 7;
    while (!((K >> top) & 1)) top--;
    for (int bit = top - 1; bit >= 0; bit--) {
      dbl(acc, acc);
      if ((K >> bit) & 1) add(acc, acc, base);
    }
    r = acc;
  }

  // Montgomery <-> integer
  B200_HD static void from_mont(Fp &r, const Fp &a) {
    Fp one_int;
    set_zero(one_int);
    one_int.l[0] = 1;
    mul(r, a, one_int);
  }
  B200_HD static void to_mont(Fp &r, const Fp &a) {
    Fp r2;
#pragma unroll
    for (int i = 0; i < kLimbs; i++) r2.l[i] = P::r2(i);
    mul(r, a, r2);
  }
  // r = a^e, e given as little-endian u32 words (plain integer)
  B200_HD static void pow_words(Fp &r, const Fp &a, const uint32_t *e, int nwords) {
    Fp acc;
    set_one(acc);
    bool started = false;
    for (int w = nwords - 1; w >= 0; w--) {
      for (int bit = 31; bit >= 0; bit--) {
        if (started) sqr(acc, acc);
        if ((e[w] >> bit) & 1) {
          if (started)
            mul(acc, acc, a);
          else {
            acc = a;
            started = true;
          }
        }
      }
    }
    r = acc;
  }
  // r = a^-1 by a right-shift binary extended gcd that needs only add / sub / shift / select on 24 limbs - on the GPU it
  // runs on the ALU pipe, which the multiplier-bound kernels leave ~90 % idle, instead of ~1130 multiplications on the
  // IMAD pipe. Branch-free per iteration (all lanes of a warp stay converged); at most 2*753 iterations.
  //   u = x, v = p, A = 1, C = 0 with A*x = u, C*x = v (mod p); while u != 0:
  //     if u odd: (if u < v swap (u,A) <-> (v,C));  u -= v;  A = A - C mod p;     u >>= 1;  A = A/2 mod p
  //   => v = 1 and C = x^-1 (plain integer inverse of the Montgomery residue x = a*R, i.e. a^-1 * R^-1);
  //   one Montgomery multiplication by R^3 brings it back to a^-1 * R (the reference does the same, fp.tcc:677-683).
  B200_HD static void inv_binary(Fp &r, const Fp &a) {
    uint32_t u[kLimbs], v[kLimbs], A[kLimbs], C[kLimbs];
#pragma unroll
    for (int i = 0; i < kLimbs; i++) {
      u[i] = a.l[i];
      v[i] = P::p(i);
      A[i] = 0;
      C[i] = 0;
    }
    A[0] = 1;
    for (int iter = 0; iter < 2 * 753; iter++) {
      uint32_t nz = 0;
#pragma unroll
      for (int i = 0; i < kLimbs; i++) nz |= u[i];
      if (nz == 0) break;  // per-thread exit: lanes that finish early simply idle until the warp's slowest lane is done
      const uint32_t odd = 0u - (u[0] & 1u);  // all-ones when u is odd
      // d = u - v (borrow -> lt), e = A - C (borrow -> bl)
      uint32_t d[kLimbs], e[kLimbs];
      uint32_t brw = 0, brw2 = 0;
#pragma unroll
      for (int i = 0; i < kLimbs; i++) {
        uint64_t t = (uint64_t)u[i] - v[i] - brw;
        d[i] = (uint32_t)t;
        brw = (uint32_t)(t >> 32) & 1u;
        uint64_t t2 = (uint64_t)A[i] - C[i] - brw2;
        e[i] = (uint32_t)t2;
        brw2 = (uint32_t)(t2 >> 32) & 1u;
      }
      const uint32_t lt = 0u - brw;          // u < v
      const uint32_t swap = odd & lt;
      // new v = swap ? u : v ; new C = swap ? A : C
      // new u = odd ? (lt ? -d : d) : u ; new A = odd ? (lt ? (C - A mod p) : (A - C mod p)) : A
      //   A - C mod p = e + (bl ? p : 0) ;  C - A mod p = -e + (bl ? 0 : p)
      const uint32_t bl = 0u - brw2;
      const uint32_t addp = lt ? ~bl : bl;   // all-ones when p must be added
      uint32_t cn = lt & 1u, ce = lt & 1u;   // +1 of the two's-complement negation
      uint32_t cp = 0;
#pragma unroll
      for (int i = 0; i < kLimbs; i++) {
        uint64_t nd = (uint64_t)(d[i] ^ lt) + cn;  // lt ? -d : d
        cn = (uint32_t)(nd >> 32);
        uint64_t ne = (uint64_t)(e[i] ^ lt) + ce;  // lt ? -e : e
        ce = (uint32_t)(ne >> 32);
        uint64_t na = (uint64_t)(uint32_t)ne + (P::p(i) & addp) + cp;
        cp = (uint32_t)(na >> 32);
        const uint32_t ui = u[i], ai = A[i];
        u[i] = (ui & ~odd) | ((uint32_t)nd & odd);
        A[i] = (ai & ~odd) | ((uint32_t)na & odd);
        v[i] = (v[i] & ~swap) | (ui & swap);
        C[i] = (C[i] & ~swap) | (ai & swap);
      }
      // u >>= 1 ; A = (A + (A odd ? p : 0)) >> 1   (A + p < 2^754 fits)
      const uint32_t aodd = 0u - (A[0] & 1u);
      uint32_t c2 = 0;
      uint32_t t[kLimbs];
#pragma unroll
      for (int i = 0; i < kLimbs; i++) {
        uint64_t s2 = (uint64_t)A[i] + (P::p(i) & aodd) + c2;
        t[i] = (uint32_t)s2;
        c2 = (uint32_t)(s2 >> 32);
      }
#pragma unroll
      for (int i = 0; i < kLimbs; i++) {
        const uint32_t un = i + 1 < kLimbs ? u[i + 1] : 0u;
        const uint32_t tn = i + 1 < kLimbs ? t[i + 1] : c2;
        u[i] = (u[i] >> 1) | (un << 31);
        A[i] = (t[i] >> 1) | (tn << 31);
      }
    }
    // C may equal p (== 0 mod p) only if the inverse were 0: impossible for a != 0. Bring back to Montgomery form.
    Fp ci, r3;
#pragma unroll
    for (int i = 0; i < kLimbs; i++) {
      ci.l[i] = C[i];
      r3.l[i] = P::r3(i);
    }
    mul(r, ci, r3);
  }
  // r = a^-1 by the batched binary gcd (Pornin, "Optimized Binary GCD for Modular Inversion", 2020): the 2*753
  // single-bit steps of inv_binary are grouped 31 at a time; each group runs on 64-bit approximations of (a, b) (low 31
  // bits exact, top 33 bits) and yields factors |f|,|g| <= 2^31 that are then applied ONCE to the full-width values:
  //   (a, b) <- ((f0 a + g0 b) / 2^31, (f1 a + g1 b) / 2^31),  (u, v) <- the same combination mod p (the division by
  //   2^31 made exact by adding the right multiple of p).
  // 49 rounds; ~25 K instructions instead of ~770 K. Host and device code (no divergence inside a round).
  // Round-2 building block for the batch-affine accumulation; the host uses it for to_affine.
  B200_HD static void inv_bingcd(Fp &r, const Fp &x) {
    constexpr int kW = kLimbs + 1;  // 25 limbs: products by a 32-bit factor, and p << 31
    uint32_t a[kLimbs], b[kLimbs], u[kLimbs], v[kLimbs];
    for (int i = 0; i < kLimbs; i++) {
      a[i] = x.l[i];
      b[i] = P::p(i);
      u[i] = 0;
      v[i] = 0;
    }
    u[0] = 1;
    // |f| * s (25 limbs)
    auto mul_small = [](uint32_t (&out)[kW], const uint32_t (&s)[kLimbs], uint32_t f) {
      uint64_t c = 0;
      for (int i = 0; i < kLimbs; i++) {
        c += (uint64_t)s[i] * f;
        out[i] = (uint32_t)c;
        c >>= 32;
      }
      out[kLimbs] = (uint32_t)c;
    };
    for (int round = 0; round < 49; round++) {
      // ---- 64-bit approximations: n = bit length of max(a, b)
      int top = kLimbs - 1;
      while (top > 1 && (a[top] | b[top]) == 0) top--;
      uint64_t abar, bbar;
      {
        const uint32_t hi = a[top] | b[top];
        int lz = 0;
        for (uint32_t t = hi; lz < 32 && !(t & 0x80000000u); t <<= 1) lz++;
        const int n = 32 * (top + 1) - lz;
        if (n <= 64) {
          abar = (uint64_t)a[0] | ((uint64_t)a[1] << 32);
          bbar = (uint64_t)b[0] | ((uint64_t)b[1] << 32);
        } else {
          const int pos = n - 33, w = pos >> 5, off = pos & 31;  // bits [pos, pos + 33)
          auto window = [&](const uint32_t (&s)[kLimbs]) -> uint64_t {
            uint64_t lo = (uint64_t)s[w] | ((uint64_t)(w + 1 < kLimbs ? s[w + 1] : 0u) << 32);
            uint64_t hi2 = w + 2 < kLimbs ? s[w + 2] : 0u;
            uint64_t val = lo >> off;
            if (off) val |= hi2 << (64 - off);
            return val & 0x1ffffffffull;
          };
          abar = (window(a) << 31) | (a[0] & 0x7fffffffu);
          bbar = (window(b) << 31) | (b[0] & 0x7fffffffu);
        }
      }
      // ---- 31 binary-gcd steps on the approximations, branch-free
      int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
      for (int it = 0; it < 31; it++) {
        const uint64_t odd = 0ull - (abar & 1ull);
        const uint64_t sw = odd & (0ull - (uint64_t)(abar < bbar));
        const uint64_t tx = (abar ^ bbar) & sw;
        abar ^= tx;
        bbar ^= tx;
        const int64_t tf = (f0 ^ f1) & (int64_t)sw, tg = (g0 ^ g1) & (int64_t)sw;
        f0 ^= tf;
        f1 ^= tf;
        g0 ^= tg;
        g1 ^= tg;
        abar -= bbar & odd;
        f0 -= f1 & (int64_t)odd;
        g0 -= g1 & (int64_t)odd;
        abar >>= 1;
        f1 <<= 1;
        g1 <<= 1;
      }
      // ---- apply to (a, b): t = f*a + g*b (signed), exact division by 2^31, then make it non-negative
      int64_t fs[2] = {f0, f1}, gs[2] = {g0, g1};
      uint32_t na[2][kLimbs];
      for (int k = 0; k < 2; k++) {
        const bool fneg = fs[k] < 0, gneg = gs[k] < 0;
        const uint32_t fa = (uint32_t)(fneg ? -fs[k] : fs[k]), ga = (uint32_t)(gneg ? -gs[k] : gs[k]);
        uint32_t pf[kW], pg[kW], mag[kW];
        mul_small(pf, a, fa);
        mul_small(pg, b, ga);
        bool neg;
        if (fneg == gneg) {
          uint64_t c = 0;
          for (int i = 0; i < kW; i++) {
            c += (uint64_t)pf[i] + pg[i];
            mag[i] = (uint32_t)c;
            c >>= 32;
          }
          neg = fneg;  // both factors negative -> negative (a zero factor never has the sign bit set)
        } else {
          // pf*sign(f) + pg*sign(g): compute pos - negpart, negate on borrow
          const uint32_t(&posv)[kW] = fneg ? pg : pf;
          const uint32_t(&negv)[kW] = fneg ? pf : pg;
          uint64_t brw = 0;
          for (int i = 0; i < kW; i++) {
            uint64_t t = (uint64_t)posv[i] - negv[i] - brw;
            mag[i] = (uint32_t)t;
            brw = (t >> 32) & 1u;
          }
          neg = brw != 0;
          if (neg) {
            uint64_t c = 1;
            for (int i = 0; i < kW; i++) {
              c += (uint64_t)(uint32_t)~mag[i];
              mag[i] = (uint32_t)c;
              c >>= 32;
            }
          }
        }
        for (int i = 0; i < kLimbs; i++) na[k][i] = (mag[i] >> 31) | (mag[i + 1] << 1);
        if (neg) {  // keep (a, b) non-negative: flip the factors so that the (u, v) update matches
          fs[k] = -fs[k];
          gs[k] = -gs[k];
        }
      }
      // ---- apply to (u, v) mod p:  acc = p*2^31 + f*u + g*v >= 0; acc += q*p with q = acc * (-p^-1) mod 2^31;
      //      acc / 2^31 < 3p; two conditional subtractions
      uint32_t nu[2][kLimbs];
      for (int k = 0; k < 2; k++) {
        const bool fneg = fs[k] < 0, gneg = gs[k] < 0;
        const uint32_t fa = (uint32_t)(fneg ? -fs[k] : fs[k]), ga = (uint32_t)(gneg ? -gs[k] : gs[k]);
        uint32_t acc[kW], t[kW];
        acc[0] = 0;
        {
          uint32_t prev = 0;
          for (int i = 0; i < kLimbs; i++) {
            const uint32_t pi = P::p(i);
            acc[i] = (pi << 31) | (prev >> 1);
            prev = pi;
          }
          acc[kLimbs] = prev >> 1;
        }
        for (int part = 0; part < 2; part++) {
          mul_small(t, part == 0 ? u : v, part == 0 ? fa : ga);
          const bool sub = part == 0 ? fneg : gneg;
          if (!sub) {
            uint64_t c = 0;
            for (int i = 0; i < kW; i++) {
              c += (uint64_t)acc[i] + t[i];
              acc[i] = (uint32_t)c;
              c >>= 32;
            }
          } else {
            uint64_t brw = 0;
            for (int i = 0; i < kW; i++) {
              uint64_t d = (uint64_t)acc[i] - t[i] - brw;
              acc[i] = (uint32_t)d;
              brw = (d >> 32) & 1u;
            }
          }
        }
        const uint32_t q = (acc[0] * P::kInv32) & 0x7fffffffu;
        {
          uint64_t c = 0;
          for (int i = 0; i < kLimbs; i++) {
            c += (uint64_t)P::p(i) * q + acc[i];
            acc[i] = (uint32_t)c;
            c >>= 32;
          }
          acc[kLimbs] = (uint32_t)(c + acc[kLimbs]);
        }
        uint32_t res[kLimbs];
        for (int i = 0; i < kLimbs; i++) res[i] = (acc[i] >> 31) | (acc[i + 1] << 1);
        for (int rep = 0; rep < 2; rep++) {  // res < 3p
          uint32_t d[kLimbs];
          uint64_t brw = 0;
          for (int i = 0; i < kLimbs; i++) {
            uint64_t t2 = (uint64_t)res[i] - P::p(i) - brw;
            d[i] = (uint32_t)t2;
            brw = (t2 >> 32) & 1u;
          }
          const uint32_t keep = 0u - (uint32_t)brw;  // all-ones: res < p, keep it
          for (int i = 0; i < kLimbs; i++) res[i] = (res[i] & keep) | (d[i] & ~keep);
        }
        for (int i = 0; i < kLimbs; i++) nu[k][i] = res[i];
      }
      for (int i = 0; i < kLimbs; i++) {
        a[i] = na[0][i];
        b[i] = na[1][i];
        u[i] = nu[0][i];
        v[i] = nu[1][i];
      }
    }
    // b == 1 and v = x^-1 as a plain integer (x = the Montgomery residue): one multiplication by R^3 -> a^-1 * R
    Fp vi, r3;
    for (int i = 0; i < kLimbs; i++) {
      vi.l[i] = v[i];
      r3.l[i] = P::r3(i);
    }
    mul(r, vi, r3);
  }
  // r = a^-1; a != 0 (0 maps to 0). The reference uses an extended gcd (fp.tcc:641-685); the inverse is unique, so any
  // algorithm gives the same canonical bytes. Host and device: batched binary gcd (~10x fewer instructions than the
  // bitwise one; it is what makes the shared inversion of the batch-affine accumulation cheap). -DB200_INV_BITWISE
  // selects the bitwise routine on the device (round 1's default).
  // (out of line on the device: its ~1 KB of scratch arrays and their live ranges then stay out of the calling kernel's
  // register allocation)
  B200_HD static B200_NOINLINE void inv(Fp &r, const Fp &a) {
#if defined(__CUDA_ARCH__) && defined(B200_INV_BITWISE)
    inv_binary(r, a);
#else
    inv_bingcd(r, a);
#endif
  }
  // a^(p-2) (Fermat): kept as an independent cross-check of the gcd-based inversions (tests)
  B200_HD static void inv_fermat(Fp &r, const Fp &a) {
    uint32_t e[kLimbs];
    for (int i = 0; i < kLimbs; i++) e[i] = P::p(i);
    e[0] -= 2;  // p is odd and p mod 2^32 >= 3, no borrow
    pow_words(r, a, e, kLimbs);
  }
};

// ------------------------------------------------------------------------------------------------ Fq2 = Fq[u]/(u^2 - 13)
// (depends/libff/libff/algebra/fields/fp2.tcc:78-126; non-residue 13: mnt4753_init.cpp:105)
template <class P, unsigned NR>
struct alignas(16) Fp2 {
  typedef Fp<P> B;
  typedef P Prime;
  static constexpr int kDegree = 2;
  B c0, c1;

  B200_HD static void set_zero(Fp2 &r) { B::set_zero(r.c0); B::set_zero(r.c1); }
  B200_HD static void set_one(Fp2 &r) { B::set_one(r.c0); B::set_zero(r.c1); }
  B200_HD static bool is_zero(const Fp2 &a) { return B::is_zero(a.c0) && B::is_zero(a.c1); }
  B200_HD static bool eq(const Fp2 &a, const Fp2 &b) { return B::eq(a.c0, b.c0) && B::eq(a.c1, b.c1); }
  B200_HD static B200_NOINLINE void add(Fp2 &r, const Fp2 &a, const Fp2 &b) { B::add_ni(r.c0, a.c0, b.c0); B::add_ni(r.c1, a.c1, b.c1); }
  B200_HD static B200_NOINLINE void sub(Fp2 &r, const Fp2 &a, const Fp2 &b) { B::sub_ni(r.c0, a.c0, b.c0); B::sub_ni(r.c1, a.c1, b.c1); }
  B200_HD static B200_NOINLINE void dbl(Fp2 &r, const Fp2 &a) { B::add_ni(r.c0, a.c0, a.c0); B::add_ni(r.c1, a.c1, a.c1); }
  B200_HD static B200_NOINLINE void neg(Fp2 &r, const Fp2 &a) { B::neg_ni(r.c0, a.c0); B::neg_ni(r.c1, a.c1); }
  B200_HD static B200_INLINE void neg_ni(Fp2 &r, const Fp2 &a) { neg(r, a); }
  B200_HD static B200_NOINLINE void mul(Fp2 &r, const Fp2 &a, const Fp2 &b) {  // Karatsuba, 3 base multiplications
    B aA, bB, s, t;
    B::mul(aA, a.c0, b.c0);
    B::mul(bB, a.c1, b.c1);
    B::add_ni(s, a.c0, a.c1);
    B::add_ni(t, b.c0, b.c1);
    B::mul(s, s, t);
    B::sub_ni(s, s, aA);
    B::sub_ni(r.c1, s, bB);
    B::template mul_small_ni<NR>(t, bB);
    B::add_ni(r.c0, aA, t);
  }
  B200_HD static B200_NOINLINE void sqr(Fp2 &r, const Fp2 &a) {  // complex squaring, 2 base multiplications
    B ab, s, t;
    B::mul(ab, a.c0, a.c1);
    B::add_ni(s, a.c0, a.c1);
    B::template mul_small_ni<NR>(t, a.c1);
    B::add_ni(t, t, a.c0);
    B::mul(s, s, t);
    B::sub_ni(s, s, ab);
    B::template mul_small_ni<NR>(t, ab);
    B::sub_ni(r.c0, s, t);
    B::add_ni(r.c1, ab, ab);
  }
  B200_HD static B200_INLINE void sqr_fast(Fp2 &r, const Fp2 &a) { sqr(r, a); }  // complex squaring has no base squarings
  B200_HD static void inv(Fp2 &r, const Fp2 &a) {  // fp2.tcc:128-142
    B t0, t1, t2;
    B::sqr(t0, a.c0);
    B::sqr(t1, a.c1);
    B::template mul_small_ni<NR>(t1, t1);
    B::sub_ni(t2, t0, t1);
    B::inv(t2, t2);
    B::mul(r.c0, a.c0, t2);
    B::mul(t0, a.c1, t2);
    B::neg_ni(r.c1, t0);
  }
};

// ------------------------------------------------------------------------------------------------ Fq3 = Fq[u]/(u^3 - 11)
// (depends/libff/libff/algebra/fields/fp3.tcc:82-123; non-residue 11: mnt6753_init.cpp:109)
template <class P, unsigned NR>
struct alignas(16) Fp3 {
  typedef Fp<P> B;
  typedef P Prime;
  static constexpr int kDegree = 3;
  B c0, c1, c2;

  B200_HD static void set_zero(Fp3 &r) { B::set_zero(r.c0); B::set_zero(r.c1); B::set_zero(r.c2); }
  B200_HD static void set_one(Fp3 &r) { B::set_one(r.c0); B::set_zero(r.c1); B::set_zero(r.c2); }
  B200_HD static bool is_zero(const Fp3 &a) { return B::is_zero(a.c0) && B::is_zero(a.c1) && B::is_zero(a.c2); }
  B200_HD static bool eq(const Fp3 &a, const Fp3 &b) {
    return B::eq(a.c0, b.c0) && B::eq(a.c1, b.c1) && B::eq(a.c2, b.c2);
  }
  B200_HD static B200_NOINLINE void add(Fp3 &r, const Fp3 &a, const Fp3 &b) {
    B::add_ni(r.c0, a.c0, b.c0); B::add_ni(r.c1, a.c1, b.c1); B::add_ni(r.c2, a.c2, b.c2);
  }
  B200_HD static B200_NOINLINE void sub(Fp3 &r, const Fp3 &a, const Fp3 &b) {
    B::sub_ni(r.c0, a.c0, b.c0); B::sub_ni(r.c1, a.c1, b.c1); B::sub_ni(r.c2, a.c2, b.c2);
  }
  B200_HD static B200_NOINLINE void dbl(Fp3 &r, const Fp3 &a) { B::add_ni(r.c0, a.c0, a.c0); B::add_ni(r.c1, a.c1, a.c1); B::add_ni(r.c2, a.c2, a.c2); }
  B200_HD static B200_NOINLINE void neg(Fp3 &r, const Fp3 &a) { B::neg_ni(r.c0, a.c0); B::neg_ni(r.c1, a.c1); B::neg_ni(r.c2, a.c2); }
  B200_HD static B200_INLINE void neg_ni(Fp3 &r, const Fp3 &a) { neg(r, a); }
  B200_HD static B200_NOINLINE void mul(Fp3 &r, const Fp3 &a, const Fp3 &b) {  // Karatsuba, 6 base multiplications
    B aA, bB, cC, s, t, u;
    B::mul(aA, a.c0, b.c0);
    B::mul(bB, a.c1, b.c1);
    B::mul(cC, a.c2, b.c2);
    // c0 = aA + nr*((b+c)(B+C) - bB - cC)
    B::add_ni(s, a.c1, a.c2);
    B::add_ni(t, b.c1, b.c2);
    B::mul(s, s, t);
    B::sub_ni(s, s, bB);
    B::sub_ni(s, s, cC);
    B::template mul_small_ni<NR>(s, s);
    // c1 = (a+b)(A+B) - aA - bB + nr*cC
    B::add_ni(t, a.c0, a.c1);
    B::add_ni(u, b.c0, b.c1);
    B::mul(t, t, u);
    B::sub_ni(t, t, aA);
    B::sub_ni(t, t, bB);
    B::template mul_small_ni<NR>(u, cC);
    B::add_ni(t, t, u);
    // c2 = (a+c)(A+C) - aA + bB - cC
    B u2, v2;
    B::add_ni(u2, a.c0, a.c2);
    B::add_ni(v2, b.c0, b.c2);
    B::mul(u2, u2, v2);
    B::sub_ni(u2, u2, aA);
    B::add_ni(u2, u2, bB);
    B::sub_ni(r.c2, u2, cC);
    B::add_ni(r.c0, aA, s);
    r.c1 = t;
  }
  B200_HD static B200_NOINLINE void sqr(Fp3 &r, const Fp3 &a) {  // CH-SQR2: 3 squarings + 2 multiplications
    B s0, s1, s2, s3, s4, t;
    B::sqr(s0, a.c0);
    B::mul(s1, a.c0, a.c1);
    B::add_ni(s1, s1, s1);
    B::sub_ni(t, a.c0, a.c1);
    B::add_ni(t, t, a.c2);
    B::sqr(s2, t);
    B::mul(s3, a.c1, a.c2);
    B::add_ni(s3, s3, s3);
    B::sqr(s4, a.c2);
    // c0 = s0 + nr*s3 ; c1 = s1 + nr*s4 ; c2 = s1 + s2 + s3 - s0 - s4
    B::template mul_small_ni<NR>(t, s3);
    B::add_ni(r.c0, s0, t);
    B::template mul_small_ni<NR>(t, s4);
    B::add_ni(r.c1, s1, t);
    B::add_ni(t, s1, s2);
    B::add_ni(t, t, s3);
    B::sub_ni(t, t, s0);
    B::sub_ni(r.c2, t, s4);
  }
  B200_HD static B200_NOINLINE void sqr_fast(Fp3 &r, const Fp3 &a) {  // CH-SQR2 with the dedicated base squaring
    B s0, s1, s2, s3, s4, t;
    B::sqr_fast(s0, a.c0);
    B::mul(s1, a.c0, a.c1);
    B::add_ni(s1, s1, s1);
    B::sub_ni(t, a.c0, a.c1);
    B::add_ni(t, t, a.c2);
    B::sqr_fast(s2, t);
    B::mul(s3, a.c1, a.c2);
    B::add_ni(s3, s3, s3);
    B::sqr_fast(s4, a.c2);
    B::template mul_small_ni<NR>(t, s3);
    B::add_ni(r.c0, s0, t);
    B::template mul_small_ni<NR>(t, s4);
    B::add_ni(r.c1, s1, t);
    B::add_ni(t, s1, s2);
    B::add_ni(t, t, s3);
    B::sub_ni(t, t, s0);
    B::sub_ni(r.c2, t, s4);
  }
  B200_HD static void inv(Fp3 &r, const Fp3 &a) {  // fp3.tcc:125-143
    B t0, t1, t2, t3, t4, t5, c0, c1, c2, t6, u;
    B::sqr(t0, a.c0);
    B::sqr(t1, a.c1);
    B::sqr(t2, a.c2);
    B::mul(t3, a.c0, a.c1);
    B::mul(t4, a.c0, a.c2);
    B::mul(t5, a.c1, a.c2);
    B::template mul_small_ni<NR>(u, t5);
    B::sub_ni(c0, t0, u);
    B::template mul_small_ni<NR>(u, t2);
    B::sub_ni(c1, u, t3);
    B::sub_ni(c2, t1, t4);
    B::mul(t6, a.c0, c0);
    B::mul(t0, a.c2, c1);
    B::mul(t1, a.c1, c2);
    B::add_ni(t0, t0, t1);
    B::template mul_small_ni<NR>(t0, t0);
    B::add_ni(t6, t6, t0);
    B::inv(t6, t6);
    B::mul(r.c0, t6, c0);
    B::mul(r.c1, t6, c1);
    B::mul(r.c2, t6, c2);
  }
};

}  // namespace b200
