// Short-Weierstrass group arithmetic for MNT4753 / MNT6753 G1 and G2, homogeneous projective (X:Y:Z), O = (0:1:0).
// Same coordinate system, same special-case behaviour (O, P+P falls into doubling, P+(-P) yields X=Z=0) as the
// reference: depends/libff/libff/algebra/curves/mnt753/mnt4753/mnt4753_g1.cpp:95-98 (is_zero), 134-207 (operator+),
// 265-313 (mixed_add), 315-346 (dbl), 68-83 (to_affine); G2: mnt4753_g2.cpp:31-34 and mnt6753_g2.cpp:38-41 (mul_by_a).
// Formulas are the EFD "add-1998-cmo-2" / "dbl-2007-bl" ones those files cite. Host and device share this code.
#pragma once
#include "field.cuh"

namespace b200 {

template <class F>
struct alignas(16) Affine {  // wire format of libsnark/serialization.hpp:43-67,83-111: (x, y); y == 0 encodes O
  F x, y;
};

template <class F>
struct alignas(16) Proj {
  F X, Y, Z;
};

// ---- group descriptors -------------------------------------------------------------------------------------
// curve coefficient a: MNT4753 G1 a=2 (mnt4753_init.cpp:119), G2 a'=(2*13, 0) applied componentwise (:127-128);
// MNT6753 G1 a=11 (mnt6753_init.cpp:130), G2 a'=(0,0,11): (c0,c1,c2) -> (11*11*c1, 11*11*c2, 11*c0) (:140-142).
struct Mnt4G1 {
  typedef Fp<PrimeB> F;
  typedef PrimeA ScalarPrime;
  B200_HD static void mul_by_a(F &r, const F &x) { F::dbl(r, x); }
};
struct Mnt4G2 {
  typedef Fp2<PrimeB, 13> F;
  typedef PrimeA ScalarPrime;
  B200_HD static void mul_by_a(F &r, const F &x) {
    F::B::template mul_small_ni<26>(r.c0, x.c0);
    F::B::template mul_small_ni<26>(r.c1, x.c1);
  }
};
struct Mnt6G1 {
  typedef Fp<PrimeA> F;
  typedef PrimeB ScalarPrime;
  B200_HD static void mul_by_a(F &r, const F &x) { F::template mul_small<11>(r, x); }
};
struct Mnt6G2 {
  typedef Fp3<PrimeA, 11> F;
  typedef PrimeB ScalarPrime;
  B200_HD static void mul_by_a(F &r, const F &x) {
    typename F::B t0, t1, t2;
    F::B::template mul_small_ni<121>(t0, x.c1);
    F::B::template mul_small_ni<121>(t1, x.c2);
    F::B::template mul_small_ni<11>(t2, x.c0);
    r.c0 = t0;
    r.c1 = t1;
    r.c2 = t2;
  }
};

// ---- point operations --------------------------------------------------------------------------------------
template <class F>
B200_HD inline void proj_set_zero(Proj<F> &p) {
  F::set_zero(p.X);
  F::set_one(p.Y);
  F::set_zero(p.Z);
}
template <class F>
B200_HD inline bool proj_is_zero(const Proj<F> &p) {
  return F::is_zero(p.X) && F::is_zero(p.Z);
}
template <class F>
B200_HD inline bool affine_is_zero(const Affine<F> &p) {
  return F::is_zero(p.y);  // reader rule: y == 0 -> O (serialization.hpp:87-89,107-109)
}
template <class F>
B200_HD inline void proj_from_affine(Proj<F> &r, const Affine<F> &a) {
  if (affine_is_zero(a)) {
    proj_set_zero(r);
  } else {
    r.X = a.x;
    r.Y = a.y;
    F::set_one(r.Z);
  }
}

template <class G>
B200_HD B200_NOINLINE void proj_dbl(Proj<typename G::F> &r, const Proj<typename G::F> &p) {
  typedef typename G::F F;
  if (proj_is_zero(p)) {
    r = p;
    return;
  }
  F XX, w, s, t, R, RR, B;
  F::sqr(XX, p.X);          // XX = X1^2
  F::sqr(t, p.Z);           // ZZ
  G::mul_by_a(w, t);        // a*ZZ
  F::add(w, w, XX);
  F::add(w, w, XX);
  F::add(w, w, XX);         // w = a*ZZ + 3*XX
  F::mul(s, p.Y, p.Z);
  F::dbl(s, s);             // s = 2*Y1*Z1
  F::mul(R, p.Y, s);        // R = Y1*s
  F::sqr(RR, R);            // RR
  F::add(B, p.X, R);
  F::sqr(B, B);
  F::sub(B, B, XX);
  F::sub(B, B, RR);         // B = (X1+R)^2 - XX - RR
  F::sqr(t, w);
  F::sub(t, t, B);
  F::sub(t, t, B);          // h = w^2 - 2B
  F::mul(r.X, t, s);        // X3 = h*s
  F::sub(B, B, t);
  F::mul(B, w, B);
  F::sub(B, B, RR);
  F::sub(r.Y, B, RR);       // Y3 = w*(B-h) - 2*RR
  F::sqr(t, s);
  F::mul(r.Z, s, t);        // Z3 = s^3
}

// r = p + q, both projective (general addition with the reference's built-in doubling detection)
template <class G>
B200_HD void proj_add(Proj<typename G::F> &r, const Proj<typename G::F> &p, const Proj<typename G::F> &q) {
  typedef typename G::F F;
  if (proj_is_zero(p)) {
    r = q;
    return;
  }
  if (proj_is_zero(q)) {
    r = p;
    return;
  }
  F X1Z2, t0, Y1Z2, t1, t2, t3, t4;
  F::mul(X1Z2, p.X, q.Z);
  F::mul(t0, p.Z, q.X);     // X2Z1
  F::mul(Y1Z2, p.Y, q.Z);
  F::mul(t1, p.Z, q.Y);     // Y2Z1
  if (F::eq(X1Z2, t0) && F::eq(Y1Z2, t1)) {
    proj_dbl<G>(r, p);
    return;
  }
  F::mul(t4, p.Z, q.Z);     // Z1Z2
  F::sub(t1, t1, Y1Z2);     // u
  F::sub(t0, t0, X1Z2);     // v
  F::sqr(t2, t1);           // uu
  F::sqr(t3, t0);           // vv
  F::mul(t2, t2, t4);       // uu*Z1Z2
  F::mul(X1Z2, t3, X1Z2);   // R = vv*X1Z2
  F::mul(t3, t0, t3);       // vvv
  F::sub(t2, t2, t3);
  F::sub(t2, t2, X1Z2);
  F::sub(t2, t2, X1Z2);     // A = uu*Z1Z2 - vvv - 2R
  F::mul(r.X, t0, t2);      // X3 = v*A
  F::sub(X1Z2, X1Z2, t2);   // R - A
  F::mul(X1Z2, t1, X1Z2);   // u*(R-A)
  F::mul(Y1Z2, t3, Y1Z2);   // vvv*Y1Z2
  F::sub(r.Y, X1Z2, Y1Z2);  // Y3
  F::mul(r.Z, t3, t4);      // Z3 = vvv*Z1Z2
}

// acc += q, q affine and NOT the point at infinity (callers filter y == 0)
template <class G>
B200_HD void proj_madd(Proj<typename G::F> &acc, const Affine<typename G::F> &q) {
  typedef typename G::F F;
  if (proj_is_zero(acc)) {
    acc.X = q.x;
    acc.Y = q.y;
    F::set_one(acc.Z);
    return;
  }
  F t0, t1, t2, t3, t4;
  F::mul(t0, q.x, acc.Z);   // X2Z1
  F::mul(t1, q.y, acc.Z);   // Y2Z1
  if (F::eq(t0, acc.X) && F::eq(t1, acc.Y)) {
    proj_dbl<G>(acc, acc);
    return;
  }
  F::sub(t1, t1, acc.Y);    // u
  F::sub(t0, t0, acc.X);    // v
  F::sqr(t2, t1);           // uu
  F::sqr(t3, t0);           // vv
  F::mul(t4, t0, t3);       // vvv
  F::mul(t3, t3, acc.X);    // R = vv*X1
  F::mul(t2, t2, acc.Z);    // uu*Z1
  F::sub(t2, t2, t4);
  F::sub(t2, t2, t3);
  F::sub(t2, t2, t3);       // A
  F::mul(acc.X, t0, t2);    // X3 = v*A
  F::sub(t3, t3, t2);       // R - A
  F::mul(t3, t1, t3);       // u*(R-A)
  F::mul(t2, t4, acc.Y);    // vvv*Y1
  F::sub(acc.Y, t3, t2);    // Y3
  F::mul(acc.Z, t4, acc.Z); // Z3 = vvv*Z1
}

// ---- XYZZ accumulator (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; O <=> ZZ == 0) ----------------------------------------
// Used only inside the MSM bucket accumulation: the mixed addition costs 8M + 2S (EFD "madd-2008-s") instead of the
// 9M + 2S of the homogeneous-projective formula the reference uses (mnt4753_g1.cpp:265-313). Same special cases:
// O + Q = Q, P + P -> doubling ("mdbl-2008-s-1", needs the curve's a), P + (-P) = O. The result is converted back
// to the reference's (X:Y:Z) before anything else sees it; the group element is the same.
template <class F>
struct alignas(16) XYZZ {
  F X, Y, ZZ, ZZZ;
};
template <class F>
B200_HD inline void xyzz_set_zero(XYZZ<F> &p) {
  F::set_zero(p.X);
  F::set_zero(p.Y);
  F::set_zero(p.ZZ);
  F::set_zero(p.ZZZ);
}
template <class G>
B200_HD void xyzz_madd(XYZZ<typename G::F> &acc, const Affine<typename G::F> &q) {
  typedef typename G::F F;
  if (F::is_zero(acc.ZZ)) {
    acc.X = q.x;
    acc.Y = q.y;
    F::set_one(acc.ZZ);
    F::set_one(acc.ZZZ);
    return;
  }
  F P, R, t0, t1;
  F::mul(P, q.x, acc.ZZ);    // U2
  F::mul(R, q.y, acc.ZZZ);   // S2
  F::sub(P, P, acc.X);       // P = U2 - X1
  F::sub(R, R, acc.Y);       // R = S2 - Y1
  if (F::is_zero(P)) {
    if (!F::is_zero(R)) {    // Q = -acc
      xyzz_set_zero(acc);
      return;
    }
    // Q == acc: double the affine point Q
    F one, a;
    F::set_one(one);
    G::mul_by_a(a, one);
    F::dbl(t0, q.y);         // U = 2*Y1
    F::sqr(acc.ZZ, t0);      // V = U^2
    F::mul(acc.ZZZ, t0, acc.ZZ);  // W = U*V
    F::mul(t1, q.x, acc.ZZ); // S = X1*V
    F::sqr(t0, q.x);
    F::add(P, t0, t0);
    F::add(P, P, t0);
    F::add(P, P, a);         // M = 3*X1^2 + a
    F::sqr(acc.X, P);
    F::sub(acc.X, acc.X, t1);
    F::sub(acc.X, acc.X, t1);  // X3 = M^2 - 2S
    F::sub(t1, t1, acc.X);
    F::mul(t1, P, t1);       // M*(S - X3)
    F::mul(t0, acc.ZZZ, q.y);  // W*Y1
    F::sub(acc.Y, t1, t0);
    return;
  }
  F::sqr(t0, P);             // PP
  F::mul(t1, P, t0);         // PPP
  F::mul(P, acc.X, t0);      // Q = X1*PP
  F::mul(acc.ZZ, acc.ZZ, t0);    // ZZ3 = ZZ1*PP
  F::mul(acc.ZZZ, acc.ZZZ, t1);  // ZZZ3 = ZZZ1*PPP
  F::sqr(t0, R);
  F::sub(t0, t0, t1);
  F::sub(t0, t0, P);
  F::sub(t0, t0, P);         // X3 = R^2 - PPP - 2Q
  F::sub(P, P, t0);          // Q - X3
  F::mul(P, R, P);           // R*(Q - X3)
  F::mul(t1, acc.Y, t1);     // Y1*PPP
  F::sub(acc.Y, P, t1);      // Y3
  acc.X = t0;
}
// (X:Y:Z) = (X*ZZZ : Y*ZZ : ZZ*ZZZ)
template <class F>
B200_HD inline void xyzz_to_proj(Proj<F> &r, const XYZZ<F> &p) {
  if (F::is_zero(p.ZZ)) {
    proj_set_zero(r);
    return;
  }
  F::mul(r.X, p.X, p.ZZZ);
  F::mul(r.Y, p.Y, p.ZZ);
  F::mul(r.Z, p.ZZ, p.ZZZ);
}

// ---- Jacobian doubling (x = X/Z^2, y = Y/Z^3) --------------------------------------------------------------------
// Used only by the base-table builder (msm_precompute_kernel), which is nothing but doublings - 753 per base: EFD
// "dbl-2007-bl" costs 1M + 8S (+ one multiplication by the curve's small a) against the 6M + 6S of the homogeneous
// doubling above, and its squarings go through the dedicated squaring (sqr_fast). The affine results are the same
// canonical bytes whatever the coordinate system. Never called on O (the builder filters y == 0); a point of order 2
// does not exist in these prime-order groups.
template <class G>
B200_HD void jac_dbl(Proj<typename G::F> &r, const Proj<typename G::F> &p) {
  typedef typename G::F F;
  F XX, YY, YYYY, ZZ, S, M, t;
  F::sqr_fast(XX, p.X);
  F::sqr_fast(YY, p.Y);
  F::sqr_fast(YYYY, YY);
  F::sqr_fast(ZZ, p.Z);
  F::add(S, p.X, YY);
  F::sqr_fast(S, S);
  F::sub(S, S, XX);
  F::sub(S, S, YYYY);
  F::dbl(S, S);              // S = 2((X+YY)^2 - XX - YYYY)
  F::sqr_fast(t, ZZ);
  G::mul_by_a(M, t);         // a*ZZ^2
  F::add(M, M, XX);
  F::add(M, M, XX);
  F::add(M, M, XX);          // M = 3XX + a*ZZ^2
  F::add(t, p.Y, p.Z);
  F::sqr_fast(t, t);
  F::sub(t, t, YY);
  F::sub(r.Z, t, ZZ);        // Z3 = (Y+Z)^2 - YY - ZZ   (p.Y, p.Z no longer needed after this line)
  F::sqr_fast(t, M);
  F::sub(t, t, S);
  F::sub(t, t, S);           // X3 = M^2 - 2S
  F::sub(S, S, t);
  F::mul(S, M, S);           // M(S - X3)
  F::dbl(YYYY, YYYY);
  F::dbl(YYYY, YYYY);
  F::dbl(YYYY, YYYY);        // 8 YYYY
  F::sub(r.Y, S, YYYY);
  r.X = t;
}

template <class F>
B200_HD inline void proj_neg(Proj<F> &r, const Proj<F> &p) {
  r.X = p.X;
  F::neg(r.Y, p.Y);
  r.Z = p.Z;
}

// MSB-first double-and-add, scalar = plain integer, little-endian u32 words (curve_utils.tcc:13-34)
template <class G>
B200_HD void proj_scalar_mul(Proj<typename G::F> &r, const Proj<typename G::F> &p, const uint32_t *k, int nwords) {
  Proj<typename G::F> acc;
  proj_set_zero(acc);
  bool found = false;
  for (int w = nwords - 1; w >= 0; w--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (found) proj_dbl<G>(acc, acc);
      if ((k[w] >> bit) & 1) {
        found = true;
        proj_add<G>(acc, acc, p);
      }
    }
  }
  r = acc;
}

// affine (x, y) in wire format; O -> (0, 0)  (serialization.hpp:43-54 + mnt4753_g1.cpp:68-83)
template <class G>
B200_HD void proj_to_affine(Affine<typename G::F> &a, const Proj<typename G::F> &p) {
  typedef typename G::F F;
  if (proj_is_zero(p)) {
    F::set_zero(a.x);
    F::set_zero(a.y);
    return;
  }
  F zi;
  F::inv(zi, p.Z);
  F::mul(a.x, p.X, zi);
  F::mul(a.y, p.Y, zi);
}

}  // namespace b200
