/* gmp.h shim (oracle build infrastructure, NOT product code).
 * The image has libgmp.so.10 but no gmp.h; this declares exactly the 28 GMP entry points that the
 * reference (libff / libfqfft / libsnark) uses so that oracle/build_ref.sh can compile the UNMODIFIED reference
 * sources where they lie under /root/reference. See SURVEY.md section 8c / Appendix A. */
#ifndef __GMP_H__
#define __GMP_H__
#include <stddef.h>
#include <stdio.h>
#ifdef __cplusplus
#include <iosfwd>
extern "C" {
#endif
typedef unsigned long mp_limb_t;
typedef long mp_limb_signed_t;
typedef unsigned long mp_bitcnt_t;
typedef long mp_size_t;
typedef long mp_exp_t;
#define GMP_LIMB_BITS 64
#define GMP_NAIL_BITS 0
#define GMP_NUMB_BITS 64
typedef struct { int _mp_alloc; int _mp_size; mp_limb_t *_mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef mp_limb_t *mp_ptr; typedef const mp_limb_t *mp_srcptr;
typedef __mpz_struct *mpz_ptr; typedef const __mpz_struct *mpz_srcptr;
#define mpn_add_1 __gmpn_add_1
#define mpn_add_n __gmpn_add_n
#define mpn_addmul_1 __gmpn_addmul_1
#define mpn_cmp __gmpn_cmp
#define mpn_copyi __gmpn_copyi
#define mpn_gcdext __gmpn_gcdext
#define mpn_mul __gmpn_mul
#define mpn_mul_n __gmpn_mul_n
#define mpn_rshift __gmpn_rshift
#define mpn_set_str __gmpn_set_str
#define mpn_sub __gmpn_sub
#define mpn_sub_1 __gmpn_sub_1
#define mpn_sub_n __gmpn_sub_n
#define mpn_tdiv_qr __gmpn_tdiv_qr
#define mpn_zero __gmpn_zero
#define mpz_add_ui __gmpz_add_ui
#define mpz_clear __gmpz_clear
#define mpz_export __gmpz_export
#define mpz_fdiv_q __gmpz_fdiv_q
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
#define mpz_get_ui __gmpz_get_ui
#define mpz_import __gmpz_import
#define mpz_init __gmpz_init
#define mpz_init_set __gmpz_init_set
#define mpz_mod __gmpz_mod
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_set_ui __gmpz_set_ui
#define gmp_printf __gmp_printf
mp_limb_t mpn_add_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t mpn_add_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_addmul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
int mpn_cmp(mp_srcptr, mp_srcptr, mp_size_t);
void mpn_copyi(mp_ptr, mp_srcptr, mp_size_t);
mp_size_t mpn_gcdext(mp_ptr, mp_ptr, mp_size_t *, mp_ptr, mp_size_t, mp_ptr, mp_size_t);
mp_limb_t mpn_mul(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
void mpn_mul_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_rshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_size_t mpn_set_str(mp_ptr, const unsigned char *, size_t, int);
mp_limb_t mpn_sub(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
mp_limb_t mpn_sub_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t mpn_sub_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
void mpn_tdiv_qr(mp_ptr, mp_ptr, mp_size_t, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
void mpn_zero(mp_ptr, mp_size_t);
void mpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_clear(mpz_ptr);
void *mpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
void mpz_fdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
unsigned long mpz_get_ui(mpz_srcptr);
void mpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void mpz_init(mpz_ptr);
void mpz_init_set(mpz_ptr, mpz_srcptr);
void mpz_mod(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_set_ui(mpz_ptr, unsigned long);
int gmp_printf(const char *, ...);
#define mpz_sgn(Z) ((Z)->_mp_size < 0 ? -1 : (Z)->_mp_size > 0)
#ifdef __cplusplus
}
#endif
#endif
