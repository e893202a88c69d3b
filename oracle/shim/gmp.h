/* Minimal declaration shim for the GMP entry points that libff/libfqfft use.
 *
 * TEST INFRASTRUCTURE ONLY.  The image ships the GMP runtime (libgmp.so.10) but not its
 * development header, so the recipe in oracle/Makefile compiles the reference's libff sources
 * against these prototypes and links with -l:libgmp.so.10.  Only the 24 functions libff calls
 * are declared; the names map onto the exported __gmpn_* / __gmpz_* symbols exactly as the
 * real <gmp.h> does.
 */
#ifndef ORACLE_SHIM_GMP_H
#define ORACLE_SHIM_GMP_H

#include <stddef.h>
#include <stdio.h>

typedef unsigned long mp_limb_t;
typedef long mp_limb_signed_t;
typedef long mp_size_t;
typedef unsigned long mp_bitcnt_t;
typedef mp_limb_t *mp_ptr;
typedef const mp_limb_t *mp_srcptr;

#define GMP_LIMB_BITS 64
#define GMP_NAIL_BITS 0
#define GMP_NUMB_BITS 64

typedef struct {
    int _mp_alloc;
    int _mp_size;
    mp_limb_t *_mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

#ifdef __cplusplus
extern "C" {
#endif

mp_limb_t __gmpn_add_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t __gmpn_add_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_addmul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
int __gmpn_cmp(mp_srcptr, mp_srcptr, mp_size_t);
void __gmpn_copyi(mp_ptr, mp_srcptr, mp_size_t);
mp_size_t __gmpn_gcdext(mp_ptr, mp_ptr, mp_size_t *, mp_ptr, mp_size_t, mp_ptr, mp_size_t);
mp_limb_t __gmpn_mul(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
void __gmpn_mul_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_rshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_size_t __gmpn_set_str(mp_ptr, const unsigned char *, size_t, int);
mp_limb_t __gmpn_sub(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_sub_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t __gmpn_sub_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
void __gmpn_tdiv_qr(mp_ptr, mp_ptr, mp_size_t, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
void __gmpn_zero(mp_ptr, mp_size_t);

void __gmpz_init(mpz_ptr);
void __gmpz_init_set(mpz_ptr, mpz_srcptr);
void __gmpz_clear(mpz_ptr);
void __gmpz_set_ui(mpz_ptr, unsigned long);
unsigned long __gmpz_get_ui(mpz_srcptr);
void __gmpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void __gmpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);

int __gmp_printf(const char *, ...);

#ifdef __cplusplus
}
#endif

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

#define mpz_init __gmpz_init
#define mpz_init_set __gmpz_init_set
#define mpz_clear __gmpz_clear
#define mpz_set_ui __gmpz_set_ui
#define mpz_get_ui __gmpz_get_ui
#define mpz_add_ui __gmpz_add_ui
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
#define mpz_sgn(Z) ((Z)->_mp_size < 0 ? -1 : (Z)->_mp_size > 0)

#define gmp_printf __gmp_printf

#endif
