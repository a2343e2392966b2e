/* oracle/shim/gmp.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Minimal declaration shim for the GNU MP C ABI, so that the UNMODIFIED
 * reference sources under /root/reference/src can be compiled into
 * oracle/_ref/ in an image that ships the GMP runtime (libgmp.so.10) but not
 * its development headers.  Only declarations live here; every function is
 * resolved at link time against the system libgmp.so.10.  The struct layout
 * and the __gmpz_ symbol prefix are GMP's documented, stable ABI (GMP 4.x-6.x).
 */
#ifndef TB_ORACLE_SHIM_GMP_H
#define TB_ORACLE_SHIM_GMP_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t;
typedef unsigned long mp_bitcnt_t;

typedef struct {
    int _mp_alloc;
    int _mp_size;
    mp_limb_t *_mp_d;
} __mpz_struct;

typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

#define TB_GMPZ(name) __gmpz_##name

void TB_GMPZ(init)(mpz_ptr);
void TB_GMPZ(inits)(mpz_ptr, ...);
void TB_GMPZ(clear)(mpz_ptr);
void TB_GMPZ(clears)(mpz_ptr, ...);
void TB_GMPZ(init_set)(mpz_ptr, mpz_srcptr);
void TB_GMPZ(init_set_si)(mpz_ptr, long);
void TB_GMPZ(init_set_ui)(mpz_ptr, unsigned long);
void TB_GMPZ(set)(mpz_ptr, mpz_srcptr);
void TB_GMPZ(set_si)(mpz_ptr, long);
void TB_GMPZ(set_ui)(mpz_ptr, unsigned long);
int TB_GMPZ(set_str)(mpz_ptr, const char *, int);
void TB_GMPZ(add)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void TB_GMPZ(add_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void TB_GMPZ(sub)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void TB_GMPZ(sub_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void TB_GMPZ(mul)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void TB_GMPZ(mul_si)(mpz_ptr, mpz_srcptr, long);
void TB_GMPZ(mul_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void TB_GMPZ(mul_2exp)(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void TB_GMPZ(addmul)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void TB_GMPZ(submul)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void TB_GMPZ(neg)(mpz_ptr, mpz_srcptr);
void TB_GMPZ(abs)(mpz_ptr, mpz_srcptr);
void TB_GMPZ(swap)(mpz_ptr, mpz_ptr);
int TB_GMPZ(cmp)(mpz_srcptr, mpz_srcptr);
int TB_GMPZ(cmp_si)(mpz_srcptr, long);
int TB_GMPZ(cmp_ui)(mpz_srcptr, unsigned long);
int TB_GMPZ(cmpabs)(mpz_srcptr, mpz_srcptr);
void TB_GMPZ(fdiv_q)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void TB_GMPZ(fdiv_q_2exp)(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void TB_GMPZ(tdiv_q)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void TB_GMPZ(tdiv_r)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void TB_GMPZ(tdiv_q_2exp)(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void TB_GMPZ(divexact)(mpz_ptr, mpz_srcptr, mpz_srcptr);
unsigned long TB_GMPZ(fdiv_r_ui)(mpz_ptr, mpz_srcptr, unsigned long);
long TB_GMPZ(get_si)(mpz_srcptr);
unsigned long TB_GMPZ(get_ui)(mpz_srcptr);
double TB_GMPZ(get_d)(mpz_srcptr);
char *TB_GMPZ(get_str)(char *, int, mpz_srcptr);
size_t TB_GMPZ(sizeinbase)(mpz_srcptr, int);
int TB_GMPZ(fits_slong_p)(mpz_srcptr);
int TB_GMPZ(jacobi)(mpz_srcptr, mpz_srcptr);
int TB_GMPZ(invert)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void TB_GMPZ(nextprime)(mpz_ptr, mpz_srcptr);

#ifdef __cplusplus
}
#endif

#define mpz_init        __gmpz_init
#define mpz_inits       __gmpz_inits
#define mpz_clear       __gmpz_clear
#define mpz_clears      __gmpz_clears
#define mpz_init_set    __gmpz_init_set
#define mpz_init_set_si __gmpz_init_set_si
#define mpz_init_set_ui __gmpz_init_set_ui
#define mpz_set         __gmpz_set
#define mpz_set_si      __gmpz_set_si
#define mpz_set_ui      __gmpz_set_ui
#define mpz_set_str     __gmpz_set_str
#define mpz_add         __gmpz_add
#define mpz_add_ui      __gmpz_add_ui
#define mpz_sub         __gmpz_sub
#define mpz_sub_ui      __gmpz_sub_ui
#define mpz_mul         __gmpz_mul
#define mpz_mul_si      __gmpz_mul_si
#define mpz_mul_ui      __gmpz_mul_ui
#define mpz_mul_2exp    __gmpz_mul_2exp
#define mpz_addmul      __gmpz_addmul
#define mpz_submul      __gmpz_submul
#define mpz_neg         __gmpz_neg
#define mpz_abs         __gmpz_abs
#define mpz_swap        __gmpz_swap
#define mpz_cmp         __gmpz_cmp
#define mpz_cmp_si      __gmpz_cmp_si
#define mpz_cmp_ui      __gmpz_cmp_ui
#define mpz_cmpabs      __gmpz_cmpabs
#define mpz_fdiv_q      __gmpz_fdiv_q
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
#define mpz_tdiv_q      __gmpz_tdiv_q
#define mpz_tdiv_r      __gmpz_tdiv_r
#define mpz_tdiv_q_2exp __gmpz_tdiv_q_2exp
#define mpz_divexact    __gmpz_divexact
#define mpz_fdiv_r_ui   __gmpz_fdiv_r_ui
#define mpz_mod_ui      __gmpz_fdiv_r_ui
#define mpz_get_si      __gmpz_get_si
#define mpz_get_ui      __gmpz_get_ui
#define mpz_get_d       __gmpz_get_d
#define mpz_get_str     __gmpz_get_str
#define mpz_sizeinbase  __gmpz_sizeinbase
#define mpz_fits_slong_p __gmpz_fits_slong_p
#define mpz_jacobi      __gmpz_jacobi
#define mpz_legendre    __gmpz_jacobi
#define mpz_invert      __gmpz_invert
#define mpz_nextprime   __gmpz_nextprime

#endif /* TB_ORACLE_SHIM_GMP_H */
