/* oracle/port/tb_oracle.c -- TEST INFRASTRUCTURE ONLY.  See tb_oracle.h.
 *
 * Plain-C restatement of the reference's p-neighbor Hecke hot path.  Every
 * function cites the reference lines it follows (paths relative to
 * /root/reference/src).  Compile with -fwrapv: the int64 mode reproduces the
 * reference's wrap-around Genus<Z64> arithmetic.
 */
#include "tb_oracle.h"

#include <stdlib.h>
#include <string.h>

#define TBO_REDUCE_CAP 4096

static _Thread_local int64_t tbo_err_pos = -1;
int64_t tbo_last_error_position(void) { return tbo_err_pos; }

/* ------------------------------------------------------------------------ */
/* RNG: std::mt19937(seed) + std::uniform_int_distribution<>(0, p-1), as used */
/* by Fp (Fp.h:12-15, 165-168).  The distribution is libstdc++ 13's Lemire    */
/* downscaling for a 32-bit generator (bits/uniform_int_dist.h, _S_nd).       */
/* ------------------------------------------------------------------------ */
typedef struct { uint32_t mt[624]; int idx; } tbo_rng;

static void tbo_rng_seed(tbo_rng *r, uint64_t seed)
{
    r->mt[0] = (uint32_t)seed;
    for (int i = 1; i < 624; i++)
        r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->idx = 624;
}

static uint32_t tbo_rng_next(tbo_rng *r)
{
    if (r->idx >= 624)
    {
        for (int i = 0; i < 624; i++)
        {
            uint32_t y = (r->mt[i] & 0x80000000u) | (r->mt[(i + 1) % 624] & 0x7fffffffu);
            r->mt[i] = r->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        r->idx = 0;
    }
    uint32_t y = r->mt[r->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* Fp::random (Fp.h:165-168): uniform on [0, p-1]. */
static uint64_t fp_random(tbo_rng *r, uint64_t p)
{
    uint32_t range = (uint32_t)p;
    uint64_t product = (uint64_t)tbo_rng_next(r) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range)
    {
        uint32_t threshold = (uint32_t)(-range) % range;
        while (low < threshold)
        {
            product = (uint64_t)tbo_rng_next(r) * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return product >> 32;
}

/* ------------------------------------------------------------------------ */
/* Fp arithmetic on canonical residues.  The reference keeps values lazily in */
/* [0, kp) (Fp.h:20-21, 52-78); every decision it takes is on a canonical     */
/* value (explicit mod() before each compare), so canonical arithmetic yields */
/* the same residues and the same control flow.                               */
/* ------------------------------------------------------------------------ */
static inline uint64_t fp_mul(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)(((unsigned __int128)a * b) % p); }
static inline uint64_t fp_add(uint64_t a, uint64_t b, uint64_t p) { uint64_t s = a + b; return s >= p ? s - p : s; }
static inline uint64_t fp_sub(uint64_t a, uint64_t b, uint64_t p) { return a >= b ? a - b : a + p - b; }
static inline uint64_t fp_mod_i64(int64_t a, uint64_t p)           /* Fp::mod, Fp.h:29-35 */
{
    int64_t v = a % (int64_t)p;
    return (uint64_t)(v < 0 ? v + (int64_t)p : v);
}

static uint64_t fp_pow(uint64_t a, uint64_t e, uint64_t p)
{
    uint64_t r = 1 % p;
    while (e)
    {
        if (e & 1) r = fp_mul(r, a, p);
        a = fp_mul(a, a, p);
        e >>= 1;
    }
    return r;
}

/* Fp::legendre (Fp.h:90-95, mpz_legendre): Euler's criterion for prime p. */
static int fp_legendre(uint64_t a, uint64_t p)
{
    a %= p;
    if (a == 0) return 0;
    return fp_pow(a, (p - 1) / 2, p) == 1 ? 1 : -1;
}

/* Fp::inverse (Fp.h:147-150, 182-195); F2::inverse (Fp.h:303-306) for p = 2. */
static uint64_t fp_inverse(uint64_t a, uint64_t p)
{
    if (p == 2) return a & 1;
    a %= p;
    if (a == 0) return 0;
    return fp_pow(a, p - 2, p);
}

/* Fp::sqrt, Fp.h:97-145: Tonelli-Shanks with an RNG-drawn non-residue. */
static uint64_t fp_sqrt(tbo_rng *rng, uint64_t a, uint64_t p)
{
    a = a % p;
    if (a == 1) return 1;
    if (a == 0) return 0;
    if (fp_legendre(a, p) != 1) return 0;

    uint64_t q = p - 1, s = 0;
    while (q % 2 == 0) { q >>= 1; s++; }
    if (s == 1) return fp_pow(a, (p + 1) / 4, p);

    uint64_t z = fp_random(rng, p);
    while (z == 0 || fp_legendre(z, p) == 1) z = fp_random(rng, p);

    int m = (int)s;
    uint64_t c = fp_pow(z, q, p);
    uint64_t r = fp_pow(a, (q + 1) / 2, p);
    uint64_t t = fp_pow(a, q, p);
    while (1)
    {
        if (t == 1) return r;
        int i = 0;
        uint64_t t1 = t;
        while (t1 != 1) { t1 = fp_mul(t1, t1, p); i++; }
        uint64_t e = 1;
        for (int j = 0; j < m - i - 1; j++) e <<= 1;
        uint64_t b = fp_pow(c, e, p);
        r = fp_mul(r, b, p);
        c = fp_mul(b, b, p);
        t = fp_mul(t, c, p);
        m = i;
    }
}

/* QuadFormFp::isotropic_vector (QuadForm.h:847-907; p = 2: 915-935), called
 * from the NeighborManager constructor (NeighborManager.h:13-27). */
static void tbo_base_point_one(tbo_rng *rng, const int64_t q[6], uint64_t p, int64_t base[3])
{
    uint64_t a = fp_mod_i64(q[0], p), b = fp_mod_i64(q[1], p), c = fp_mod_i64(q[2], p);
    uint64_t f = fp_mod_i64(q[3], p), g = fp_mod_i64(q[4], p), h = fp_mod_i64(q[5], p);

    if (p == 2)
    {
        int64_t temp[3] = {0, 0, 0};
        int index = 0;
        /* temp[] has room for the three isotropic lines a smooth conic over F2 has */
#define TBO_PUSH(v) do { if (index < 3) temp[index] = (v); index++; } while (0)
        if (c == 0) TBO_PUSH(1);
        if (b == 0) TBO_PUSH(2);
        if ((b ^ f ^ c) == 0) TBO_PUSH(3);
        if (a == 0) TBO_PUSH(4);
        if ((a ^ g ^ c) == 0) TBO_PUSH(5);
        if ((a ^ h ^ b) == 0) TBO_PUSH(6);
        if ((a ^ b ^ c ^ f ^ g ^ h) == 0) TBO_PUSH(7);
#undef TBO_PUSH
        base[0] = temp[0]; base[1] = temp[1]; base[2] = temp[2];
        return;
    }

    while (1)
    {
        uint64_t r = 0, alpha = 0, beta = 0;
        while (alpha == 0)
        {
            r = fp_random(rng, p);
            beta = fp_mul(b, r, p);
            alpha = fp_add(beta, h, p);
            alpha = fp_mul(alpha, r, p);
            alpha = fp_add(alpha, a, p);                          /* Q(1,r,0) */
        }
        uint64_t s = fp_random(rng, p);

        beta = fp_add(beta, beta, p);
        beta = fp_add(beta, h, p);
        beta = fp_mul(beta, s, p);
        beta = fp_add(beta, fp_mul(f, r, p), p);
        beta = fp_add(beta, g, p);                                /* (2br+h)s+fr+g */

        uint64_t gamma = fp_mul(b, s, p);
        gamma = fp_add(gamma, f, p);
        gamma = fp_mul(gamma, s, p);
        gamma = fp_add(gamma, c, p);                              /* Q(0,s,1) */

        uint64_t disc = fp_mul(beta, beta, p);
        gamma = fp_mul(gamma, alpha, p);
        gamma = fp_add(gamma, gamma, p);
        gamma = fp_add(gamma, gamma, p);
        disc = fp_sub(disc, gamma, p);

        if (fp_legendre(disc, p) >= 0)
        {
            uint64_t root = fp_sqrt(rng, disc, p);
            root = fp_sub(root, beta, p);
            alpha = fp_add(alpha, alpha, p);
            uint64_t x = fp_mul(root, fp_inverse(alpha, p), p);
            uint64_t y = fp_add(fp_mul(r, x, p), s, p);
            base[0] = (int64_t)x; base[1] = (int64_t)y; base[2] = 1;
            return;
        }
    }
}

int tbo_base_points(const tbo_genus *g, uint64_t p, int64_t *base)
{
    tbo_rng rng;
    tbo_rng_seed(&rng, g->seed);
    for (int64_t n = 0; n < g->N; n++) tbo_base_point_one(&rng, g->q + 6 * n, p, base + 3 * n);
    return TBO_OK;
}

/* NeighborManager ctor constants (NeighborManager.h:34-47) and
 * isotropic_vector(t) (50-144); p = 2: isotropic_vector_p2 (370-394). */
void tbo_isotropic_vector(const int64_t q[6], const int64_t base[3], uint64_t p, uint64_t t, int64_t vec[3])
{
    if (p == 2)
    {
        int64_t w = base[t == 0 ? 0 : (t == 1 ? 1 : 2)];
        vec[2] = w & 1;
        vec[1] = !!(w & 2);
        vec[0] = !!(w & 4);
        return;
    }

    uint64_t a = fp_mod_i64(q[0], p), b = fp_mod_i64(q[1], p);
    uint64_t f = fp_mod_i64(q[3], p), g = fp_mod_i64(q[4], p), h = fp_mod_i64(q[5], p);
    uint64_t x0 = (uint64_t)base[0] % p, y0 = (uint64_t)base[1] % p;

    uint64_t temp = fp_mul(a, x0, p);
    temp = fp_add(temp, temp, p);
    uint64_t a0 = fp_sub(a, temp, p);
    a0 = fp_sub(a0, g, p);
    a0 = fp_sub(a0, fp_mul(h, y0, p), p);                         /* a-2ax-g-hy */

    temp = fp_mul(b, y0, p);
    temp = fp_add(temp, temp, p);
    uint64_t delta = fp_sub(b, temp, p);
    delta = fp_sub(delta, f, p);
    delta = fp_add(delta, h, p);
    delta = fp_sub(delta, fp_mul(h, x0, p), p);                   /* b-2by-f+h-hx */

    uint64_t rx, ry, rz;
    if (t == p)                                                   /* 58-84 */
    {
        uint64_t at = fp_sub(delta, h, p);
        if (at == 0)
        {
            rx = x0; ry = y0 ? y0 - 1 : p - 1; rz = 1;
        }
        else if (b == 0)
        {
            rx = 0; ry = 1; rz = 0;
        }
        else
        {
            uint64_t inv = fp_mul(fp_inverse(b, p), at, p);
            rx = x0; ry = fp_add(y0, fp_sub(inv, 1, p), p); rz = 1;
        }
    }
    else                                                          /* 85-136 */
    {
        uint64_t at = a0;
        if (t == 1) at = fp_add(at, delta, p);
        else if (t >= 2)
        {
            uint64_t tmp = fp_mul(t - 1, b, p);
            tmp = fp_add(tmp, delta, p);
            tmp = fp_mul(tmp, t, p);
            at = fp_add(tmp, at, p);
        }
        if (at == 0)
        {
            rx = x0 ? x0 - 1 : p - 1; ry = fp_sub(y0, t, p); rz = 1;
        }
        else
        {
            uint64_t ct = fp_mul(b, t, p);
            ct = fp_add(ct, h, p);
            ct = fp_mul(ct, t, p);
            ct = fp_add(ct, a, p);
            if (ct == 0)
            {
                if (t == 0) { rx = 1; ry = 0; rz = 0; }
                else { rx = fp_inverse(t, p); ry = 1; rz = 0; }
            }
            else
            {
                uint64_t inv = fp_mul(at, fp_inverse(ct, p), p);
                rx = fp_add(x0, fp_sub(inv, 1, p), p);
                ry = fp_sub(y0, t, p);
                ry = fp_add(ry, fp_mul(t, inv, p), p);
                rz = 1;
            }
        }
    }
    vec[0] = (int64_t)rx; vec[1] = (int64_t)ry; vec[2] = (int64_t)rz;
}

/* ------------------------------------------------------------------------ */
/* Genus hash table: QuadForm::hash_value (QuadForm.cpp:779-803), HashMap     */
/* (HashMap.h:10-24, 68-89, 107-147).                                         */
/* ------------------------------------------------------------------------ */
uint64_t tbo_hash_form(const int64_t q[6])
{
    uint64_t fnv = 0x811c9dc5ULL;
    for (int i = 0; i < 6; i++) fnv = (fnv ^ (uint64_t)q[i]) * 0x01000193ULL;
    return fnv;
}

typedef struct { int64_t *keyptr; uint64_t mask; } tbo_table;

static int tbo_table_build(tbo_table *tab, const tbo_genus *g)
{
    int lg2 = 4;
    while (((int64_t)1 << lg2) < g->N) lg2++;
    uint64_t slots = (uint64_t)1 << (lg2 + 1);
    tab->mask = slots - 1;
    tab->keyptr = (int64_t *)malloc(sizeof(int64_t) * slots);
    if (!tab->keyptr) return TBO_ERR_NOMEM;
    for (uint64_t i = 0; i < slots; i++) tab->keyptr[i] = -1;
    for (int64_t n = 0; n < g->N; n++)
    {
        uint64_t index = tbo_hash_form(g->q + 6 * n) & tab->mask;
        while (tab->keyptr[index] != -1) index = (index + 1) & tab->mask;
        tab->keyptr[index] = n;
    }
    return TBO_OK;
}

static void tbo_table_free(tbo_table *tab) { free(tab->keyptr); tab->keyptr = NULL; }

static int64_t tbo_table_indexof(const tbo_table *tab, const int64_t *keys, const int64_t key[6])
{
    uint64_t index = tbo_hash_form(key) & tab->mask;
    while (1)
    {
        int64_t offset = tab->keyptr[index];
        if (offset == -1) return -1;
        if (memcmp(keys + 6 * offset, key, 6 * sizeof(int64_t)) == 0) return offset;
        index = (index + 1) & tab->mask;
    }
}

/* ------------------------------------------------------------------------ */
/* The two integer instantiations.                                            */
/* ------------------------------------------------------------------------ */
#define TI int64_t
#define FN(x) x##_i64
#define TBO_CHECK_DISC 1
#include "tb_oracle_int.inc"
#undef TI
#undef FN
#undef TBO_CHECK_DISC

#define TI __int128
#define FN(x) x##_i128
#define TBO_CHECK_DISC 0
#include "tb_oracle_int.inc"
#undef TI
#undef FN
#undef TBO_CHECK_DISC

int tbo_reduced_neighbor(const int64_t q[6], const int64_t vec[3], uint64_t p, int precise,
                         int64_t out_q[6], int64_t out_s[9])
{
    int err = TBO_OK;
    if (precise)
    {
        form_i128 cur = load_form_i128(q), nq;
        iso_i128 s;
        int rc = build_neighbor_i128(&cur, vec, p, &nq, &s);
        if (rc) return rc;
        reduce_i128(&nq, &s, &err);
        store_form_i128(&nq, out_q);
        store_iso_i128(&s, out_s);
    }
    else
    {
        form_i64 cur = load_form_i64(q), nq;
        iso_i64 s;
        int rc = build_neighbor_i64(&cur, vec, p, &nq, &s);
        if (rc) return rc;
        reduce_i64(&nq, &s, &err);
        store_form_i64(&nq, out_q);
        store_iso_i64(&s, out_s);
    }
    return err;
}

/* Loop trips QuadForm::reduce takes on the neighbor along `vec` (profiling aid). */
int tbo_reduce_iterations(const int64_t q[6], const int64_t vec[3], uint64_t p)
{
    int err = TBO_OK;
    form_i128 cur = load_form_i128(q), nq;
    iso_i128 s;
    if (build_neighbor_i128(&cur, vec, p, &nq, &s)) return -1;
    return reduce_i128(&nq, &s, &err);
}

void tbo_reduce(int64_t q[6], int64_t s[9], int precise)
{
    int err = TBO_OK;
    if (precise)
    {
        form_i128 f = load_form_i128(q);
        iso_i128 m = load_iso_i128(s);
        reduce_i128(&f, &m, &err);
        store_form_i128(&f, q);
        store_iso_i128(&m, s);
    }
    else
    {
        form_i64 f = load_form_i64(q);
        iso_i64 m = load_iso_i64(s);
        reduce_i64(&f, &m, &err);
        store_form_i64(&f, q);
        store_iso_i64(&m, s);
    }
}

static int bad_prime(const tbo_genus *g, uint64_t p)
{
    /* Genus.h:334-337: disc % p == 0; the discriminant's primes are g->primes. */
    for (int64_t i = 0; i < g->K; i++) if ((uint64_t)g->primes[i] == p) return 1;
    return 0;
}

int tbo_neighbor_records(const tbo_genus *g, uint64_t p, int precise, int64_t *rec)
{
    if (bad_prime(g, p)) return TBO_ERR_BAD_PRIME;
    return precise ? neighbor_records_i128(g, p, rec) : neighbor_records_i64(g, p, rec);
}

int tbo_hecke_sparse(const tbo_genus *g, uint64_t p, int precise, tbo_csr *out)
{
    if (bad_prime(g, p)) return TBO_ERR_BAD_PRIME;
    return precise ? hecke_sparse_i128(g, p, out) : hecke_sparse_i64(g, p, out);
}

void tbo_csr_free(tbo_csr *m, int64_t count)
{
    for (int64_t k = 0; k < count; k++)
    {
        free(m[k].data); free(m[k].indices); free(m[k].indptr);
        m[k].data = m[k].indices = m[k].indptr = NULL;
    }
}

int tbo_eigenvalues(const tbo_genus *g, uint64_t p, int precise, int64_t V,
                    const int32_t *vecs, const int64_t *cond, const int64_t *rep, int32_t *out)
{
    return precise ? eigenvalues_i128(g, p, V, vecs, cond, rep, out)
                   : eigenvalues_i64(g, p, V, vecs, cond, rep, out);
}

int tbo_isometry_sequence(const tbo_genus *g, uint64_t p, int precise, int64_t *rec)
{
    return precise ? isometry_sequence_i128(g, p, rec) : isometry_sequence_i64(g, p, rec);
}
