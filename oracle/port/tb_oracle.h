/* oracle/port/tb_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's p-neighbor Hecke hot path
 * (jefferyphein/ternary-birch).  It is the CHECKER for the
 * CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
 * bit-for-bit against outputs of the unmodified reference compiled into
 * oracle/_ref/birch_ref (fixtures under tests/golden/, generator
 * tests/golden/make_golden.py).
 *
 * Integer modes:  precise = 0  int64 with wrap-around arithmetic, the
 *                              reference's Genus<Z64> (compile with -fwrapv);
 *                 precise = 1  __int128, standing in for the reference's GMP
 *                              Genus<Z>.  Exact while every intermediate is
 *                              below 2^127; overflow beyond that is NOT modelled.
 */
#ifndef TB_ORACLE_H
#define TB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    TBO_OK = 0,
    TBO_ERR_OVERFLOW = 1,   /* NeighborManager.h:344-354 discriminant check */
    TBO_ERR_NOT_FOUND = 2,  /* HashMap.h:77 "Key not found." */
    TBO_ERR_BAD_PRIME = 3,  /* Genus.h:334-337 / 343-346 */
    TBO_ERR_NOMEM = 4
};

/* The genus table the reference keeps in Genus<R> (Genus.h:423-434, 12-26). */
typedef struct {
    int64_t N;               /* number of genus representatives */
    int64_t K;               /* number of prime divisors of the level */
    uint64_t seed;           /* Genus::seed_ */
    const int64_t *primes;   /* [K]     Genus::prime_divisors */
    const int64_t *q;        /* [N][6]  rep.q  = (a,b,c,f,g,h) */
    const int64_t *s;        /* [N][9]  rep.s  row-major a11..a33 */
    const int64_t *sinv;     /* [N][9]  rep.sinv */
    const int64_t *scalar;   /* [N]     my_pow(rep.es) */
    const int32_t *lut;      /* [2^K][N] Genus::lut_positions */
    const int64_t *dims;     /* [2^K]   Genus::dims */
} tbo_genus;

typedef struct {
    int64_t dim;
    int64_t nnz;
    int32_t *data;
    int32_t *indices;
    int32_t *indptr;         /* [dim+1] */
} tbo_csr;

/* Per-rep conic base points, consuming the shared mt19937 stream in rep order
 * exactly as one Fp object does across NeighborManager constructions
 * (Fp.h:10-25,165-168; QuadForm.h:847-935; NeighborManager.h:13-48).
 * base[n] = (x0, y0, z0); for p = 2 the three entries are the bit-packed
 * isotropic vectors of QuadForm.h:915-935. */
int tbo_base_points(const tbo_genus *g, uint64_t p, int64_t *base /* [N][3] */);

/* NeighborManager::isotropic_vector(t) (NeighborManager.h:50-144, 370-394),
 * canonical residues. */
void tbo_isotropic_vector(const int64_t q[6], const int64_t base[3], uint64_t p, uint64_t t,
                          int64_t vec[3]);

/* One neighbor: build (NeighborManager.h:207-356) + reduce (QuadForm.h:530-767).
 * Outputs the reduced form and the isometry cur.q -> neighbor. */
int tbo_reduced_neighbor(const int64_t q[6], const int64_t vec[3], uint64_t p, int precise,
                         int64_t out_q[6], int64_t out_s[9]);

/* Loop trips of QuadForm::reduce on the neighbor along `vec` (profiling aid). */
int tbo_reduce_iterations(const int64_t q[6], const int64_t vec[3], uint64_t p);

/* QuadForm::reduce alone, on int64 input (s is updated in place). */
void tbo_reduce(int64_t q[6], int64_t s[9], int precise);

/* The loop body of Genus.h:562-618 for every (n, t): rec[n][t] =
 * vec[3], q[6], s[9], r, spin  (20 int64 per neighbor). */
int tbo_neighbor_records(const tbo_genus *g, uint64_t p, int precise, int64_t *rec);

/* Genus::hecke_matrix_sparse (Genus.h:341-348, 533-672): out[2^K]. */
int tbo_hecke_sparse(const tbo_genus *g, uint64_t p, int precise, tbo_csr *out);
void tbo_csr_free(tbo_csr *m, int64_t count);

/* Genus::_eigenvectors (Genus.h:446-514).  vecs[V][N] are the lifted
 * eigenvectors (Genus::eigenvector, Genus.h:350-390), cond[V] their conductor
 * bitmask, rep[V] the representative index each one is evaluated at
 * (EigenvectorManager::finalize, Eigenvector.h:170-185). */
int tbo_eigenvalues(const tbo_genus *g, uint64_t p, int precise, int64_t V,
                    const int32_t *vecs, const int64_t *cond, const int64_t *rep,
                    int32_t *out /* [V] */);

/* IsometrySequence::next for every (n, t) in order (IsometrySequence.h:45-92):
 * rec = isometry[9], denominator, src, dst  (12 int64 per neighbor). */
int tbo_isometry_sequence(const tbo_genus *g, uint64_t p, int precise, int64_t *rec);

/* QuadForm::hash_value (QuadForm.cpp:779-803). */
uint64_t tbo_hash_form(const int64_t q[6]);

/* Neighbors the last failing call had processed before the error (debug aid). */
int64_t tbo_last_error_position(void);

#ifdef __cplusplus
}
#endif
#endif /* TB_ORACLE_H */
