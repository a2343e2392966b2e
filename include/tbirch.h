/* include/tbirch.h -- C ABI of libtbirch.so, the B200 (sm_100a) implementation of
 * ternary-birch's p-neighbor Hecke hot path.
 *
 * This is the drop-in boundary: plain C, POD arguments, no torch / C++ types.
 * Each entry point names the reference interface it stands in for (paths are
 * relative to the reference's src/).  The C++ facade in
 * ternary_birch_b200/host/ re-creates the reference's class API
 * (Genus<R>::hecke_matrix_sparse, ...) on top of these calls, and INTEGRATION.md
 * shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - Every function returning int returns a tb_status.  On failure the
 *     calling thread's tb_last_error() holds the SAME what() string the
 *     reference's C++ exception carries, and tb_status says which exception
 *     type the facade must rethrow.
 *   - `precise` = 0: int64 wrap-around arithmetic (the reference's Genus<Z64>,
 *     pyx precise=False).  `precise` = 1: overflow-checked 128-bit arithmetic
 *     replacing GMP (the reference's Genus<Z>, pyx precise=True); an operation
 *     that would exceed 127 bits reports TB_ERR_OVERFLOW.
 *   - There is no CPU fallback: without a usable CUDA device every compute
 *     entry point fails with TB_ERR_RUNTIME.
 *   - Destination pointers of the tb_*_copy_* calls may be host OR device
 *     memory (unified virtual addressing decides).
 *   - A tb_genus and the objects made from it must be used by one host thread
 *     at a time (the reference's Genus has the same restriction).
 */
#ifndef TBIRCH_H
#define TBIRCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TBIRCH_ABI_VERSION 2

typedef enum tb_status {
    TB_OK = 0,
    TB_ERR_INVALID_ARGUMENT = 1,  /* std::invalid_argument: "Prime must not divide the discriminant." (Genus.h:336,345), "Key not found." (HashMap.h:77), "Invalid conductor." / "Eigenvector has incorrect dimension." (Genus.h:367,373) */
    TB_ERR_OVERFLOW = 2,          /* std::overflow_error: NeighborManager.h:350-352 */
    TB_ERR_DOMAIN = 3,            /* std::domain_error: "No more isometries." (IsometrySequence.h:52), "Must have 63 or fewer prime divisors." (Genus.h:59) */
    TB_ERR_RUNTIME = 4            /* std::runtime_error: CUDA failures, missing device, unsupported sizes */
} tb_status;

typedef struct tb_genus tb_genus;
typedef struct tb_hecke tb_hecke;
typedef struct tb_isoseq tb_isoseq;

/* The state Genus<R> keeps after construction (Genus.h:423-434; GenusRep,
 * Genus.h:12-26), flattened.  All arrays are host memory, copied by
 * tb_genus_create; the caller may free them afterwards. */
typedef struct tb_genus_desc {
    int64_t N;                  /* number of genus representatives (Genus::size) */
    int64_t K;                  /* number of prime divisors; 2^K conductors */
    uint64_t seed;              /* Genus::seed_ (orders the isotropic lines, Fp.h:12-15) */
    int64_t disc;               /* Genus::disc */
    const int64_t *primes;      /* [K]      Genus::prime_divisors */
    const int64_t *conductors;  /* [2^K]    Genus::conductors */
    const int64_t *dims;        /* [2^K]    Genus::dims */
    const int64_t *q;           /* [N][6]   GenusRep::q = (a,b,c,f,g,h) */
    const int64_t *s;           /* [N][9]   GenusRep::s    (row-major a11..a33) */
    const int64_t *sinv;        /* [N][9]   GenusRep::sinv */
    const int64_t *scalar;      /* [N]      birch_util::my_pow(GenusRep::es) */
    const int32_t *lut;         /* [2^K][N] Genus::lut_positions */
} tb_genus_desc;

/* ---- library ---------------------------------------------------------- */
int tb_abi_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char *tb_last_error(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's
 * "gpu_launches" claim). */
uint64_t tb_launch_count(void);
/* Device time in ms of the last neighbor-kernel launch sequence of `g`
 * (CUDA events on g's stream), and the number of neighbors it covered. */
int tb_last_kernel_time(tb_genus *g, float *ms_neighbors, float *ms_rows, int64_t *neighbors);
/* Integer mode of g's last neighbor-kernel launch (diagnostic; results never depend on it):
 * 0 = wrap-around int64 (precise = 0, Genus<Z64>); precise = 1 (Genus<Z>) runs as
 * 1 = 64-bit under host-proven magnitude bounds, 2 = unchecked 128-bit under proven bounds,
 * 3 = overflow-checked 128-bit.  -1 before the first launch. */
int tb_last_int_mode(const tb_genus *g);

/* ---- genus table ------------------------------------------------------ */
/* Replaces the state built by Genus<R>::Genus (Genus.h:37-246) and copied by the
 * converting constructor Genus<Z64>(const Genus<Z>&) (Genus.h:248-303): uploads
 * the table, builds the device hash table of representatives
 * (HashMap.h:107-147, keyed by QuadForm::hash_value, QuadForm.cpp:779-803). */
int tb_genus_create(const tb_genus_desc *desc, tb_genus **out);
void tb_genus_destroy(tb_genus *g);
/* Run all of g's work on this cudaStream_t (default: the legacy default stream). */
int tb_genus_set_stream(tb_genus *g, void *cuda_stream);
/* Re-upload the table into the existing device buffers (same N, K). */
int tb_genus_upload(tb_genus *g, const tb_genus_desc *desc);
int64_t tb_genus_size(const tb_genus *g);            /* Genus::size, Genus.h:306 */
int64_t tb_genus_num_conductors(const tb_genus *g);
int64_t tb_genus_conductor(const tb_genus *g, int64_t k);
int64_t tb_genus_dim(const tb_genus *g, int64_t k);  /* Genus::dimension_map, Genus.h:321-330 */

/* ---- Hecke matrices --------------------------------------------------- */
/* Genus<R>::hecke_matrix_sparse / hecke_matrix_dense (Genus.h:332-348) and the
 * loops behind them (533-672, 674-848): T_p for ALL 2^K conductors in one pass.
 * The matrices stay on the device inside *out until copied. `rep_begin/rep_end`
 * restrict the pass to rows of representatives [rep_begin, rep_end) -- the
 * multi-GPU shard; pass 0, N for the whole matrix. */
int tb_hecke_compute(tb_genus *g, uint64_t p, int precise, tb_hecke **out);
int tb_hecke_compute_rows(tb_genus *g, uint64_t p, int precise,
                          int64_t rep_begin, int64_t rep_end, tb_hecke **out);
int64_t tb_hecke_num_conductors(const tb_hecke *h);
int64_t tb_hecke_conductor(const tb_hecke *h, int64_t k);
int64_t tb_hecke_dim(const tb_hecke *h, int64_t k);   /* columns; full rows of T^(k) */
int64_t tb_hecke_rows(const tb_hecke *h, int64_t k);  /* rows present in this shard */
int64_t tb_hecke_row_begin(const tb_hecke *h, int64_t k);
int64_t tb_hecke_nnz(const tb_hecke *h, int64_t k);
/* CSR triple in the reference's order (data, indices, indptr), Genus.h:662-671:
 * data[nnz], indices[nnz], indptr[rows+1]. */
int tb_hecke_copy_sparse(tb_hecke *h, int64_t k, int32_t *data, int32_t *indices, int32_t *indptr);
/* All conductors in one call, into three concatenated arrays: data_all / indices_all hold the
 * tb_hecke_nnz(h, k) entries of conductor 0, 1, ... back to back, indptr_all the
 * (tb_hecke_rows(h, k) + 1)-blocks.  Same values as 2^K tb_hecke_copy_sparse calls, a fraction
 * of their fixed costs. */
int tb_hecke_copy_sparse_all(tb_hecke *h, int32_t *data_all, int32_t *indices_all, int32_t *indptr_all);
/* tb_hecke_copy_sparse without synchronising: lets the caller overlap the read-back of one T_p with the
 * next one.  Device destinations and small matrices are enqueued on `cuda_stream`; large host
 * destinations are served by the library's background read-back.  Either way
 * tb_hecke_copy_wait(h) is the completion point. */
int tb_hecke_copy_sparse_async(tb_hecke *h, int64_t k, int32_t *data, int32_t *indices, int32_t *indptr,
                               void *cuda_stream);
/* Blocks until every asynchronous copy of `h` issued so far has landed: the event of the last
 * stream-ordered copy, and the background read-back of large host copies (which runs on the
 * library's own copy stream and host threads, not on the caller's stream). */
int tb_hecke_copy_wait(tb_hecke *h);
/* Bytes that cross PCIe when data + indices + indptr of conductor k are read back to host
 * memory (diagnostic for bench.py's d2h_bytes_per_step).  Host read-backs of large matrices go
 * through a 3-byte-per-entry device format (uint16 column, int8 entry) that the library widens
 * on the host; the arrays the caller receives are the reference's int32 CSR either way. */
int64_t tb_hecke_transfer_bytes(const tb_hecke *h, int64_t k);
/* Which of the two read-back formats large host copies use is decided per rank from measured
 * delivery rates (profiles/r2a_d2h_probe_n4.md: what a rank gets depends on the host and on the
 * other ranks): running GB/s of delivered int32 CSR for the plain DMA and for the 3-byte format,
 * the number of large read-backs behind each figure, and the format currently preferred
 * (0 plain, 1 narrow).  TBIRCH_NO_NARROW_COPY / TBIRCH_FORCE_NARROW_COPY pin the choice. */
int tb_readback_stats(double *plain_gbps, double *narrow_gbps, int *plain_samples, int *narrow_samples,
                      int *current);
/* Host-side widening throughput of this machine (no GPU involved; diagnostic): output GB/s of the
 * library's thread pool on `entries` packed entries, the pool size, and whether the AVX2
 * streaming-store path is in use. */
int tb_host_widen_probe(int64_t entries, double *out_gbps, int *threads, int *avx2);

/* ---- final integer gather of the rows (multi-GPU) ---------- */
/* The path needs no collective until the end: rank i holds the rows (or the primes) it computed.
 * tb_hecke_pack lays ALL conductors of `h` out in one device buffer of tb_hecke_packed_bytes(h)
 * bytes -- per conductor: columns[nnz] | entries[nnz] | indptr[rows + 1], each section padded to
 * 16 bytes; *entry_bytes = 3: uint16 columns + int8 entries, 8: int32 + int32 (a column or an
 * entry does not fit) -- which the caller ships to the root with grouped send/recv
 * (ternary_birch_b200/shard.py does it over torch.distributed/NCCL, unpadded).  On the root,
 * tb_unpack_rows widens one such buffer into the reference's int32 CSR on the device, conductors
 * back to back (the layout of tb_hecke_copy_sparse_all).  Both enqueue on `cuda_stream`. */
int64_t tb_hecke_packed_bytes(const tb_hecke *h, int *entry_bytes);
int tb_hecke_pack(tb_hecke *h, void *dst_device, void *cuda_stream);
int tb_unpack_rows(const void *packed_device, int entry_bytes, int64_t num_conductors, const int64_t *nnz,
                   const int64_t *rows, int32_t *data_all, int32_t *indices_all, int32_t *indptr_all,
                   void *cuda_stream);

/* Row-major rows x dim block, Genus.h:811-846. */
int tb_hecke_copy_dense(tb_hecke *h, int64_t k, int32_t *matrix);
void tb_hecke_destroy(tb_hecke *h);

/* ---- eigenvalues ------------------------------------------------------ */
/* Genus<R>::eigenvalues (Genus.h:392-421) / _eigenvectors (446-514).
 *   vecs[V][N]   eigenvectors lifted to the full genus (Genus::eigenvector, 350-390)
 *   cond_mask[V] conductor index k of each vector (Eigenvector::conductor_index)
 *   rep_index[V] representative each vector is evaluated at
 *                (EigenvectorManager::finalize, Eigenvector.h:170-185)
 *   out[V]       the eigenvalues, in vector order. */
int tb_eigenvalues(tb_genus *g, uint64_t p, int precise, int64_t V, const int32_t *vecs,
                   const int64_t *cond_mask, const int64_t *rep_index, int32_t *out);

/* ---- isometry sequence ------------------------------------------------ */
/* IsometrySequence<R,S,T> (IsometrySequence.h:19-92): rep-major, t-minor stream
 * of (isometry, denominator, src, dst).  tb_isoseq_next past the end fails with
 * TB_ERR_DOMAIN "No more isometries.". */
int tb_isoseq_open(tb_genus *g, uint64_t p, int precise, tb_isoseq **out);
int tb_isoseq_done(const tb_isoseq *it);
int tb_isoseq_next(tb_isoseq *it, int64_t isometry[9], int64_t *denominator, int64_t *src, int64_t *dst);
/* Bulk form: up to max_records records of 12 int64 (isometry[9], denominator,
 * src, dst); returns the count through *got. */
int tb_isoseq_read(tb_isoseq *it, int64_t max_records, int64_t *records, int64_t *got);
void tb_isoseq_close(tb_isoseq *it);

/* ---- per-neighbor records (parity tests) ------------------------------ */
/* The loop body of Genus.h:567-618 for every (n, t) with n in
 * [rep_begin, rep_end): 20 int64 per neighbor = isotropic vector[3], reduced
 * form[6], isometry rep -> neighbor[9], r, spin.  `records` is host or device. */
int tb_neighbor_records(tb_genus *g, uint64_t p, int precise, int64_t rep_begin, int64_t rep_end,
                        int64_t *records);

/* ---- genus construction (the caller side of the hot path) -- */
typedef struct tb_prime_symbol {   /* PrimeSymbol<R>, birch.h:107-112 */
    int64_t p;
    int32_t power;
    int32_t ramified;
} tb_prime_symbol;

typedef struct tb_table tb_table;

/* QuadForm<Z>::get_quad_form (QuadForm.cpp:444-777): the positive definite ternary form
 * (a,b,c,f,g,h) with the given discriminant prod p^power and ramification. */
int tb_quad_form(const tb_prime_symbol *symbols, int64_t count, int64_t form[6]);
/* Genus<R>::Genus(q, symbols, seed) (Genus.h:37-246): enumerate the genus of `form` by
 * p-neighbor walking in the reference's discovery order (seed = 0 draws a random seed, as
 * the reference does), compose the isometries to the mother form (189-217) and derive
 * dims / lut_positions from the automorphisms' spinor norms (219-244). */
int tb_genus_enumerate(const int64_t form[6], const tb_prime_symbol *symbols, int64_t count,
                       uint64_t seed, tb_table **out);
/* Fill `desc` with pointers into the table (valid until tb_table_destroy); the extra arrays
 * are GenusRep::parent, GenusRep::p and QuadForm::num_automorphisms per representative. */
int tb_table_desc(const tb_table *t, tb_genus_desc *desc, const int64_t **parent,
                  const int64_t **rep_prime, const int64_t **num_auts);
void tb_table_destroy(tb_table *t);

/* ---- modular linear algebra for the eigenvector search ----------------- */
/* The consumer of T_p: the reference's
 * BirchGenus.rational_eigenvectors (ternary_birch.pyx:246-379) asks Sage for
 * (matrix - e).right_kernel(); these compute that kernel modulo a prime P < 2^29 on the
 * GPU (panel-blocked Gauss-Jordan elimination); the Python layer lifts it to Q (CRT,
 * rational reconstruction) and verifies it exactly.  The basis is the canonical one of the
 * reduced row echelon form: one vector per free column f (ascending), with x[f] = 1 and 0 in
 * the other free columns, so bases for different P agree entrywise modulo the P's. */
typedef struct tb_modkernel tb_modkernel;
/* right kernel {x : (A - shift I) x = 0 mod P} of an n x n CSR int32 matrix (host pointers) */
int tb_modkernel_csr(const int32_t *data, const int32_t *indices, const int32_t *indptr, int64_t n,
                     int64_t shift, uint32_t P, tb_modkernel **out);
/* right kernel {x : M x = 0 mod P} of a dense rows x cols row-major int64 matrix (host pointer) */
int tb_modkernel_dense(const int64_t *M, int64_t rows, int64_t cols, uint32_t P, tb_modkernel **out);
int64_t tb_modkernel_dim(const tb_modkernel *k);
int64_t tb_modkernel_cols(const tb_modkernel *k);
/* free (non-pivot) columns, ascending: free_cols[dim] */
int tb_modkernel_free_cols(const tb_modkernel *k, int64_t *free_cols);
/* basis[dim][cols], residues in [0, P) */
int tb_modkernel_basis(const tb_modkernel *k, uint32_t *basis);
/* device time of the elimination in ms (CUDA events) */
float tb_modkernel_ms(const tb_modkernel *k);
void tb_modkernel_destroy(tb_modkernel *k);

/* ---- integer-pipe microbenchmark -------------------------------------- */
/* Measures the int32 issue peak used as the roofline denominator: unrolled chains of single
 * SASS instructions on every SM, counted per instruction.  Returns int32 op/s for
 * out[0] IMAD only (fma pipe), out[1] LOP3 only (alu pipe), out[2] ADD only (ptxas splits them
 * over both pipes), out[3] IMAD + ADD alternating, out[4] IMAD + LOP3 alternating, out[5] DFMA
 * only (fp64 pipe), out[6] DFMA + IMAD + LOP3 in turn (instructions/s, all three pipes).  The nominal
 * bound of the multi-pipe arms is the issue slot: SMs x 4 x 32 x clock. */
int tb_int_peak_detail(double out[7]);
/* The same, condensed: fma-pipe-only, alu-pipe-only, and the best two-pipe arm. */
int tb_int_peak(double *imad_ops_per_s, double *alu_ops_per_s, double *mixed_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* TBIRCH_H */
