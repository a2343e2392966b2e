// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// Command-line driver around the UNMODIFIED reference sources, which are
// compiled from where they lie under /root/reference/src by oracle/Makefile
// into oracle/_ref/birch_ref.  It builds a genus the way the reference's Python
// binding does (ternary_birch.pyx:139-205: factor the level, default
// ramification, QuadForm::get_quad_form, Genus<Z>(q, symbols, seed)) and then
// runs the reference's own entry points, dumping their results as flat
// little-endian int64 streams that tests/ and bench.py read with numpy.
//
//   birch_ref --level L [--ramified p,q,..] [--seed S] <command>...
//
// commands (any number, executed in order on the same genus):
//   genus  OUT                    genus table (format: see write_genus below)
//   hecke  P MODE FMT OUT         Genus::hecke_matrix_{sparse,dense}(P); MODE = z64|z, FMT = sparse|dense
//   spin   P MODE OUT             per-neighbor records of the loop at Genus.h:562-618
//   eigen  VECS PRIMES MODE OUT   Genus::eigenvalues for every prime in the comma list
//   isoseq P MODE OUT             IsometrySequence<W16,W32,T> stream (IsometrySequence.h:45-92)
//   time   P MODE FMT REPEAT      time the reference call, print one JSON line per call
//
// The only non-API access is `#define private public`, used to READ the genus
// table (Genus::hash, lut_positions, ...) that the reference keeps private.

#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <iterator>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <gmpxx.h>

#define private public
#define protected public
#include "birch.h"
#include "Genus.h"
#include "IsometrySequence.h"
#undef private
#undef protected

typedef std::vector<int64_t> Stream;

static void write_stream(const std::string& path, const Stream& out)
{
    FILE *fp = std::fopen(path.c_str(), "wb");
    if (!fp) { std::perror(path.c_str()); std::exit(2); }
    if (!out.empty() && std::fwrite(out.data(), sizeof(int64_t), out.size(), fp) != out.size())
    {
        std::perror("fwrite"); std::exit(2);
    }
    std::fclose(fp);
}

static Stream read_stream(const std::string& path)
{
    FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp) { std::perror(path.c_str()); std::exit(2); }
    std::fseek(fp, 0, SEEK_END);
    long bytes = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    Stream in(bytes / sizeof(int64_t));
    if (!in.empty() && std::fread(in.data(), sizeof(int64_t), in.size(), fp) != in.size())
    {
        std::perror("fread"); std::exit(2);
    }
    std::fclose(fp);
    return in;
}

static int64_t to_i64(const Z& x)
{
    if (!mpz_fits_slong_p(x.get_mpz_t()))
    {
        std::fprintf(stderr, "birch_ref: value %s does not fit int64\n", x.get_str().c_str());
        std::exit(3);
    }
    return mpz_get_si(x.get_mpz_t());
}
static int64_t to_i64(const Z64& x) { return x; }

static std::vector<int64_t> parse_list(const std::string& s)
{
    std::vector<int64_t> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ',')) if (!item.empty()) out.push_back(std::stoll(item));
    return out;
}

template<typename R>
static void push_iso(Stream& out, const Isometry<R>& s)
{
    out.push_back(to_i64(s.a11)); out.push_back(to_i64(s.a12)); out.push_back(to_i64(s.a13));
    out.push_back(to_i64(s.a21)); out.push_back(to_i64(s.a22)); out.push_back(to_i64(s.a23));
    out.push_back(to_i64(s.a31)); out.push_back(to_i64(s.a32)); out.push_back(to_i64(s.a33));
}

template<typename R>
static void push_form(Stream& out, const QuadForm<R>& q)
{
    out.push_back(to_i64(q.a())); out.push_back(to_i64(q.b())); out.push_back(to_i64(q.c()));
    out.push_back(to_i64(q.f())); out.push_back(to_i64(q.g())); out.push_back(to_i64(q.h()));
}

// Genus table layout (all int64):
//   magic 0x3153554e45474254 ("TBGENUS1"), N, K, seed, disc,
//   prime_divisors[K], conductors[2^K], dims[2^K],
//   q[N][6], s[N][9], sinv[N][9], scalar[N] (= my_pow(es)), parent[N], p[N],
//   num_auts[N] (QuadForm::num_automorphisms), lut_positions[2^K][N]
static void write_genus(const Z_Genus& g, const std::string& path)
{
    Stream out;
    size_t N = g.size();
    size_t K = g.prime_divisors.size();
    size_t C = g.conductors.size();
    out.push_back(0x3153554e45474254LL);
    out.push_back((int64_t)N);
    out.push_back((int64_t)K);
    out.push_back((int64_t)g.seed());
    out.push_back(to_i64(g.disc));
    for (const Z& p : g.prime_divisors) out.push_back(to_i64(p));
    for (const Z& c : g.conductors) out.push_back(to_i64(c));
    for (size_t d : g.dims) out.push_back((int64_t)d);
    for (size_t n = 0; n < N; n++) push_form(out, g.hash->get(n).q);
    for (size_t n = 0; n < N; n++) push_iso(out, g.hash->get(n).s);
    for (size_t n = 0; n < N; n++) push_iso(out, g.hash->get(n).sinv);
    for (size_t n = 0; n < N; n++) out.push_back(to_i64(birch_util::my_pow(g.hash->get(n).es)));
    for (size_t n = 0; n < N; n++) out.push_back(g.hash->get(n).parent);
    for (size_t n = 0; n < N; n++) out.push_back(to_i64(g.hash->get(n).p));
    for (size_t n = 0; n < N; n++) out.push_back(Z_QuadForm::num_automorphisms(g.hash->get(n).q));
    for (size_t k = 0; k < C; k++)
        for (size_t n = 0; n < N; n++) out.push_back(g.lut_positions[k][n]);
    write_stream(path, out);
}

// Layout: C, then for each conductor index k in 0..C-1:
//   sparse: conductor, dim, nnz, data[nnz], indices[nnz], indptr[dim+1]
//   dense:  conductor, dim, matrix[dim*dim]
template<typename R>
static Stream run_hecke(const Genus<R>& g, int64_t p, bool sparse)
{
    Stream out;
    size_t C = g.conductors.size();
    out.push_back((int64_t)C);
    R prime = birch_util::convert_Integer<Z64,R>(p);
    if (sparse)
    {
        std::map<R,std::vector<std::vector<int>>> csr = g.hecke_matrix_sparse(prime);
        for (size_t k = 0; k < C; k++)
        {
            const std::vector<std::vector<int>>& m = csr[g.conductors[k]];
            out.push_back(to_i64(g.conductors[k]));
            out.push_back((int64_t)g.dims[k]);
            out.push_back((int64_t)m[0].size());
            for (int x : m[0]) out.push_back(x);
            for (int x : m[1]) out.push_back(x);
            for (int x : m[2]) out.push_back(x);
        }
    }
    else
    {
        std::map<R,std::vector<int>> dense = g.hecke_matrix_dense(prime);
        for (size_t k = 0; k < C; k++)
        {
            const std::vector<int>& m = dense[g.conductors[k]];
            out.push_back(to_i64(g.conductors[k]));
            out.push_back((int64_t)g.dims[k]);
            for (int x : m) out.push_back(x);
        }
    }
    return out;
}

// Per-neighbor records, following the loop body of hecke_matrix_sparse_internal
// (Genus.h:562-618) call for call.  Layout: N, p, then N*(p+1) records of
//   vec[3] (isotropic vector, canonical mod p), q[6] (reduced neighbor),
//   s[9] (cur.q -> neighbor isometry after reduce), r, spin.
template<typename R>
static Stream run_spin(const Genus<R>& g, int64_t p)
{
    Stream out;
    size_t N = g.size();
    out.push_back((int64_t)N);
    out.push_back(p);
    W16 prime = (W16)p;
    R pR = birch_util::convert_Integer<Z64,R>(p);
    std::shared_ptr<W16_Fp> GF;
    if (prime == 2) GF = std::make_shared<W16_F2>(2, g.seed());
    else GF = std::make_shared<W16_Fp>(prime, g.seed(), true);
    const GenusRep<R>& mother = g.hash->keys()[0];
    for (size_t n = 0; n < N; n++)
    {
        const GenusRep<R>& cur = g.hash->get(n);
        NeighborManager<W16,W32,R> manager(cur.q, GF);
        for (W32 t = 0; t <= prime; t++)
        {
            W16_Vector3 vec = manager.isotropic_vector((W16)t);
            GenusRep<R> foo = manager.get_reduced_neighbor_rep((W16)t);
            size_t r = g.hash->indexof(foo);
            Isometry<R> nbr_s = foo.s;
            W64 spin_vals;
            if (r == n)
            {
                spin_vals = g.spinor->norm(foo.q, foo.s, pR);
            }
            else
            {
                const GenusRep<R>& rep = g.hash->get(r);
                foo.s = cur.s * foo.s;
                R scalar = pR;
                foo.s = foo.s * rep.sinv;
                scalar *= birch_util::my_pow(cur.es);
                scalar *= birch_util::my_pow(rep.es);
                spin_vals = g.spinor->norm(mother.q, foo.s, scalar);
            }
            if (prime == 2)
            {
                out.push_back(vec.x); out.push_back(vec.y); out.push_back(vec.z);
            }
            else
            {
                out.push_back(GF->mod(vec.x)); out.push_back(GF->mod(vec.y)); out.push_back(GF->mod(vec.z));
            }
            push_form(out, foo.q);
            push_iso(out, nbr_s);
            out.push_back((int64_t)r);
            out.push_back((int64_t)spin_vals);
        }
    }
    return out;
}

// Eigenvector file layout (int64): V, then per vector: conductor, dim, coords[dim].
// Output layout: V, P, then for each prime: p, aps[V].
template<typename R>
static Stream run_eigen(const Genus<R>& g, const Stream& vecs, const std::vector<int64_t>& primes)
{
    EigenvectorManager<R> manager;
    size_t pos = 0;
    int64_t V = vecs.at(pos++);
    for (int64_t v = 0; v < V; v++)
    {
        int64_t cond = vecs.at(pos++);
        int64_t dim = vecs.at(pos++);
        std::vector<Z32> data(dim);
        for (int64_t i = 0; i < dim; i++) data[i] = (Z32)vecs.at(pos++);
        manager.add_eigenvector(g.eigenvector(data, birch_util::convert_Integer<Z64,R>(cond)));
    }
    manager.finalize();

    Stream out;
    out.push_back(V);
    out.push_back((int64_t)primes.size());
    for (int64_t p : primes)
    {
        std::vector<Z32> aps = g.eigenvalues(manager, birch_util::convert_Integer<Z64,R>(p));
        out.push_back(p);
        for (Z32 a : aps) out.push_back(a);
    }
    return out;
}

// Layout: count, then per record: isometry[9], denominator, src, dst.
template<typename R>
static Stream run_isoseq(std::shared_ptr<Genus<R>> g, int64_t p)
{
    Stream out;
    out.push_back(0);
    IsometrySequence<W16,W32,R> seq(g, birch_util::convert_Integer<Z64,R>(p));
    int64_t count = 0;
    while (!seq.done())
    {
        IsometrySequenceData<R> d = seq.next();
        push_iso(out, d.isometry);
        out.push_back(to_i64(d.denominator));
        out.push_back((int64_t)d.src);
        out.push_back((int64_t)d.dst);
        ++count;
    }
    out[0] = count;
    return out;
}

template<typename R>
static void run_time(const Genus<R>& g, int64_t p, bool sparse, int repeat, const char *mode)
{
    R prime = birch_util::convert_Integer<Z64,R>(p);
    for (int it = 0; it < repeat; it++)
    {
        auto t0 = std::chrono::steady_clock::now();
        size_t sink = 0;
        if (sparse)
        {
            auto m = g.hecke_matrix_sparse(prime);
            sink = m.size();
        }
        else
        {
            auto m = g.hecke_matrix_dense(prime);
            sink = m.size();
        }
        auto t1 = std::chrono::steady_clock::now();
        double sec = std::chrono::duration<double>(t1 - t0).count();
        std::printf("{\"p\": %lld, \"mode\": \"%s\", \"format\": \"%s\", \"reps\": %zu, \"neighbors\": %lld, "
                    "\"seconds\": %.6f, \"conductors\": %zu}\n",
                    (long long)p, mode, sparse ? "sparse" : "dense", g.size(),
                    (long long)(g.size() * (size_t)(p + 1)), sec, sink);
        std::fflush(stdout);
    }
}

static std::vector<Z_PrimeSymbol> make_symbols(int64_t level, const std::vector<int64_t>& ramified, bool have_ramified)
{
    // ternary_birch.pyx:146-167: factor the level; when no ramification is
    // given, ramify all primes if their number is odd, else all but the largest.
    std::vector<std::pair<int64_t,int>> facs;
    int64_t m = level;
    for (int64_t d = 2; d * d <= m; d++)
    {
        int e = 0;
        while (m % d == 0) { m /= d; ++e; }
        if (e) facs.push_back(std::make_pair(d, e));
    }
    if (m > 1) facs.push_back(std::make_pair(m, 1));

    std::vector<int64_t> ram;
    if (!have_ramified)
    {
        for (auto& pe : facs) ram.push_back(pe.first);
        if (facs.size() % 2 == 0) ram.pop_back();
    }
    else
    {
        for (int64_t p : ramified)
            for (auto& pe : facs) if (pe.first == p) ram.push_back(p);
    }

    std::vector<Z_PrimeSymbol> symbols;
    for (auto& pe : facs)
    {
        Z_PrimeSymbol symb;
        symb.p = Z(pe.first);
        symb.power = pe.second;
        symb.ramified = std::find(ram.begin(), ram.end(), pe.first) != ram.end();
        symbols.push_back(symb);
    }
    return symbols;
}

int main(int argc, char **argv)
{
    int64_t level = 0;
    W64 seed = 1;
    std::vector<int64_t> ramified;
    bool have_ramified = false;

    int i = 1;
    for (; i < argc; i++)
    {
        std::string a = argv[i];
        if (a == "--level" && i + 1 < argc) level = std::stoll(argv[++i]);
        else if (a == "--seed" && i + 1 < argc) seed = std::stoull(argv[++i]);
        else if (a == "--ramified" && i + 1 < argc) { ramified = parse_list(argv[++i]); have_ramified = true; }
        else break;
    }
    if (level <= 0)
    {
        std::fprintf(stderr, "usage: birch_ref --level L [--ramified p,..] [--seed S] <command>...\n");
        return 2;
    }

    try
    {
        std::vector<Z_PrimeSymbol> symbols = make_symbols(level, ramified, have_ramified);
        Z_QuadForm q = Z_QuadForm::get_quad_form(symbols);
        std::shared_ptr<Z_Genus> gz = std::make_shared<Z_Genus>(q, symbols, seed);
        std::shared_ptr<Z64_Genus> g64;
        auto need64 = [&]() { if (!g64) g64 = std::make_shared<Z64_Genus>(*gz); };

        while (i < argc)
        {
            std::string cmd = argv[i++];
            auto arg = [&]() -> std::string {
                if (i >= argc) { std::fprintf(stderr, "birch_ref: missing argument for %s\n", cmd.c_str()); std::exit(2); }
                return argv[i++];
            };
            if (cmd == "genus")
            {
                write_genus(*gz, arg());
            }
            else if (cmd == "hecke")
            {
                int64_t p = std::stoll(arg()); std::string mode = arg(); std::string fmt = arg(); std::string out = arg();
                bool sparse = (fmt == "sparse");
                if (mode == "z64") { need64(); write_stream(out, run_hecke(*g64, p, sparse)); }
                else write_stream(out, run_hecke(*gz, p, sparse));
            }
            else if (cmd == "spin")
            {
                int64_t p = std::stoll(arg()); std::string mode = arg(); std::string out = arg();
                if (mode == "z64") { need64(); write_stream(out, run_spin(*g64, p)); }
                else write_stream(out, run_spin(*gz, p));
            }
            else if (cmd == "eigen")
            {
                Stream vecs = read_stream(arg()); std::vector<int64_t> primes = parse_list(arg());
                std::string mode = arg(); std::string out = arg();
                if (mode == "z64") { need64(); write_stream(out, run_eigen(*g64, vecs, primes)); }
                else write_stream(out, run_eigen(*gz, vecs, primes));
            }
            else if (cmd == "isoseq")
            {
                int64_t p = std::stoll(arg()); std::string mode = arg(); std::string out = arg();
                if (mode == "z64") { need64(); write_stream(out, run_isoseq(g64, p)); }
                else write_stream(out, run_isoseq(gz, p));
            }
            else if (cmd == "time")
            {
                int64_t p = std::stoll(arg()); std::string mode = arg(); std::string fmt = arg(); int repeat = std::stoi(arg());
                bool sparse = (fmt == "sparse");
                if (mode == "z64") { need64(); run_time(*g64, p, sparse, repeat, "z64"); }
                else run_time(*gz, p, sparse, repeat, "z");
            }
            else
            {
                std::fprintf(stderr, "birch_ref: unknown command %s\n", cmd.c_str());
                return 2;
            }
        }
    }
    catch (const std::overflow_error& e)
    {
        std::printf("{\"error\": \"overflow_error\", \"what\": \"%s\"}\n", e.what());
        return 10;
    }
    catch (const std::invalid_argument& e)
    {
        std::printf("{\"error\": \"invalid_argument\", \"what\": \"%s\"}\n", e.what());
        return 11;
    }
    catch (const std::exception& e)
    {
        std::printf("{\"error\": \"exception\", \"what\": \"%s\"}\n", e.what());
        return 12;
    }
    return 0;
}
