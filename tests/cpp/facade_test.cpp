// tests/cpp/facade_test.cpp -- exercises the C++ facade headers (ternary_birch_b200/host)
// the way the reference's binding and demo use the reference's classes.
//
//   facade_test <genus stream file (raw int64)> <p> <eigenvector file or -> <out file>
//
// Writes an int64 stream: for R = int64_t (precise=False) and R = Big (a stand-in
// arbitrary-precision type, precise=True): C, then per conductor (cond, dim, nnz,
// data, indices, indptr) from hecke_matrix_sparse, the dense matrices' checksums,
// and the first isometry-sequence records.
#include <cstdio>
#include <fstream>
#include <iostream>

#ifdef TB_FACADE_GMPXX
#include <gmpxx.h>      // -I oracle/shim: a plain mpz_class over libgmp.so (this image has no GMP headers)
#endif

#include "../../ternary_birch_b200/host/Genus.h"
#include "../../ternary_birch_b200/host/IsometrySequence.h"

#ifdef TB_FACADE_GMPXX
// the binding's own instantiation: R = mpz_class through host/birch.h's IntegerTraits<mpz_class>
typedef mpz_class Big;
#else
// A minimal "big integer" to instantiate the precise (GMP-replacing) mode without GMP headers.
struct Big {
    long long v;
    Big() : v(0) {}
    Big(long long x) : v(x) {}
    bool operator<(const Big& o) const { return v < o.v; }
    bool operator==(const Big& o) const { return v == o.v; }
};
namespace tbirch {
template<> struct IntegerTraits<Big, void> {
    static const bool precise = true;
    static int64_t to_i64(const Big& x) { return x.v; }
    static Big from_i64(int64_t x) { return Big(x); }
};
}
#endif

static std::vector<int64_t> read_stream(const char *path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { std::perror(path); std::exit(2); }
    std::vector<int64_t> a((size_t)f.tellg() / 8);
    f.seekg(0);
    f.read((char *)a.data(), (std::streamsize)(a.size() * 8));
    return a;
}

template<typename R>
static void run(std::shared_ptr<tbirch::Table> table, long long p, std::vector<int64_t>& out)
{
    typedef tbirch::IntegerTraits<R> Tr;
    std::shared_ptr<Genus<R>> g = std::make_shared<Genus<R>>(Genus<R>::from_table(table));
    auto sparse = g->hecke_matrix_sparse(Tr::from_i64(p));
    auto dense = g->hecke_matrix_dense(Tr::from_i64(p));
    auto dims = g->dimension_map();
    out.push_back((int64_t)sparse.size());
    for (auto& kv : sparse)
    {
        const auto& m = kv.second;
        out.push_back(Tr::to_i64(kv.first));
        out.push_back((int64_t)dims[kv.first]);
        out.push_back((int64_t)m[0].size());
        for (int x : m[0]) out.push_back(x);
        for (int x : m[1]) out.push_back(x);
        for (int x : m[2]) out.push_back(x);
        // dense must agree with sparse
        const auto& d = dense[kv.first];
        const size_t dim = dims[kv.first];
        std::vector<int> rebuilt(dim * dim, 0);
        for (size_t r = 0; r < dim; r++)
            for (int i = m[2][r]; i < m[2][r + 1]; i++) rebuilt[r * dim + (size_t)m[1][i]] = m[0][i];
        out.push_back(rebuilt == d ? 1 : 0);
    }
    IsometrySequence<W16,W32,R> seq(g, Tr::from_i64(p));
    int count = 0;
    while (!seq.done() && count < 8)
    {
        IsometrySequenceData<R> d = seq.next();
        out.push_back(Tr::to_i64(d.isometry.a11)); out.push_back(Tr::to_i64(d.isometry.a22)); out.push_back(Tr::to_i64(d.isometry.a33));
        out.push_back(Tr::to_i64(d.denominator)); out.push_back((int64_t)d.src); out.push_back((int64_t)d.dst);
        ++count;
    }
    try { g->hecke_matrix_sparse(Tr::from_i64(table->primes[0])); out.push_back(0); }
    catch (const std::invalid_argument& e) { out.push_back(std::string(e.what()) == "Prime must not divide the discriminant." ? 1 : 0); }
}

int main(int argc, char **argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: facade_test <genus stream> <p> <out>\n"); return 2; }
    std::vector<int64_t> stream = read_stream(argv[1]);
    const long long p = std::atoll(argv[2]);
    std::vector<int64_t> out;
    try
    {
        std::shared_ptr<tbirch::Table> table = tbirch::Table::from_stream(stream.data(), stream.size());
        run<int64_t>(table, p, out);
        run<Big>(table, p, out);
        // converting constructor, as the binding does: Genus<Z64>(const Genus<Z>&)
        Genus<Big> gz = Genus<Big>::from_table(table);
        Genus<int64_t> g64(gz);
        out.push_back((int64_t)g64.size());
        // the binding's constructor path (pyx:164-186): get_quad_form, then Genus(q, symbols, seed)
        std::vector<PrimeSymbol<int64_t>> symbols;
        for (size_t i = 0; i < table->primes.size(); i++) symbols.push_back(PrimeSymbol<int64_t>{table->primes[i], 1, true});
        QuadForm<int64_t> q = QuadForm<int64_t>::get_quad_form(symbols);
        Genus<int64_t> built(q, symbols, table->seed);
        out.push_back((int64_t)built.size());
        out.push_back(built.table()->q == table->q && built.table()->s == table->s && built.table()->sinv == table->sinv &&
                      built.table()->lut == table->lut && built.table()->scalar == table->scalar ? 1 : 0);
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "facade_test: %s\n", e.what());
        return 1;
    }
    std::ofstream f(argv[3], std::ios::binary);
    f.write((const char *)out.data(), (std::streamsize)(out.size() * 8));
    return 0;
}
