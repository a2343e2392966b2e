// Genus.h -- facade twin of the reference's Genus<R> (Genus.h:27-857) for the
// p-neighbor hot path: same class template, same method names, argument meaning,
// return types and exception behaviour as the binding expects (pyx:78-96); every
// computation goes through libtbirch.so onto the GPU.
#ifndef TBIRCH_FACADE_GENUS_H
#define TBIRCH_FACADE_GENUS_H

#include "birch.h"
#include "QuadForm.h"
#include "Isometry.h"
#include "Eigenvector.h"

template<typename R>
class Genus
{
    template<typename T> friend class Genus;
    template<typename A, typename B, typename C> friend class IsometrySequence;
    typedef tbirch::IntegerTraits<R> Traits;

public:
    Genus() = default;

    // Genus<R>::Genus (Genus.h:37-246): enumerate the genus of q in the reference's discovery
    // order -- tb_genus_enumerate in the library (seed = 0 draws a random seed).
    Genus(const QuadForm<R>& q, const std::vector<PrimeSymbol<R>>& symbols, W64 seed = 0)
    {
        std::vector<tb_prime_symbol> syms;
        for (const PrimeSymbol<R>& s : symbols)
        {
            tb_prime_symbol t;
            t.p = Traits::to_i64(s.p); t.power = s.power; t.ramified = s.ramified ? 1 : 0;
            syms.push_back(t);
        }
        const int64_t form[6] = { Traits::to_i64(q.a()), Traits::to_i64(q.b()), Traits::to_i64(q.c()),
                                  Traits::to_i64(q.f()), Traits::to_i64(q.g()), Traits::to_i64(q.h()) };
        tb_table *built = nullptr;
        tbirch::check(tb_genus_enumerate(form, syms.data(), (int64_t)syms.size(), seed, &built));
        tb_genus_desc d;
        tbirch::check(tb_table_desc(built, &d, nullptr, nullptr, nullptr));
        table_ = tbirch::Table::from_desc(d);
        tb_table_destroy(built);
    }

    static Genus<R> from_table(std::shared_ptr<tbirch::Table> table)
    {
        Genus<R> g;
        g.table_ = table;
        return g;
    }
    static Genus<R> from_stream(const int64_t *stream, size_t len)
    {
        return from_table(tbirch::Table::from_stream(stream, len));
    }

    // Genus<Z64>(const Genus<Z>&), Genus.h:248-303: same table, other integer mode.
    template<typename T>
    Genus(const Genus<T>& src) : table_(src.table_) {}

    template<typename T>
    static Genus<T> convert(const Genus<R>& src) { return Genus<T>(src); }

    size_t size(void) const { return (size_t)table_->N; }                       // Genus.h:306
    W64 seed(void) const { return table_->seed; }                                // Genus.h:311

    std::map<R,size_t> dimension_map(void) const                                // Genus.h:316-325
    {
        std::map<R,size_t> temp;
        for (size_t k = 0; k < table_->conductors.size(); k++)
            temp[Traits::from_i64(table_->conductors[k])] = (size_t)table_->dims[k];
        return temp;
    }

    std::map<R,std::vector<int>> hecke_matrix_dense(const R& p) const           // Genus.h:332-339
    {
        Hecke h(*this, p);
        std::map<R,std::vector<int>> matrices;
        for (int64_t k = 0; k < tb_hecke_num_conductors(h.ptr); k++)
        {
            const size_t dim = (size_t)tb_hecke_dim(h.ptr, k);
            std::vector<int> m(dim * dim);
            tbirch::check(tb_hecke_copy_dense(h.ptr, k, m.data()));
            matrices[Traits::from_i64(tb_hecke_conductor(h.ptr, k))] = std::move(m);
        }
        return matrices;
    }

    // order of the inner vectors: data, indices, indptr (Genus.h:662-671)
    std::map<R,std::vector<std::vector<int>>> hecke_matrix_sparse(const R& p) const   // Genus.h:341-348
    {
        Hecke h(*this, p);
        std::map<R,std::vector<std::vector<int>>> csr;
        for (int64_t k = 0; k < tb_hecke_num_conductors(h.ptr); k++)
        {
            const size_t nnz = (size_t)tb_hecke_nnz(h.ptr, k), rows = (size_t)tb_hecke_rows(h.ptr, k);
            std::vector<std::vector<int>> m(3);
            m[0].resize(nnz); m[1].resize(nnz); m[2].resize(rows + 1);
            tbirch::check(tb_hecke_copy_sparse(h.ptr, k, m[0].data(), m[1].data(), m[2].data()));
            csr[Traits::from_i64(tb_hecke_conductor(h.ptr, k))] = std::move(m);
        }
        return csr;
    }

    Eigenvector<R> eigenvector(const std::vector<Z32>& vec, const R& conductor) const   // Genus.h:350-390
    {
        const int64_t cond = Traits::to_i64(conductor);
        size_t k = 0;
        const size_t C = table_->conductors.size();
        for (; k < C; k++) if (table_->conductors[k] == cond) break;
        if (k == C) throw std::invalid_argument("Invalid conductor.");
        if ((size_t)table_->dims[k] != vec.size()) throw std::invalid_argument("Eigenvector has incorrect dimension.");
        const size_t N = size();
        std::vector<Z32> temp(N);
        const int32_t *lut = &table_->lut[k * N];
        for (size_t n = 0; n < N; n++) if (lut[n] != -1) temp[n] = vec[(size_t)lut[n]];
        return Eigenvector<R>(std::move(temp), k);
    }

    std::vector<Z32> eigenvalues(EigenvectorManager<R>& manager, const R& p) const      // Genus.h:392-421
    {
        const size_t V = manager.size(), N = size();
        std::vector<Z32> out(V);
        if (V == 0) return out;
        std::vector<Z32> vecs(V * N);
        std::vector<int64_t> cond(V), rep(V);
        for (size_t v = 0; v < V; v++)
        {
            const Eigenvector<R>& e = manager[v];
            std::copy(e.data().begin(), e.data().end(), vecs.begin() + v * N);
            cond[v] = (int64_t)e.conductor_index();
            rep[v] = (int64_t)e.rep_index();
        }
        tbirch::check(tb_eigenvalues(table_->device(), (uint64_t)Traits::to_i64(p), Traits::precise ? 1 : 0,
                                     (int64_t)V, vecs.data(), cond.data(), rep.data(), out.data()));
        return out;
    }

    const std::shared_ptr<tbirch::Table>& table(void) const { return table_; }

private:
    std::shared_ptr<tbirch::Table> table_;

    struct Hecke {
        tb_hecke *ptr = nullptr;
        Hecke(const Genus<R>& g, const R& p)
        {
            tbirch::check(tb_hecke_compute(g.table_->device(), (uint64_t)Traits::to_i64(p), Traits::precise ? 1 : 0, &ptr));
        }
        ~Hecke() { tb_hecke_destroy(ptr); }
    };
};

#endif  // TBIRCH_FACADE_GENUS_H
