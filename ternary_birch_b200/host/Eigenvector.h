// Eigenvector.h -- facade twins of Eigenvector<R> / EigenvectorManager<R>
// (reference Eigenvector.h:7-240; binding: pyx:68-76).
#ifndef TBIRCH_FACADE_EIGENVECTOR_H
#define TBIRCH_FACADE_EIGENVECTOR_H

#include <algorithm>
#include "birch.h"

template<typename R>
class Eigenvector
{
public:
    Eigenvector() = default;
    Eigenvector(std::vector<Z32>&& data, W64 conductor_index)
        : data_(std::move(data)), conductor_index_(conductor_index), rep_index_(0) {}

    const std::vector<Z32>& data(void) const { return data_; }
    size_t size(void) const { return data_.size(); }
    W64 conductor_index(void) const { return conductor_index_; }
    void rep_index(size_t pos) { rep_index_ = pos; }
    size_t rep_index(void) const { return rep_index_; }
    Z32 operator[](size_t pos) const { return data_[pos]; }

private:
    std::vector<Z32> data_;
    W64 conductor_index_ = 0;
    size_t rep_index_ = 0;
};

template<typename R>
class EigenvectorManager
{
    friend class Genus<R>;

public:
    EigenvectorManager() = default;

    void add_eigenvector(const Eigenvector<R>& vector)
    {
        if (finalized) throw std::logic_error("Cannot add eigenvectors once finalized.");
        if (dimension > 0)
        {
            if (dimension != vector.size()) throw std::invalid_argument("Eigenvector dimensions must match.");
        }
        else dimension = vector.size();
        eigenvectors.push_back(vector);
    }

    size_t size(void) const { return eigenvectors.size(); }
    const Eigenvector<R>& operator[](size_t index) const { return eigenvectors[index]; }

    // Chooses, for every eigenvector, a genus representative where its coordinate is
    // nonzero (Eigenvector.h:95-185).  The reference solves a set cover with its
    // SetCover class; any cover gives the same eigenvalues, so a greedy one is used.
    void finalize(void)
    {
        if (finalized) throw std::logic_error("Cannot finalize again.");
        finalized = true;
        const size_t V = eigenvectors.size();
        if (V == 0) return;
        std::vector<bool> covered(V, false);
        size_t left = V;
        while (left)
        {
            size_t best = 0, best_gain = 0;
            for (size_t pos = 0; pos < dimension; ++pos)
            {
                size_t gain = 0;
                for (size_t v = 0; v < V; ++v) if (!covered[v] && eigenvectors[v][pos]) ++gain;
                if (gain > best_gain) { best_gain = gain; best = pos; }
            }
            if (best_gain == 0) throw std::runtime_error("Eigenvector is identically zero.");
            indices.push_back((int64_t)best);
            for (size_t v = 0; v < V; ++v)
                if (!covered[v] && eigenvectors[v][best]) { covered[v] = true; --left; }
        }
        std::sort(indices.begin(), indices.end());
        for (size_t v = 0; v < V; ++v)
            for (int64_t index : indices)
                if (eigenvectors[v][(size_t)index]) { eigenvectors[v].rep_index((size_t)index); break; }
    }

private:
    bool finalized = false;
    size_t dimension = 0;
    std::vector<Eigenvector<R>> eigenvectors;
    std::vector<int64_t> indices;
};

#endif  // TBIRCH_FACADE_EIGENVECTOR_H
