// IsometrySequence.h -- facade twin of IsometrySequence<R,S,T> (reference
// IsometrySequence.h:7-102; binding: pyx:109-118): rep-major, t-minor stream of
// (isometry, denominator, src, dst), computed on the GPU in chunks.
#ifndef TBIRCH_FACADE_ISOMETRY_SEQUENCE_H
#define TBIRCH_FACADE_ISOMETRY_SEQUENCE_H

#include "birch.h"
#include "Genus.h"

template<typename T>
class IsometrySequenceData
{
public:
    Isometry<T> isometry;
    T denominator;
    size_t src;
    size_t dst;
};

template<typename R, typename S, typename T>
class IsometrySequence
{
    typedef tbirch::IntegerTraits<T> Traits;

public:
    IsometrySequence(std::shared_ptr<Genus<T>> genus, const T& p) : genus_(genus)
    {
        tbirch::check(tb_isoseq_open(genus->table_->device(), (uint64_t)Traits::to_i64(p),
                                     Traits::precise ? 1 : 0, &it_));
    }
    ~IsometrySequence() { tb_isoseq_close(it_); }
    IsometrySequence(const IsometrySequence&) = delete;
    IsometrySequence& operator=(const IsometrySequence&) = delete;

    bool done() const { return tb_isoseq_done(it_) != 0; }

    IsometrySequenceData<T> next()
    {
        int64_t iso[9], den = 0, src = 0, dst = 0;
        tbirch::check(tb_isoseq_next(it_, iso, &den, &src, &dst));     // "No more isometries." past the end
        IsometrySequenceData<T> d;
        d.isometry.a11 = Traits::from_i64(iso[0]); d.isometry.a12 = Traits::from_i64(iso[1]); d.isometry.a13 = Traits::from_i64(iso[2]);
        d.isometry.a21 = Traits::from_i64(iso[3]); d.isometry.a22 = Traits::from_i64(iso[4]); d.isometry.a23 = Traits::from_i64(iso[5]);
        d.isometry.a31 = Traits::from_i64(iso[6]); d.isometry.a32 = Traits::from_i64(iso[7]); d.isometry.a33 = Traits::from_i64(iso[8]);
        d.denominator = Traits::from_i64(den);
        d.src = (size_t)src;
        d.dst = (size_t)dst;
        return d;
    }

private:
    std::shared_ptr<Genus<T>> genus_;
    tb_isoseq *it_ = nullptr;
};

#endif  // TBIRCH_FACADE_ISOMETRY_SEQUENCE_H
