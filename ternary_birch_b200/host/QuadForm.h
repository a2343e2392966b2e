// QuadForm.h -- facade twin of the reference's QuadForm<R> as the binding uses it
// (pyx:57-66): coefficient accessors and the static get_quad_form.
#ifndef TBIRCH_FACADE_QUADFORM_H
#define TBIRCH_FACADE_QUADFORM_H

#include "birch.h"

template<typename R>
class QuadForm
{
public:
    QuadForm() = default;
    QuadForm(const R& a, const R& b, const R& c, const R& f, const R& g, const R& h)
        : a_(a), b_(b), c_(c), f_(f), g_(g), h_(h) {}

    const R& a(void) const { return a_; }
    const R& b(void) const { return b_; }
    const R& c(void) const { return c_; }
    const R& f(void) const { return f_; }
    const R& g(void) const { return g_; }
    const R& h(void) const { return h_; }

    R discriminant(void) const        // QuadForm.h:27-32
    {
        return a_ * (4 * b_ * c_ - f_ * f_) - b_ * g_ * g_ + h_ * (f_ * g_ - c_ * h_);
    }
    bool operator==(const QuadForm<R>& q) const
    {
        return a_ == q.a_ && b_ == q.b_ && c_ == q.c_ && f_ == q.f_ && g_ == q.g_ && h_ == q.h_;
    }

    // QuadForm.cpp:444-777 builds the form of a given ramification by a search that is
    // not on the accelerated path (one call per genus, milliseconds).  A deployment keeps
    // the reference's own translation unit for it; this facade does not re-implement it.
    static QuadForm<R> get_quad_form(const std::vector<PrimeSymbol<R>>&)
    {
        throw std::logic_error("QuadForm::get_quad_form is outside the accelerated p-neighbor path; "
                               "link the reference's QuadForm.cpp or build the genus from a table.");
    }

protected:
    R a_, b_, c_, f_, g_, h_;
};

#endif  // TBIRCH_FACADE_QUADFORM_H
