// Isometry.h -- facade twin of Isometry<R> as the binding reads it (pyx:98-107):
// the nine public fields, row-major.
#ifndef TBIRCH_FACADE_ISOMETRY_H
#define TBIRCH_FACADE_ISOMETRY_H

#include "birch.h"

template<typename R>
class Isometry
{
public:
    Isometry() : a11(1), a12(0), a13(0), a21(0), a22(1), a23(0), a31(0), a32(0), a33(1) {}
    R a11, a12, a13;
    R a21, a22, a23;
    R a31, a32, a33;
};

#endif  // TBIRCH_FACADE_ISOMETRY_H
