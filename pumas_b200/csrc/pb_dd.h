/*
 * pb_dd.h -- double-double arithmetic (unevaluated sum of two binary64, ~106
 * bit significand) for the one place the reference relies on x87 extended
 * precision at run time: the pole reduction of the screened Coulomb DCS and
 * its integrals (src/pumas.c:7524-7683), which cancel catastrophically in
 * binary64. GPUs have no 80 bit type; double-double is strictly more precise.
 *
 * Error free transformations after Dekker (1971) and Knuth, with the product
 * error from a fused multiply-add; algorithms as in Hida, Li & Bailey,
 * "Library for double-double and quad-double arithmetic" (2007). The explicit
 * fma() calls are required (and are not subject to -fmad=false).
 */
#ifndef PB_DD_H
#define PB_DD_H

#include "pb_common.h"

namespace pb {

struct dd {
        double hi, lo;
        PB_FN dd() : hi(0.), lo(0.) {}
        PB_FN explicit dd(double x) : hi(x), lo(0.) {}
        PB_FN dd(double h, double l) : hi(h), lo(l) {}
        PB_FN explicit operator double() const { return hi + lo; }
};

PB_FN dd dd_quick_two_sum(double a, double b)
{
        const double s = a + b;
        return dd(s, b - (s - a));
}

PB_FN dd dd_two_sum(double a, double b)
{
        const double s = a + b;
        const double bb = s - a;
        return dd(s, (a - (s - bb)) + (b - bb));
}

PB_FN dd dd_two_prod(double a, double b)
{
        const double pr = a * b;
        return dd(pr, fma(a, b, -pr));
}

PB_FN dd operator+(const dd & a, const dd & b)
{
        dd s = dd_two_sum(a.hi, b.hi);
        const dd t = dd_two_sum(a.lo, b.lo);
        s.lo += t.hi;
        s = dd_quick_two_sum(s.hi, s.lo);
        s.lo += t.lo;
        return dd_quick_two_sum(s.hi, s.lo);
}

PB_FN dd operator-(const dd & a) { return dd(-a.hi, -a.lo); }

PB_FN dd operator-(const dd & a, const dd & b) { return a + (-b); }

PB_FN dd operator*(const dd & a, const dd & b)
{
        dd pr = dd_two_prod(a.hi, b.hi);
        pr.lo += (a.hi * b.lo + a.lo * b.hi);
        return dd_quick_two_sum(pr.hi, pr.lo);
}

PB_FN dd operator/(const dd & a, const dd & b)
{
        const double q1 = a.hi / b.hi;
        dd r = a - b * dd(q1);
        const double q2 = r.hi / b.hi;
        r = r - b * dd(q2);
        const double q3 = r.hi / b.hi;
        const dd q = dd_quick_two_sum(q1, q2);
        return q + dd(q3);
}

/* exp(a) for |a| of order 1: argument reduction by ln2 and 2^-9, Taylor
 * series, nine squarings. */
PB_FN dd dd_exp(const dd & a)
{
        const dd ln2(6.931471805599452862e-01, 2.319046813846299558e-17);
        const double k = floor(a.hi / ln2.hi + 0.5);
        const double inv_k = 1.0 / 512.0;
        dd r = a - ln2 * dd(k);
        r = dd(r.hi * inv_k, r.lo * inv_k);

        /* s = r + r^2/2! + r^3/3! + ... */
        dd term = r;
        dd s = r;
        for (int i = 2; i <= 10; i++) {
                term = term * r;
                term = term / dd((double)i);
                s = s + term;
        }
        /* (1 + s)^512 - 1 by repeated s <- 2 s + s^2 */
        for (int i = 0; i < 9; i++) {
                s = dd(2. * s.hi, 2. * s.lo) + s * s;
        }
        s = s + dd(1.);
        const int ik = (int)k;
        return dd(ldexp(s.hi, ik), ldexp(s.lo, ik));
}

/* log(a), a > 0: one Newton step on exp from the binary64 estimate. */
PB_FN dd dd_log(const dd & a)
{
        const dd x(log(a.hi));
        return x + a * dd_exp(-x) - dd(1.);
}

} /* namespace pb */

#endif
