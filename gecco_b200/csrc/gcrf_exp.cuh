// gcrf_exp.cuh — exp() for the f64 reference-order kernel (gcrf_exact.cu), evaluated in double-double arithmetic and
// rounded once: the result is the correctly rounded double except when exp(x) lies within ~2^-100 (relative) of a
// rounding boundary.
//
// Why not the CUDA math library's exp(): it is specified to 1 ulp, the host libm the reference's tagger calls
// (CRFsuite crf1d_context.c, `exp` of every state and transition score) to well under 1 ulp.  The marginals are
// compared with the reference's own numbers digit by digit (tests/golden/bgc0001866.json holds python-crfsuite's
// output to 17 digits), so the exponentials have to agree with a good host libm to the last bit in practice.
//
// Scheme: x = n (ln2/64) + r, n = 64 k + j, |r| <= ln2/128; exp(x) = 2^k * T[j] * P(r), T[j] = exp(j ln2/64) as
// (hi, lo), P = Taylor polynomial: degrees 0..6 in double-double Horner form, degrees 7..12 as a plain-double tail
// (those terms are below 2^-65).  Constants: tools/gen_exp_tables.py (80-digit decimals).
//
// The file compiles as host code too (tests/test_exp_dd.py builds tools/exp_dd_check.cpp with g++ and checks the
// algorithm against 60-digit decimals), which is why it uses fma() and no device intrinsics.
#pragma once

#include <cmath>

#ifdef __CUDACC__
#define GCRF_HD __device__ __forceinline__
#else
#define GCRF_HD inline
#define __device__
#endif

namespace gcrf {
namespace expdd {

#include "gcrf_exp_tables.inc"

struct dd {
    double h, l;
};

// s + e = a + b exactly (Knuth)
GCRF_HD dd two_sum(double a, double b) {
    const double s = a + b;
    const double bb = s - a;
    const double e = (a - (s - bb)) + (b - bb);
    return {s, e};
}
// s + e = a + b exactly when |a| >= |b|
GCRF_HD dd fast_two_sum(double a, double b) {
    const double s = a + b;
    return {s, b - (s - a)};
}
GCRF_HD dd two_prod(double a, double b) {
    const double p = a * b;
    return {p, fma(a, b, -p)};
}
GCRF_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.h, b.h);
    p.l += a.h * b.l + a.l * b.h;
    return fast_two_sum(p.h, p.l);
}
GCRF_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.h, b.h);
    const dd t = two_sum(a.l, b.l);
    s.l += t.h;
    s = fast_two_sum(s.h, s.l);
    s.l += t.l;
    return fast_two_sum(s.h, s.l);
}

// exp(x) for finite x in the range where neither the result nor the scaling underflows; other arguments take the
// library function (inf / 0 / NaN semantics; the CRF never produces them on meaningful input).
GCRF_HD double exp_cr(double x) {
    if (!(x > -700.0 && x < 700.0)) return exp(x);
    const double nd = rint(x * kInvStep);
    const int n = (int)nd;
    // r = x - n ln2/64 as (hi, lo): n * kStep1 is exact (35-bit constant), the subtraction is exact (Sterbenz)
    const double r0 = fma(-nd, kStep1, x);
    const dd p2 = two_prod(nd, kStep2);
    dd r = two_sum(r0, -p2.h);
    r.l -= p2.l + nd * kStep3;
    r = fast_two_sum(r.h, r.l);
    // tail: sum_{i=7..12} r^i / i!, plain double
    const double rh = r.h;
    double t = kInvFactTail[5];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 4; i >= 0; --i) t = fma(t, rh, kInvFactTail[i]);
    const double r2 = rh * rh, r4 = r2 * r2;
    t *= r4 * r2 * rh;
    // head: degrees 6..0, double-double Horner
    dd acc = {kInvFact[6][0], kInvFact[6][1]};
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 5; i >= 0; --i) {
        acc = dd_mul(acc, r);
        acc = dd_add(acc, dd{kInvFact[i][0], kInvFact[i][1]});
    }
    acc = dd_add(acc, dd{t, 0.0});
    const int j = n & 63, k = (n - j) / 64;
    acc = dd_mul(acc, dd{kExpTab[j][0], kExpTab[j][1]});
    return ldexp(acc.h + acc.l, k);
}

}  // namespace expdd
}  // namespace gcrf
