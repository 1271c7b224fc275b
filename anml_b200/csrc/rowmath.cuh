// Per-row scalar math: SmoothMapping values/derivatives and the family NLL terms.
#pragma once
#include "common.cuh"

namespace anml {

enum Link : int { LINK_IDENTITY = 0, LINK_EXP = 1, LINK_LOG = 2, LINK_EXPIT = 3, LINK_LOGIT = 4 };
enum Family : int { FAM_GAUSSIAN = 0, FAM_BINOMIAL = 1, FAM_POISSON = 2, FAM_GAUSSIAN_MEANVAR = 3 };

struct LinkVal {
    double f, d1, d2;  // f(y), f'(y), f''(y)
    bool ok;           // y inside the mapping's domain
};

// anml's SmoothMapping family, src/anml/parameter/smoothmapping.py.
// expit keeps the reference's exp(-y) formulation (:121-126) so overflow
// behaviour (nan at order >= 1 for y < -709.78) is the same.
__device__ __forceinline__ LinkVal link_eval(int link, double y) {
    LinkVal v;
    v.ok = true;
    switch (link) {
        case LINK_EXP: {  // :76-78
            const double e = exp(y);
            v.f = e; v.d1 = e; v.d2 = e;
            break;
        }
        case LINK_LOG: {  // :96-104, domain :98-99
            v.ok = (y > 0.0);
            const double inv = 1.0 / y;
            v.f = log(y); v.d1 = inv; v.d2 = -inv * inv;
            break;
        }
        case LINK_EXPIT: {  // :119-126
            // the reference's expressions with the reference's overflow structure: z / (1+z)**2 and
            // z (z-1) / (z+1)**3 are 0 once the denominator overflows and nan once the numerator does
            // too (y < -355 resp. -709.78) -- multiplying by 1/(1+z) instead would return tiny values or inf
            // at those arguments
            const double z = exp(-y);
            const double s = 1.0 + z;
            v.f = 1.0 / s; v.d1 = z / (s * s); v.d2 = z * (z - 1.0) / (s * s * s);
            break;
        }
        case LINK_LOGIT: {  // :146-156, domain :148-151
            v.ok = (y > 0.0) && (y < 1.0);
            const double s = y * (1.0 - y);
            const double inv = 1.0 / s;
            v.f = log(y / (1.0 - y)); v.d1 = inv; v.d2 = (2.0 * y - 1.0) * inv * inv;
            break;
        }
        default:  // identity :59-65
            v.f = y; v.d1 = 1.0; v.d2 = 0.0;
            break;
    }
    return v;
}

// Value of one order only (Parameter.get_params hands transform(y, order) on).
__device__ __forceinline__ double link_order(int link, double y, int order, bool& ok) {
    const LinkVal v = link_eval(link, y);
    ok = v.ok;
    return order == 0 ? v.f : (order == 1 ? v.d1 : v.d2);
}

// Per-row terms of a one-parameter family, already chained through the link:
//   l = NLL term, r = dl/dtheta * f', w = d2l/dtheta2 * f'^2 + dl/dtheta * f''
// so that grad = X^T r and Hessian = X^T diag(w) X  (SURVEY.md 3.3).
struct RowTerms1 {
    double l, r, w;
    double sw;  // sqrt(|w|), filled only when asked for
    double ls;  // DEFER_LOG only: the row's NLL term is l + log(ls) (1.0 when nothing is deferred)
    bool ok;
};

// DEFER_LOG: the canonical logistic term log(1 + e^-y) is not evaluated per row; the caller multiplies
// the returned factors ls = 1 + e^-y into a running product (LogProduct below) and takes ONE log at the
// end.  In the warp-specialised Hessian kernel every FP64 instruction of a helper warp has to squeeze into
// the DMMA stream, and libdevice's log is 22 of them per row.
template <bool WANT_SQRT = false, bool DEFER_LOG = false>
__device__ __forceinline__ RowTerms1 row_terms1(int family, int link, bool fast, double y, double obs, double se) {
    RowTerms1 o;
    o.ok = true;
    o.sw = 0.0;
    o.ls = 1.0;
    if (fast && family == FAM_BINOMIAL && link == LINK_EXPIT) {
        // canonical logistic: one exp, one log, one reciprocal per row.
        // h = exp(-y/2): z = h^2 and sqrt(w) = h q come for free (no sqrt).
        const double h = exp(-0.5 * y);
        const double z = h * h;
        const double s = 1.0 + z;
        const double q = 1.0 / s;
        if (DEFER_LOG) {
            o.ls = s;
            o.l = (1.0 - obs) * y;
        } else {
            o.l = log(s) + (1.0 - obs) * y;
        }
        o.r = q - obs;
        o.sw = h * q;
        o.w = o.sw * o.sw;
        return o;
    }
    if (fast && family == FAM_GAUSSIAN && link == LINK_IDENTITY) {
        const double is = (se == 1.0) ? 1.0 : 1.0 / se;  // the default obs_se: no division
        const double iv = is * is;
        const double res = y - obs;
        o.r = res * iv;
        o.w = iv;
        o.sw = is;
        o.l = 0.5 * res * o.r;
        return o;
    }
    if (fast && family == FAM_POISSON && link == LINK_EXP) {
        const double h = exp(0.5 * y);
        const double lam = h * h;
        o.l = lam - obs * y;
        o.r = lam - obs;
        o.w = lam;
        o.sw = h;
        return o;
    }
    const LinkVal f = link_eval(link, y);
    o.ok = f.ok;
    double l, d1, d2;
    const double th = f.f;
    if (family == FAM_BINOMIAL) {
        const double om = 1.0 - th;
        l = -(obs * log(th) + (1.0 - obs) * log(om));
        d1 = -obs / th + (1.0 - obs) / om;
        d2 = obs / (th * th) + (1.0 - obs) / (om * om);
    } else if (family == FAM_POISSON) {
        l = th - obs * log(th);
        d1 = 1.0 - obs / th;
        d2 = obs / (th * th);
    } else {  // gaussian, fixed se
        const double iv = 1.0 / (se * se);
        const double res = obs - th;
        l = 0.5 * res * res * iv;
        d1 = -res * iv;
        d2 = iv;
    }
    o.l = l;
    o.r = d1 * f.d1;
    o.w = d2 * f.d1 * f.d1 + d1 * f.d2;
    if (WANT_SQRT) o.sw = sqrt(fabs(o.w));
    return o;
}

// Running product of factors >= 1 kept as mantissa in [1, 2) x 2^exponent: sum_i log(f_i) =
// exponent * ln 2 + log(mantissa).  One DMUL per factor; the renormalisation is integer work on the
// exponent field.  Relative rounding error grows like sqrt(#factors) * 2^-53 on the mantissa, i.e. an
// ABSOLUTE error of ~1e-14 on a sum of thousands of logs -- tighter than adding rounded logs.  A factor
// that is inf or NaN (exponent field 0x7ff) stays in the mantissa and poisons the result like log would.
struct LogProduct {
    double mant = 1.0;
    int expo = 0;  // < 1075 per factor: no overflow below 2e6 factors per lane
    __device__ __forceinline__ void mul(double f) {
        mant *= f;
        const int hi = __double2hiint(mant);
        const int ef = (hi >> 20) & 0x7ff;
        if (ef != 0x7ff) {
            expo += ef - 1023;
            mant = __hiloint2double(hi - ((ef - 1023) << 20), __double2loint(mant));
        }
    }
    __device__ __forceinline__ double log_value() const { return (double)expo * 0.6931471805599453094 + log(mant); }
};

// Two-parameter Gaussian (mean mu = f0(y0), variance v = f1(y1)):
//   l = 1/2 log v + 1/2 (obs-mu)^2 / v
struct RowTerms2 {
    double l, r0, r1, w00, w01, w11;
    double sw00;  // sqrt(|w00|) when it comes for free (fast path), else < 0
    bool ok;
};

__device__ __forceinline__ RowTerms2 row_terms2(int link0, int link1, double y0, double y1, double obs, bool fast = false) {
    RowTerms2 o;
    o.sw00 = -1.0;
    if (fast && link0 == LINK_IDENTITY && link1 == LINK_EXP) {
        // mean = y0, variance = e^{y1}: one exp per row and no log, division or square root.
        // h = e^{-y1/2} = 1/sqrt(v); with a = res/v and b = res^2/v the general expressions below reduce to
        //   l = (y1 + b)/2, r0 = -a, r1 = (1 - b)/2, w00 = 1/v, w01 = a, w11 = b/2
        const double h = exp(-0.5 * y1);
        const double iv = h * h;
        const double res = obs - y0;
        const double a = res * iv;
        const double b = res * a;
        o.ok = true;
        o.l = 0.5 * (y1 + b);
        o.r0 = -a;
        o.r1 = 0.5 - 0.5 * b;
        o.w00 = iv;
        o.w01 = a;
        o.w11 = 0.5 * b;
        o.sw00 = h;
        return o;
    }
    const LinkVal f0 = link_eval(link0, y0);
    const LinkVal f1 = link_eval(link1, y1);
    o.ok = f0.ok && f1.ok;
    const double mu = f0.f, v = f1.f;
    const double iv = 1.0 / v;
    const double res = obs - mu;
    const double d0 = -res * iv;
    const double d1 = 0.5 * iv - 0.5 * res * res * iv * iv;
    const double h00 = iv;
    const double h01 = res * iv * iv;
    const double h11 = -0.5 * iv * iv + res * res * iv * iv * iv;
    o.l = 0.5 * log(v) + 0.5 * res * res * iv;
    o.r0 = d0 * f0.d1;
    o.r1 = d1 * f1.d1;
    o.w00 = h00 * f0.d1 * f0.d1 + d0 * f0.d2;
    o.w01 = h01 * f0.d1 * f1.d1;
    o.w11 = h11 * f1.d1 * f1.d1 + d1 * f1.d2;
    return o;
}

}  // namespace anml
