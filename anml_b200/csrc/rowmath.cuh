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

// ---- short exp / reciprocal for the fast paths -------------------------------------------------
// In the warp-specialised Hessian kernel every FP64 instruction of a helper warp waits ~50 cycles for a
// slot of the pipe the DMMA stream saturates, and the tensor warp waits for the helpers: the helpers' time
// is their FP64 instruction count.  libdevice's exp is ~19 FP64 instructions (no table, degree-10
// polynomial), the IEEE division ~9.  exp_short: e^x = 2^k * 2^(j/64) * e^r with a 64-entry table and a
// degree-5 polynomial on |r| <= ln2/128 -- 10 FP64 instructions, <= 1 ulp (checked against a 60-digit
// reference over [-700, 700]); the scale by 2^k is integer work on the exponent field.  Arguments that
// could leave the normal range (|x| >= 708, inf, nan) take libdevice's exp.  rcp_short: MUFU.RCP64H seed +
// two Newton steps (4 FMAs), <= 1 ulp, for arguments in the normal range (here s = 1 + z >= 1).
__device__ const double EXP2_64TH[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

__device__ __forceinline__ double exp_short(double x) {
#ifdef ANML_LIBDEVICE_EXP
    return exp(x);
#else
    if ((unsigned)(__double2hiint(x) & 0x7fffffff) >= 0x40862000u) return exp(x);  // |x| >= 708, inf, nan
    const double MAGIC = 0x1.8p52;                                                 // rounds to an integer in the low word
    double kf = fma(x, 0x1.71547652b82fep+6 /* 64/ln2 */, MAGIC);
    const int ki = __double2loint(kf);
    kf -= MAGIC;
    double r = fma(kf, -0x1.62e42fef80000p-7 /* ln2/64, 34 bits */, x);
    r = fma(kf, -0x1.1cf79abc9e3b4p-42, r);
    const double t = EXP2_64TH[ki & 63];
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p *= r;
    const double v = fma(t, p, t);
    return __hiloint2double(__double2hiint(v) + ((ki >> 6) << 20), __double2loint(v));
#endif
}

__device__ __forceinline__ double rcp_short(double s) {
#ifdef ANML_LIBDEVICE_EXP
    return 1.0 / s;
#else
    if ((unsigned)(__double2hiint(s) & 0x7fffffff) >= 0x7fd00000u) return 1.0 / s;  // huge, inf, nan: the IEEE path
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(s));
    double e = fma(-s, x, 1.0);
    x = fma(x, e, x);
    e = fma(-s, x, 1.0);
    return fma(x, e, x);
#endif
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
        const double h = exp_short(-0.5 * y);
        const double z = h * h;
        const double s = 1.0 + z;
        const double q = rcp_short(s);
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
        const double h = exp_short(0.5 * y);
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
        const double h = exp_short(-0.5 * y1);
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
