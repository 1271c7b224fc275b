// Banded evaluator for a model whose only variable is one B-spline (BASELINE config #3).
//
// A row of a spline design has degree+1 non-zeros, so the dense path (build the (n, nb) matrix,
// stream it through the DMMA kernel) moves 8*nb bytes per row to do (degree+1)^2 useful
// multiply-adds.  Here the basis is evaluated on the fly from x ("for C3 with
// on-the-fly basis P_X = 1, i.e. just x"): per row one interval lookup, the degree+1 non-zero basis
// values, the row terms and degree+1 + (degree+1)(degree+2)/2 FMAs into the band block of its knot
// interval.  The kernel is bound by FP64 instruction issue (B200 has 64 FP64 lanes per SM), so the
// basis is NOT run through the Cox-de Boor divisions per row: on one knot interval every basis
// function is a polynomial of degree `degree`, whose Taylor coefficients about the interval's left
// knot are tabulated once per spline (banded_coeff_kernel, from the same local_basis_ders as the design
// kernel: c_k = N^(k)(t_idx) / k!) and evaluated by Horner -- degree FMAs per basis value instead of
// ~150 FP64 instructions per row.  The values agree with the design kernel's to rounding (1e-16), not
// bit for bit; the design matrix itself is still produced by spline.cuh.
//
// No floating-point atomics: the rows are sorted by x once at attach (radix sort of (x, row)),
// every warp owns a contiguous range of the sorted rows, so the interval index is non-decreasing
// along its range; a lane keeps the band block of the current interval in registers, and when
// the warp moves to the next interval the 32 lane blocks are reduced in a fixed order and
// written to the warp's partial record (each (warp, interval) slot is written at most once).
// Records are summed over warps in warp order: results are bitwise reproducible.
#pragma once
#include "common.cuh"
#include "rowmath.cuh"
#include "spline.cuh"

namespace anml {

struct BandedArgs {
    const double* xs;       // abscissae, sorted ascending (n)
    const double* obs;      // per-row vectors gathered into the same order (se / offset / weights may be null)
    const double* se;
    const double* offset;
    const double* weights;
    long long n;
    const double* t;        // clamped knot vector, nk + 2*degree entries
    const double* coef;     // [nk-1][degree+1 bases][degree+1 powers] Taylor coefficients about the left knot
    int coef_in_smem;       // the table is copied to shared memory when it is small enough
    int nk;
    const double* beta;     // nb = nk - 1 + degree coefficients
    int family, link, fast, want_hess;
    long long rows_per_warp;  // multiple of 32
    double* partials;       // [warps][rec], rec = (nk-1) * slots + 1, zeroed before the launch
    int* status;
};

template <int DEG>
struct BandedCfg {
    static constexpr int NG = DEG + 1;
    static constexpr int NH = (DEG + 1) * (DEG + 2) / 2;
    static constexpr int SLOTS = NG + NH;  // per interval: gradient entries, then the lower triangle row by row
};

// COEF_SMEM: the coefficient table fits in shared memory (LDS.128 instead of generic loads);
// HESS: the Hessian band is wanted -- both compile-time, the row loop is bound by instruction issue.
template <int DEG, bool COEF_SMEM, bool HESS>
__global__ void __launch_bounds__(256, 2) spline_banded_eval_kernel(const BandedArgs a) {
    using C = BandedCfg<DEG>;
    extern __shared__ __align__(16) double sm_banded[];
    const int nt = a.nk + 2 * DEG;
    const int nb = a.nk - 1 + DEG;
    double* t_s = sm_banded;
    double* beta_s = sm_banded + ((nt + 1) & ~1);
    double* coef_s = beta_s + ((nb + 1) & ~1);
    for (int i = threadIdx.x; i < nt; i += blockDim.x) t_s[i] = a.t[i];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) beta_s[i] = a.beta[i];
    if (COEF_SMEM)
        for (int i = threadIdx.x; i < (a.nk - 1) * (DEG + 1) * (DEG + 1); i += blockDim.x) coef_s[i] = a.coef[i];
    __syncthreads();
    const double* knots = t_s + DEG;
    const uint32_t coef_u32 = smem_u32(coef_s);

    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long lo = warp * a.rows_per_warp;
    long long hi = lo + a.rows_per_warp;
    if (hi > a.n) hi = a.n;
    const int rec = (a.nk - 1) * C::SLOTS + 1;
    double* out = a.partials + warp * rec;

    double accG[C::NG], accH[C::NH];
#pragma unroll
    for (int q = 0; q < C::NG; ++q) accG[q] = 0.0;
#pragma unroll
    for (int q = 0; q < C::NH; ++q) accH[q] = 0.0;
    double objacc = 0.0;
    bool bad_domain = false;
    int cur = -1;

    auto flush = [&]() {
        // fixed-order butterfly over the 32 lane blocks, then lane q writes slot q
        double mine = 0.0;
#pragma unroll
        for (int q = 0; q < C::NG; ++q) {
            const double v = warp_sum(accG[q]);
            if (lane == q) mine = v;
            accG[q] = 0.0;
        }
        if (HESS) {
#pragma unroll
            for (int q = 0; q < C::NH; ++q) {
                const double v = warp_sum(accH[q]);
                if (lane == C::NG + q) mine = v;
                accH[q] = 0.0;
            }
        }
        if (lane < C::SLOTS) out[cur * C::SLOTS + lane] = mine;
    };

    // per-row work: interval lookup, basis, linear predictor, row terms
    // `idx` enters as the interval of the lane's previous row: x is sorted, so the search
    // clip(#(knots <= x) - 1, 0, nk-2) only ever moves forward (usually by 0 or 1 comparisons)
    struct RowIn {
        double x, ob, se, off, wt;
    };
    auto load = [&](long long i, bool active) {
        RowIn r;
        r.x = 0.0; r.ob = 0.0; r.se = 1.0; r.off = 0.0; r.wt = 1.0;
        if (active) {
            r.x = a.xs[i];
            r.ob = a.obs[i];
            if (a.se) r.se = a.se[i];
            if (a.offset) r.off = a.offset[i];
            if (a.weights) r.wt = a.weights[i];
        }
        return r;
    };
    auto compute = [&](const RowIn& in, bool active, int& idx, double* N, double& rr, double& ww) {
        if (!active) return;  // N, rr, ww of an inactive lane are never read
        const double x = in.x;
        while (idx < a.nk - 2 && knots[idx + 1] <= x) ++idx;
        const double u = x - knots[idx];
        const int cbase = idx * (DEG + 1) * (DEG + 1);
#pragma unroll
        for (int q = 0; q <= DEG; ++q) {
            double ck[DEG + 1];
            if (COEF_SMEM && (DEG + 1) % 2 == 0) {  // coefficient rows are 16-byte aligned: LDS.128
#pragma unroll
                for (int k = 0; k <= DEG; k += 2) {
                    const double2 v2 = lds128(coef_u32 + (cbase + q * (DEG + 1) + k) * 8);
                    ck[k] = v2.x;
                    ck[k + 1] = v2.y;
                }
            } else if (COEF_SMEM) {
#pragma unroll
                for (int k = 0; k <= DEG; ++k) ck[k] = lds64(coef_u32 + (cbase + q * (DEG + 1) + k) * 8);
            } else {
#pragma unroll
                for (int k = 0; k <= DEG; ++k) ck[k] = __ldg(a.coef + cbase + q * (DEG + 1) + k);
            }
            double v = ck[DEG];
#pragma unroll
            for (int k = DEG - 1; k >= 0; --k) v = fma(v, u, ck[k]);
            N[q] = v;
        }
        double y = in.off;
#pragma unroll
        for (int q = 0; q <= DEG; ++q) y = fma(N[q], beta_s[idx + q], y);
        const RowTerms1 rt = row_terms1<false>(a.family, a.link, a.fast != 0, y, in.ob, in.se);
        bad_domain |= !rt.ok;
        if (a.weights) {
            const double wt = in.wt;
            objacc += wt * rt.l;
            rr = wt * rt.r;
            ww = wt * rt.w;
        } else {
            objacc += rt.l;
            rr = rt.r;
            ww = rt.w;
        }
    };
    // lanes of one interval accumulate together; the warp walks the (few) distinct intervals in order
    auto add_row = [&](const double* N, double rr, double ww) {
#pragma unroll
        for (int q = 0; q <= DEG; ++q) accG[q] = fma(rr, N[q], accG[q]);
        if (HESS) {
#pragma unroll
            for (int p = 0; p <= DEG; ++p) {
                const double wp = ww * N[p];
#pragma unroll
                for (int q = 0; q <= p; ++q) accH[p * (p + 1) / 2 + q] = fma(wp, N[q], accH[p * (p + 1) / 2 + q]);
            }
        }
    };
    auto accumulate = [&](bool active, int idx, const double* N, double rr, double ww) {
        // common case: every active lane is still in the warp's current interval
        if (__all_sync(FULL_MASK, !active || idx == cur)) {
            if (active) add_row(N, rr, ww);
            return;
        }
        unsigned todo = __ballot_sync(FULL_MASK, active);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int v = __shfl_sync(FULL_MASK, idx, leader);
            const bool mine = active && idx == v;
            if (v != cur) {
                if (cur >= 0) flush();
                cur = v;
            }
            if (mine) add_row(N, rr, ww);
            todo &= ~__ballot_sync(FULL_MASK, mine);
        }
    };
    // The row loop is bound by the latency of its global loads (ncu: long-scoreboard stalls, FP64 pipe
    // 15 % busy), and 118 registers allow only 16 warps per SM: the inputs of the next two iterations
    // are fetched into registers before the current one is processed.
    int idx = 0;
    RowIn in0 = load(lo + lane, lo + lane < hi);
    RowIn in1 = load(lo + 32 + lane, lo + 32 + lane < hi);
    for (long long base = lo; base < hi; base += 32) {
        const long long i = base + lane;
        const bool active = i < hi;
        const RowIn in2 = load(i + 64, i + 64 < hi);
        double N[DEG + 1], rr, ww;
        compute(in0, active, idx, N, rr, ww);
        accumulate(active, idx, N, rr, ww);
        in0 = in1;
        in1 = in2;
    }
    if (cur >= 0) flush();
    const double o = warp_sum(objacc);
    if (lane == 0) out[rec - 1] = o;
    if (bad_domain) atomicOr(a.status, 1);
}

// Taylor coefficients of every basis piece about the left knot of its interval: one thread per
// (interval, derivative order)
template <int DEG>
__global__ void banded_coeff_kernel(const double* __restrict__ t, int nk, double* __restrict__ coef) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (nk - 1) * (DEG + 1)) return;
    const int idx = e / (DEG + 1), k = e % (DEG + 1);
    double N[DEG + 1];
    local_basis_ders<DEG>(t, idx + DEG, t[idx + DEG], k, N);
    double fact = 1.0;
    for (int m = 2; m <= k; ++m) fact *= (double)m;
#pragma unroll
    for (int q = 0; q <= DEG; ++q) coef[idx * (DEG + 1) * (DEG + 1) + q * (DEG + 1) + k] = N[q] / fact;
}

// partial records summed over warps in a fixed order: a 32 x 32 block owns 32 entries, row ty of the
// block walks the warps ty, ty + 32, ... (loads coalesce across the entries), then the 32 row sums are
// added in order.  (One thread per entry walking all 2368 records serially took 0.3 ms.)
__global__ void __launch_bounds__(1024) banded_reduce_kernel(const double* __restrict__ partials, long long nwarps, int rec,
                                                            double* __restrict__ reduced) {
    __shared__ double part[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int e = blockIdx.x * 32 + tx;
    double s0 = 0.0, s1 = 0.0;
    if (e < rec) {
        long long w = ty;
        for (; w + 32 < nwarps; w += 64) {
            s0 += partials[w * rec + e];
            s1 += partials[(w + 32) * rec + e];
        }
        if (w < nwarps) s0 += partials[w * rec + e];
    }
    part[ty][tx] = s0 + s1;
    __syncthreads();
    if (ty == 0 && e < rec) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) s += part[k][tx];
        reduced[e] = s;
    }
}

// band blocks -> packed [obj | grad(P) | hess(P x P)]: entry (i, j) gathers the intervals that hold both bases
__global__ void __launch_bounds__(256) banded_assemble_kernel(const double* __restrict__ reduced, int nint, int deg, int hess,
                                                             double* __restrict__ packed, int P) {
    const int ng = deg + 1, slots = ng + (deg + 1) * (deg + 2) / 2;
    const int total = 1 + P + (hess ? P * P : 0);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        if (e == 0) {
            packed[0] = reduced[nint * slots];
        } else if (e <= P) {
            const int i = e - 1;
            double s = 0.0;
            for (int idx = (i - deg > 0 ? i - deg : 0); idx <= i && idx < nint; ++idx) s += reduced[idx * slots + (i - idx)];
            packed[e] = s;
        } else {
            const int i = (e - 1 - P) / P, j = (e - 1 - P) % P;
            const int hi = i > j ? i : j, lo = i > j ? j : i;
            double s = 0.0;
            for (int idx = (hi - deg > 0 ? hi - deg : 0); idx <= lo && idx < nint; ++idx) {
                const int p = hi - idx, q = lo - idx;
                s += reduced[idx * slots + ng + p * (p + 1) / 2 + q];
            }
            packed[e] = s;
        }
    }
}

__global__ void iota_kernel(int32_t* p, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}

__global__ void gather_rows_kernel(double* __restrict__ dst, const double* __restrict__ src, const int32_t* __restrict__ perm, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[perm[i]];
}

}  // namespace anml
