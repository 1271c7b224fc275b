// On-device B-spline design-matrix builder (knot-interval lookup + Cox-de Boor).
//
// Replaces XSpline.get_design_mat(x, order) as anml calls it
// (variable/spline.py:111-113, getter/prior.py:191-194).  The arithmetic mirrors
// oracle/anml_oracle.py::_local_basis_ders operation by operation with explicit
// round-to-nearest intrinsics (no FMA contraction), so basis values inside the
// knot span are bit-identical to the NumPy oracle, and the interval index is the
// same integer as searchsorted(knots, x, side="right") - 1 clipped to [0, nk-2].
#pragma once
#include "common.cuh"

namespace anml {

constexpr int SPLINE_MAX_DEGREE = 5;
constexpr int SPLINE_MAX_KNOTS = 1024;

struct SplineArgs {
    const double* x;   // device (n)
    long long n;
    const double* t;   // device: clamped knot vector, nk + 2*degree entries
    int nk, degree, ldegree, rdegree, order;
    double* out;       // device, row stride ld, columns [col0, col0 + nk - 1 + degree)
    long long ld;
    int col0;
    int32_t* idx_out;  // device (n) or null
};

// non-zero values of the `order`-th derivative of the degree-`deg` basis at xe in span mu
__device__ __forceinline__ void local_basis_ders(const double* __restrict__ t, int deg, int mu, double xe, int order,
                                                 double* N /* [SPLINE_MAX_DEGREE+1] */) {
#pragma unroll
    for (int j = 0; j <= SPLINE_MAX_DEGREE; ++j) N[j] = 0.0;
    if (order > deg) return;
    const int low = deg - order;
    N[0] = 1.0;
    for (int j = 1; j <= low; ++j) {
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            const double right = __dsub_rn(t[mu + r + 1], xe);
            const double left = __dsub_rn(xe, t[mu + 1 - j + r]);
            const double temp = __ddiv_rn(N[r], __dadd_rn(right, left));
            N[r] = __dadd_rn(saved, __dmul_rn(right, temp));
            saved = __dmul_rn(left, temp);
        }
        N[j] = saved;
    }
    for (int q = low + 1; q <= deg; ++q) {
        double nw[SPLINE_MAX_DEGREE + 1];
#pragma unroll
        for (int a = 0; a <= SPLINE_MAX_DEGREE; ++a) nw[a] = 0.0;
        for (int a = 0; a <= q; ++a) {
            const int j = mu - q + a;
            double val = 0.0;
            if (a >= 1) val = __dadd_rn(val, __ddiv_rn(N[a - 1], __dsub_rn(t[j + q], t[j])));
            if (a <= q - 1) val = __dsub_rn(val, __ddiv_rn(N[a], __dsub_rn(t[j + q + 1], t[j + 1])));
            nw[a] = __dmul_rn((double)q, val);
        }
#pragma unroll
        for (int a = 0; a <= SPLINE_MAX_DEGREE; ++a) N[a] = nw[a];
    }
}

__global__ void __launch_bounds__(256) spline_design_kernel(const SplineArgs a) {
    extern __shared__ double t_s[];  // clamped knot vector
    const int nt = a.nk + 2 * a.degree;
    for (int i = threadIdx.x; i < nt; i += blockDim.x) t_s[i] = a.t[i];
    __syncthreads();
    const double* knots = t_s + a.degree;  // knots[0..nk-1]
    const int nb = a.nk - 1 + a.degree;
    const double k0 = knots[0], k1 = knots[a.nk - 1];

    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < a.n;
         row += (long long)gridDim.x * blockDim.x) {
        const double x = a.x[row];
        // interval index: number of knots <= x, minus one, clipped to [0, nk-2]
        int lo = 0, hi = a.nk;  // first index with knots[idx] > x
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (knots[mid] <= x) lo = mid + 1; else hi = mid;
        }
        int idx = lo - 1;
        idx = idx < 0 ? 0 : (idx > a.nk - 2 ? a.nk - 2 : idx);
        if (a.idx_out) a.idx_out[row] = idx;
        const int mu = idx + a.degree;

        double N[SPLINE_MAX_DEGREE + 1];
        if (x < k0 || x > k1) {
            // Taylor polynomial of degree ldegree / rdegree about the end knot
            const bool left = x < k0;
            const double edge = left ? k0 : k1;
            const int ext = left ? a.ldegree : a.rdegree;
            const double dx = x - edge;
            double accv[SPLINE_MAX_DEGREE + 1];
#pragma unroll
            for (int j = 0; j <= SPLINE_MAX_DEGREE; ++j) accv[j] = 0.0;
            double pw = 1.0, fact = 1.0;  // dx^(m-order), (m-order)!
            for (int m = a.order; m <= ext; ++m) {
                local_basis_ders(t_s, a.degree, mu, edge, m, N);
                const double c = pw / fact;
#pragma unroll
                for (int j = 0; j <= SPLINE_MAX_DEGREE; ++j) accv[j] = __dadd_rn(accv[j], __dmul_rn(N[j], c));
                pw *= dx;
                fact *= (double)(m - a.order + 1);
            }
#pragma unroll
            for (int j = 0; j <= SPLINE_MAX_DEGREE; ++j) N[j] = accv[j];
        } else {
            local_basis_ders(t_s, a.degree, mu, x, a.order, N);
        }
        double* o = a.out + row * a.ld + a.col0;
        for (int c = 0; c < nb; ++c) {
            const int j = c - idx;
            double v = 0.0;
#pragma unroll
            for (int q = 0; q <= SPLINE_MAX_DEGREE; ++q) v = (j == q && q <= a.degree) ? N[q] : v;
            o[c] = v;
        }
    }
}

}  // namespace anml
