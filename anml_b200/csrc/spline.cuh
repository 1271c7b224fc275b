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

// non-zero values of the `order`-th derivative of the degree-DEG basis at xe in span mu.
// Every loop has compile-time bounds (runtime limits are predicates), so N[] stays in
// registers; the operation order is the oracle's.
template <int DEG>
__device__ __forceinline__ void local_basis_ders(const double* __restrict__ t, int mu, double xe, int order, double* N /* [DEG+1] */) {
#pragma unroll
    for (int j = 0; j <= DEG; ++j) N[j] = 0.0;
    if (order > DEG) return;
    const int low = DEG - order;
    N[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= DEG; ++j) {
        if (j <= low) {
            double saved = 0.0;
#pragma unroll
            for (int r = 0; r < j; ++r) {
                const double right = __dsub_rn(t[mu + r + 1], xe);
                const double left = __dsub_rn(xe, t[mu + 1 - j + r]);
                const double temp = __ddiv_rn(N[r], __dadd_rn(right, left));
                N[r] = __dadd_rn(saved, __dmul_rn(right, temp));
                saved = __dmul_rn(left, temp);
            }
            N[j] = saved;
        }
    }
#pragma unroll
    for (int q = 1; q <= DEG; ++q) {
        if (q > low) {
            double nw[DEG + 1];
#pragma unroll
            for (int a = 0; a <= DEG; ++a) nw[a] = 0.0;
#pragma unroll
            for (int a = 0; a <= q; ++a) {
                const int j = mu - q + a;
                double val = 0.0;
                if (a >= 1) val = __dadd_rn(val, __ddiv_rn(N[a - 1], __dsub_rn(t[j + q], t[j])));
                if (a <= q - 1) val = __dsub_rn(val, __ddiv_rn(N[a], __dsub_rn(t[j + q + 1], t[j + 1])));
                nw[a] = __dmul_rn((double)q, val);
            }
#pragma unroll
            for (int a = 0; a <= DEG; ++a) N[a] = nw[a];
        }
    }
}

// One thread per row computes the (degree+1) non-zero basis values; the CTA stages its
// 256 rows in shared memory (row stride nb+1 doubles) and writes them out with
// consecutive threads on consecutive doubles, so the HBM writes are coalesced (a thread
// writing its own 8*nb-byte row directly reached only 12 % of the write bandwidth).
template <int DEG>
__global__ void __launch_bounds__(256) spline_design_kernel(const SplineArgs a) {
    extern __shared__ double sm_spline[];
    const int nt = a.nk + 2 * DEG;
    const int nb = a.nk - 1 + DEG;
    const int pitch = nb + 1;
    double* t_s = sm_spline;               // clamped knot vector
    double* tile = sm_spline + ((nt + 1) & ~1);  // [256][pitch]
    for (int i = threadIdx.x; i < nt; i += blockDim.x) t_s[i] = a.t[i];
    __syncthreads();
    const double* knots = t_s + DEG;  // knots[0..nk-1]
    const double k0 = knots[0], k1 = knots[a.nk - 1];

    const long long ntiles = (a.n + 255) / 256;
    for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const long long row = tl * 256 + threadIdx.x;
        double* mine = tile + threadIdx.x * pitch;
        for (int c = 0; c < nb; ++c) mine[c] = 0.0;
        if (row < a.n) {
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
            const int mu = idx + DEG;

            double N[DEG + 1];
            if (x < k0 || x > k1) {
                // Taylor polynomial of degree ldegree / rdegree about the end knot
                const bool left = x < k0;
                const double edge = left ? k0 : k1;
                const int ext = left ? a.ldegree : a.rdegree;
                const double dx = x - edge;
                double accv[DEG + 1];
#pragma unroll
                for (int j = 0; j <= DEG; ++j) accv[j] = 0.0;
                double pw = 1.0, fact = 1.0;  // dx^(m-order), (m-order)!
                for (int m = a.order; m <= ext; ++m) {
                    local_basis_ders<DEG>(t_s, mu, edge, m, N);
                    const double c = pw / fact;
#pragma unroll
                    for (int j = 0; j <= DEG; ++j) accv[j] = __dadd_rn(accv[j], __dmul_rn(N[j], c));
                    pw *= dx;
                    fact *= (double)(m - a.order + 1);
                }
#pragma unroll
                for (int j = 0; j <= DEG; ++j) N[j] = accv[j];
            } else {
                local_basis_ders<DEG>(t_s, mu, x, a.order, N);
            }
#pragma unroll
            for (int q = 0; q <= DEG; ++q) mine[idx + q] = N[q];
        }
        __syncthreads();
        // coalesced write-out of the tile: element e -> (row e / nb, column e % nb)
        const long long row0 = tl * 256;
        const int rows_here = (int)((a.n - row0) < 256 ? (a.n - row0) : 256);
        const int total = rows_here * nb;
        for (int e = threadIdx.x; e < total; e += 256) {
            const int r = e / nb, c = e - r * nb;
            a.out[(row0 + r) * a.ld + a.col0 + c] = tile[r * pitch + c];
        }
        __syncthreads();
    }
}

}  // namespace anml
