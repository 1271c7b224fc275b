// Streaming kernels of the general (multi-pass) path and of Parameter.get_params:
//   linpred_kernel   y = X beta (+ offset), optionally through the link at one order
//   rowterms_kernel  per-row family terms from the linear predictors (K = 1 or 2)
//   xtr_kernel       g = X^T r for any width
//   jacobian / tensor kernels for get_params(order = 1, 2)
//   add_priors_kernel  Gaussian prior terms on the packed result
// All are HBM-bound: coalesced 16-byte loads, fixed-order reductions.
#pragma once
#include <math_constants.h>
#include "common.cuh"
#include "rowmath.cuh"

namespace anml {

// One sub-warp group of GS lanes per row; lane l of the group covers column
// pairs l, l+GS, ... of the (even-ld) row.
template <int GS>
__global__ void __launch_bounds__(256) linpred_kernel(const double* __restrict__ X, long long n, int p, long long ld,
                                                      const double* __restrict__ beta, const double* __restrict__ offset,
                                                      int link, int order, double* __restrict__ out, int* status) {
    extern __shared__ double beta_s[];  // ld doubles, zero beyond p
    for (int i = threadIdx.x; i < (int)ld; i += blockDim.x) beta_s[i] = (i < p) ? beta[i] : 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int lg = lane & (GS - 1);
    constexpr int GPW = 32 / GS;  // groups per warp
    const long long group = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + lane / GS;
    const long long ngroups = (long long)gridDim.x * (blockDim.x >> 5) * GPW;
    const int npairs = (int)(ld >> 1);
    bool bad = false;
    const long long rounds = (n + ngroups - 1) / ngroups;  // keep the warp converged for the shuffles
    for (long long it = 0; it < rounds; ++it) {
        const long long row = group + it * ngroups;
        double s = 0.0;
        if (row < n) {
            const double* xr = X + row * ld;
            for (int c = lg; c < npairs; c += GS) {
                const double2 v = ldg_stream2(xr + 2 * c);
                s = fma(v.x, beta_s[2 * c], s);
                s = fma(v.y, beta_s[2 * c + 1], s);
            }
        }
#pragma unroll
        for (int m = GS >> 1; m >= 1; m >>= 1) s += shfl_xor_d(s, m);
        if (row < n && lg == 0) {
            if (offset) s += offset[row];
            if (order >= 0) {
                bool ok;
                s = link_order(link, s, order, ok);
                bad |= !ok;
            }
            out[row] = s;
        }
    }
    if (bad) atomicOr(status, 1);
}

// out[i][j] = z[i] * X[i][j]   (Parameter.get_params order 1, parameter/main.py:276-277)
__global__ void __launch_bounds__(256) jacobian_kernel(const double* __restrict__ X, long long n, int p, long long ld,
                                                       const double* __restrict__ z, double* __restrict__ out) {
    const long long total = n * p;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / p;
        const int j = (int)(e - i * p);
        out[e] = z[i] * X[i * ld + j];
    }
}

// out[i][j][k] = z[i] * (X[i][j] * X[i][k])   (order 2, parameter/main.py:278-280)
__global__ void __launch_bounds__(256) tensor2_kernel(const double* __restrict__ X, long long n, int p, long long ld,
                                                      const double* __restrict__ z, double* __restrict__ out) {
    const long long pp = (long long)p * p;
    const long long total = n * pp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / pp;
        const int jk = (int)(e - i * pp);
        const int j = jk / p, k = jk - j * p;
        out[e] = z[i] * (X[i * ld + j] * X[i * ld + k]);
    }
}

// Per-row family terms for the general path.  y0/y1: linear predictors (offset
// included).  Writes r_k and w_kl vectors and one objective partial per block.
struct RowTermsArgs {
    long long n;
    int family, nparams, link0, link1, fast;
    const double *y0, *y1, *obs, *se, *weights;
    double *r0, *r1, *w00, *w01, *w11;
    double* obj_partials;  // [gridDim.x]
    int* status;
};

__global__ void __launch_bounds__(256) rowterms_kernel(const RowTermsArgs a) {
    __shared__ double red[8];
    double l = 0.0;
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        if (a.nparams == 1) {
            const RowTerms1 t = row_terms1(a.family, a.link0, a.fast != 0, a.y0[i], a.obs[i], a.se ? a.se[i] : 1.0);
            const double wt = a.weights ? a.weights[i] : 1.0;
            l += wt * t.l;
            a.r0[i] = wt * t.r;
            a.w00[i] = wt * t.w;
            bad |= !t.ok;
        } else {
            const RowTerms2 t = row_terms2(a.link0, a.link1, a.y0[i], a.y1[i], a.obs[i], a.fast != 0);
            const double wt = a.weights ? a.weights[i] : 1.0;
            l += wt * t.l;
            a.r0[i] = wt * t.r0; a.r1[i] = wt * t.r1;
            a.w00[i] = wt * t.w00; a.w01[i] = wt * t.w01; a.w11[i] = wt * t.w11;
            bad |= !t.ok;
        }
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        a.obj_partials[blockIdx.x] = s;
    }
    if (bad) atomicOr(a.status, 1);
}

__global__ void sum_partials_kernel(const double* __restrict__ partials, int count, double* __restrict__ dst, int accumulate) {
    // single thread block, fixed order
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += partials[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x >> 1; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (accumulate) dst[0] += red[0]; else dst[0] = red[0];
    }
}

// g = X^T r for any p.  grid = (row CTAs, 64-column chunks).  Each warp owns one
// 64-column chunk (one double2 per lane) and strides over rows with 8 independent
// 512-byte loads in flight; per-CTA partials, column_sum_kernel finishes (fixed order).
__global__ void __launch_bounds__(256) xtr_kernel(const double* __restrict__ X, long long n, long long ld,
                                                  const double* __restrict__ r, double* __restrict__ partials) {
    __shared__ double gs[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int c = blockIdx.y * 32 + lane;  // column pair
    const bool live = 2 * c < ld;
    double ax = 0.0, ay = 0.0;
    const long long w0 = (long long)blockIdx.x * nw + warp, wstride = (long long)gridDim.x * nw;
    long long row = w0;
    if (live) {
        const double* xp = X + 2 * c;
        for (; row + 7 * wstride < n; row += 8 * wstride) {
            double2 v[8];
            double rv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                v[u] = ldg_stream2(xp + (row + u * wstride) * ld);
                rv[u] = r[row + u * wstride];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                ax = fma(rv[u], v[u].x, ax);
                ay = fma(rv[u], v[u].y, ay);
            }
        }
        for (; row < n; row += wstride) {
            const double rv = r[row];
            const double2 v = ldg_stream2(xp + row * ld);
            ax = fma(rv, v.x, ax);
            ay = fma(rv, v.y, ay);
        }
    }
    gs[warp][2 * lane] = ax;
    gs[warp][2 * lane + 1] = ay;
    __syncthreads();
    if (threadIdx.x < 64) {
        const int col = blockIdx.y * 64 + threadIdx.x;
        if (col < ld) {
            double s = 0.0;
            for (int w = 0; w < nw; ++w) s += gs[w][threadIdx.x];
            partials[(size_t)blockIdx.x * ld + col] = s;
        }
    }
}

// dst[c] = sum_b partials[b][c], c < count (fixed order)
__global__ void __launch_bounds__(256) column_sum_kernel(const double* __restrict__ partials, int nblocks, long long stride,
                                                         int count, double* __restrict__ dst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * stride + c];
    dst[c] = s;
}

// Gaussian prior terms of one Parameter added to the packed record
// (GaussianPrior.objective/gradient/hessian, prior/main.py:178-197, summed as
// Parameter.prior_* do, parameter/main.py:296-339).
struct PriorArgs {
    int p, off, P_total, m, want;
    const double* beta;         // device, all coefficients
    const double* mean;         // (p)
    const double* inv_var;      // (p)  1/sd^2, 0 where sd = inf
    const double* lin_mat;      // (m, p)
    const double* lin_mean;     // (m)
    const double* lin_inv_var;  // (m)
    const double* lin_hess;     // (p, p) = (M^T / sd^2) M, precomputed
    double* packed;
};

__global__ void __launch_bounds__(256) add_priors_kernel(const PriorArgs a) {
    extern __shared__ double res_s[];  // m scaled residuals
    __shared__ double red[256];
    const double* b = a.beta + a.off;
    double* grad = a.packed + 1 + a.off;
    double* H = a.packed + 1 + a.P_total;
    double o = 0.0;
    for (int j = threadIdx.x; j < a.p; j += blockDim.x) {
        const double d = b[j] - a.mean[j];
        o += 0.5 * d * d * a.inv_var[j];
    }
    for (int i = threadIdx.x; i < a.m; i += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < a.p; ++j) s = fma(a.lin_mat[(size_t)i * a.p + j], b[j], s);
        s -= a.lin_mean[i];
        o += 0.5 * s * s * a.lin_inv_var[i];
        res_s[i] = s * a.lin_inv_var[i];
    }
    red[threadIdx.x] = o;
    __syncthreads();
    for (int w = blockDim.x >> 1; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0 && (a.want & 1)) a.packed[0] += red[0];
    if (a.want & 2) {
        for (int j = threadIdx.x; j < a.p; j += blockDim.x) {
            double s = (b[j] - a.mean[j]) * a.inv_var[j];
            for (int i = 0; i < a.m; ++i) s = fma(a.lin_mat[(size_t)i * a.p + j], res_s[i], s);
            grad[j] += s;
        }
    }
    if (a.want & 4) {
        for (int e = threadIdx.x; e < a.p * a.p; e += blockDim.x) {
            const int r = e / a.p, c = e - r * a.p;
            double v = (a.m > 0) ? a.lin_hess[e] : 0.0;
            if (r == c) v += a.inv_var[r];
            H[(size_t)(a.off + r) * a.P_total + a.off + c] += v;
        }
    }
}

// Column validators on the device (reference: src/anml/data/validator.py:21-34):
// flags[c] |= 1 if column c holds a NaN, |= 2 if it holds a value <= 0.
// A thread keeps one column (consecutive threads = consecutive columns of a row, so the loads
// coalesce) and ORs its findings in a register: one atomic per thread at the end, not one per
// offending element (half of a standard-normal column is <= 0).
__global__ void __launch_bounds__(256) validate_columns_kernel(const double* __restrict__ X, long long n, long long ld, int col0,
                                                               int ncols, int* __restrict__ flags) {
    for (int cbase = 0; cbase < ncols; cbase += 256) {
        const int cw = ncols - cbase < 256 ? ncols - cbase : 256;
        const int rows_per_pass = 256 / cw;
        const int c = threadIdx.x % cw, rsub = threadIdx.x / cw;
        int f = 0;
        if (rsub < rows_per_pass) {
            const double* col = X + col0 + cbase + c;
            for (long long i = (long long)blockIdx.x * rows_per_pass + rsub; i < n; i += (long long)gridDim.x * rows_per_pass) {
                const double v = col[i * ld];
                if (v != v) f |= 1;
                if (v <= 0.0) f |= 2;
            }
        }
        if (f) atomicOr(flags + cbase + c, f);
    }
}

// per-block (min, max, saw-a-NaN) of a vector; the few hundred block records are finished on the host
__global__ void __launch_bounds__(256) minmax_kernel(const double* __restrict__ x, long long n, double* __restrict__ block_min,
                                                     double* __restrict__ block_max, int* __restrict__ block_nan) {
    double lo = CUDART_INF, hi = -CUDART_INF;
    int nan = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = x[i];
        nan |= (v != v);
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
    }
    __shared__ double s_lo[8], s_hi[8];
    __shared__ int s_nan[8];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        s_lo[w] = lo;
        s_hi[w] = hi;
        s_nan[w] = nan;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) {
            lo = fmin(lo, s_lo[k]);
            hi = fmax(hi, s_hi[k]);
            nan |= s_nan[k];
        }
        block_min[blockIdx.x] = lo;
        block_max[blockIdx.x] = hi;
        block_nan[blockIdx.x] = nan;
    }
}

__global__ void gather_kernel(const double* __restrict__ sorted, const long long* __restrict__ ranks, int n_ranks, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_ranks) out[i] = sorted[ranks[i]];
}

__global__ void fill_kernel(double* p, long long n, double v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// publish the Log/Logit domain flag of this evaluation and re-arm the status word for the next one
__global__ void set_flag_kernel(double* dst, int* status) {
    dst[0] = (double)(*status);
    *status = 0;
}

// scatter `ncols` columns from src (row stride src_ld) into X columns [col0, col0+ncols)
__global__ void __launch_bounds__(256) scatter_columns_kernel(double* __restrict__ X, long long n, long long ld, int col0,
                                                              int ncols, const double* __restrict__ src, long long src_ld) {
    const long long total = n * ncols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / ncols;
        const int c = (int)(e - i * ncols);
        X[i * ld + col0 + c] = src[i * src_ld + c];
    }
}

}  // namespace anml
