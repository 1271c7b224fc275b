// Epilogue of the fused evaluators in ONE launch: per-CTA partial records -> packed result
//     packed = [obj | grad(P_total) | hess(P_total, P_total) row-major | domain flag]
// It replaces the round-1 chain reduce_partials -> assemble -> add_priors -> set_flag (four serial
// launches, ~36 us of device time plus launch gaps; at 8 GPUs that chain was half of the
// per-evaluation fixed cost).
//
// The partial record of fused_hess_kernel / fused_eval_kernel is a list of 32-lane rows:
//     2*NT Hessian rows (tile, reg) | NB8 gradient rows | 1 objective row.
// One CTA of 4 warps owns one row: warp s sums the CTA records c = s, s+4, ... (8 loads in flight),
// the four sums are combined in a fixed order (bitwise reproducible for a given grid), and warp 0
// undoes the fragment/tile layout straight into `packed`:
//   * Hessian rows: the element lands at (row_off + c1, col_off + c2) and mirrored; the Gaussian
//     prior's Hessian (diag(1/sd^2) + (M^T/sd^2) M, prior/main.py:193-197) is added at the write;
//   * gradient rows: 4 lanes per column are summed, the prior gradient (prior/main.py:185-190) added;
//   * objective row: 32 lanes summed, the prior objectives (prior/main.py:178-183) added, and the
//     Log/Logit domain flag is published (and re-armed for the next evaluation).
// Prior sums follow Parameter.prior_* (parameter/main.py:296-339): direct + linear Gaussian terms.
#pragma once
#include "common.cuh"

namespace anml {

struct PriorDev {
    int on;                     // 0: contributes nothing
    int p, off, m;              // width, first coefficient in x, linear rows
    const double* mean;         // (p)
    const double* inv_var;      // (p)  1/sd^2, 0 where sd = inf
    const double* lin_mat;      // (m, p)
    const double* lin_mean;     // (m)
    const double* lin_inv_var;  // (m)
    const double* lin_hess;     // (p, p) = (M^T / sd^2) M, exactly symmetric
};

struct FinishArgs {
    const double* partials;  // [ncta][el]
    int ncta, el;
    int nb16, hess, p;       // layout of the record; p = live columns
    double* packed;
    int P_total, row_off, col_off;
    int write_grad;          // gradient rows -> packed[1 + row_off + c]
    int obj_mode;            // 0: leave packed[0], 1: packed[0] = s + prior objectives, 2: packed[0] += s
    int write_flag;          // publish the domain flag and reset the status word
    const double* beta;      // all coefficients (device)
    PriorDev prior;          // prior of the parameter that owns the (row_off) block: gradient and, on a
                             // diagonal block, Hessian terms
    PriorDev prior_obj[2];   // priors whose objective is added under obj_mode 1
    int n_prior_obj;
    int* status;
};

// scaled residuals of a linear Gaussian prior: res[i] = (M_i . b - mean_i) / sd_i^2 (all threads of the CTA)
__device__ __forceinline__ void prior_residuals(const PriorDev& pr, const double* __restrict__ beta, double* res_s) {
    const double* b = beta + pr.off;
    for (int i = threadIdx.x; i < pr.m; i += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < pr.p; ++j) s = fma(pr.lin_mat[(size_t)i * pr.p + j], b[j], s);
        res_s[i] = (s - pr.lin_mean[i]) * pr.lin_inv_var[i];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128) finish_fused_kernel(const FinishArgs a) {
    extern __shared__ double res_s[];  // linear-prior residuals (gradient / objective rows)
    __shared__ double part[4][32];
    __shared__ double red[128];
    const int lane = threadIdx.x & 31, split = threadIdx.x >> 5;
    const int row = blockIdx.x;
    const int e = row * 32 + lane;
    {
        double s[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] = 0.0;
        int c = split;
        for (; c + 28 < a.ncta; c += 32) {
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] += a.partials[(size_t)(c + 4 * u) * a.el + e];
        }
        double tail = 0.0;
        for (; c < a.ncta; c += 4) tail += a.partials[(size_t)c * a.el + e];
        part[split][lane] = (((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]))) + tail;
    }
    __syncthreads();
    double v = (part[0][lane] + part[1][lane]) + (part[2][lane] + part[3][lane]);

    const int NB8 = 2 * a.nb16, NT = NB8 * (NB8 + 1) / 2;
    const int goff = a.hess ? 2 * NT : 0;
    double* H = a.packed + 1 + a.P_total;

    if (row < goff) {
        // ---- Hessian row: tile (I, J <= I), accumulator register reg
        if (split != 0) return;
        const int reg = row & 1, tile = row >> 1;
        int I = 0;  // tile = I*(I+1)/2 + J
        while ((I + 1) * (I + 2) / 2 <= tile) ++I;
        const int J = tile - I * (I + 1) / 2;
        const int m = lane >> 2, nn = 2 * (lane & 3) + reg;
        if (I == J && m < nn) return;  // one triangle of the diagonal tiles: exact symmetry
        const int c1 = 16 * (I >> 1) + 2 * m + (I & 1);
        const int c2 = 16 * (J >> 1) + 2 * nn + (J & 1);
        if (c1 >= a.p || c2 >= a.p) return;
        if (a.prior.on && a.row_off == a.col_off) {
            if (a.prior.m > 0) v += a.prior.lin_hess[(size_t)c1 * a.p + c2];
            if (c1 == c2) v += a.prior.inv_var[c1];
        }
        H[(size_t)(a.row_off + c1) * a.P_total + a.col_off + c2] = v;
        H[(size_t)(a.row_off + c2) * a.P_total + a.col_off + c1] = v;
        if (a.row_off != a.col_off) {
            H[(size_t)(a.col_off + c2) * a.P_total + a.row_off + c1] = v;
            H[(size_t)(a.col_off + c1) * a.P_total + a.row_off + c2] = v;
        }
        return;
    }
    if (row < goff + NB8) {
        // ---- gradient row jb: column c = 16*(jb>>1) + 2*g + (jb&1), summed over the 4 lanes t
        if (!a.write_grad) return;
        const bool lin = a.prior.on && a.prior.m > 0;
        if (lin) prior_residuals(a.prior, a.beta, res_s);
        if (split != 0) return;
        v += shfl_xor_d(v, 1);
        v += shfl_xor_d(v, 2);
        const int jb = row - goff;
        const int c = 16 * (jb >> 1) + 2 * (lane >> 2) + (jb & 1);
        if ((lane & 3) != 0 || c >= a.p) return;
        if (a.prior.on) {
            double s = (a.beta[a.prior.off + c] - a.prior.mean[c]) * a.prior.inv_var[c];
            if (lin)
                for (int i = 0; i < a.prior.m; ++i) s = fma(a.prior.lin_mat[(size_t)i * a.prior.p + c], res_s[i], s);
            v += s;
        }
        a.packed[1 + a.row_off + c] = v;
        return;
    }
    // ---- objective row (+ prior objectives, + domain flag)
    double total = warp_sum(v);  // meaningful in warp 0
    if (a.obj_mode == 1) {
        for (int k = 0; k < a.n_prior_obj; ++k) {
            const PriorDev& pr = a.prior_obj[k];
            if (!pr.on) continue;
            const double* b = a.beta + pr.off;
            double o = 0.0;
            for (int j = threadIdx.x; j < pr.p; j += blockDim.x) {
                const double d = b[j] - pr.mean[j];
                o += 0.5 * d * d * pr.inv_var[j];
            }
            for (int i = threadIdx.x; i < pr.m; i += blockDim.x) {
                double s = 0.0;
                for (int j = 0; j < pr.p; ++j) s = fma(pr.lin_mat[(size_t)i * pr.p + j], b[j], s);
                s -= pr.lin_mean[i];
                o += 0.5 * s * s * pr.lin_inv_var[i];
            }
            __syncthreads();  // red is reused per prior
            red[threadIdx.x] = o;
            __syncthreads();
            for (int w = 64; w >= 1; w >>= 1) {
                if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
                __syncthreads();
            }
            total += red[0];
        }
    }
    if (threadIdx.x == 0) {
        if (a.obj_mode == 1) a.packed[0] = total;
        else if (a.obj_mode == 2) a.packed[0] += total;
        if (a.write_flag) {
            a.packed[1 + a.P_total + (size_t)a.P_total * a.P_total] = (double)(*a.status);
            *a.status = 0;  // re-armed for the next evaluation on this stream
        }
    }
}

}  // namespace anml
