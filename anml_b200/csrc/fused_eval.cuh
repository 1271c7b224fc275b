// Fused single-pass evaluator for one Parameter with p <= 64 design columns.
//
// One pass over the row-major FP64 design matrix X (n, p) produces
//     obj  = sum_i l_i,   g = X^T r,   H = X^T diag(w) X
// where (l, r, w) are the per-row NLL terms chained through the SmoothMapping
// (rowmath.cuh).  This replaces the reference's three materialised objects
// theta / J (n,p) / T (n,p,p) of Parameter.get_params (parameter/main.py:265-280).
//
// Work decomposition (B200: 148 SMs, one persistent CTA of 8 warps per SM):
//   * every warp is an independent worker: it owns 32-row blocks b = gw, gw+GW, ...
//     of X, a private ring of NS shared-memory slots (16 rows x 16*NB16 columns
//     each) and private accumulators; there is no CTA-level sync in the main loop;
//   * slots are filled by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle, box =
//     16 rows x 16 columns, zero fill outside the matrix) issued by the warp's
//     lane 0 as soon as the slot has been consumed, completion on an mbarrier;
//   * phase 1 (per 32-row block): y = X beta by FMA on MMA-layout fragments and
//     an 8-lane transposing reduction, so that each lane owns one row; the lane
//     evaluates link + family terms for its row once;
//   * phase 2: for every 4-row k-step the fragments a[j] = X[k][col_j] serve as
//     A (= X^T) and, scaled by w_k, as B of DMMA.8x8x4; only the 8x8 tiles on or
//     below the diagonal are computed (NB8*(NB8+1)/2 of NB8^2);  g += r_k a[j].
//   * epilogue: warps reduce through the (now idle) ring memory, each CTA writes
//     one partial record; finish_fused_kernel (finish.cuh) sums the records in a fixed
//     order and undoes the layout (no floating-point atomics -> bitwise reproducible
//     for a given grid).
//
// Fragment/column mapping: lane (g = lane>>2, t = lane&3) loads, for 16-column
// group j, the double2 at columns (16j + 2g, 16j + 2g + 1) of row k(t).  8-column
// MMA block jb = 2j + e therefore holds matrix columns col(jb, idx) = 16*(jb>>1) +
// 2*idx + (jb&1).  With the 128-byte swizzle the 16-byte chunk index is
// g ^ (row & 7); k-step ss of a slot uses rows 2t + (ss&1) + 8*(ss>>1), which
// makes every quarter-warp LDS.128 hit 8 distinct chunks: conflict-free.
#pragma once
#include "common.cuh"
#include "rowmath.cuh"

namespace anml {

struct FusedArgs {
    long long n;        // rows
    int p;              // design columns (<= 16*NB16)
    int link, family, fast;
    const double* beta;    // device, >= 16*NB16 doubles, zero beyond p
    const double* obs;     // device (n)            [MODE 0]
    const double* se;      // device (n) or null    [MODE 0]
    const double* offset;  // device (n) or null    [MODE 0]
    const double* weights; // device (n) or null: per-row likelihood weights (trimming)
    const double* rvec;    // device (n)            [MODE 1]
    const double* wvec;    // device (n)            [MODE 1]
    double* partials;      // [gridDim.x][EL]
    int* status;           // domain-violation flag
    // MODE 2: two parameters over the same design (mean/variance model): second
    // coefficient vector / link / offset, and the per-row vectors to write
    const double* beta1;
    const double* offset1;
    int link1;
    double *out_r0, *out_r1, *out_w00, *out_w01, *out_w11;
};

// 8 partial sums per lane -> one complete sum per lane: lane (g,t) keeps entry s == g,
// summed over the 8 lanes that share t (3 shuffle stages, 7 exchanges).
__device__ __forceinline__ double transpose_reduce8(const double* part, int g) {
    double k4[4], k2[2];
    {
        const bool hi = (g & 4) != 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double send = hi ? part[q] : part[q + 4];
            const double keep = hi ? part[q + 4] : part[q];
            k4[q] = keep + shfl_xor_d(send, 16);
        }
    }
    {
        const bool hi = (g & 2) != 0;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double send = hi ? k4[q] : k4[q + 2];
            const double keep = hi ? k4[q + 2] : k4[q];
            k2[q] = keep + shfl_xor_d(send, 8);
        }
    }
    const bool hi = (g & 1) != 0;
    const double send = hi ? k2[0] : k2[1];
    const double keep = hi ? k2[1] : k2[0];
    return keep + shfl_xor_d(send, 4);
}

template <int NB16>
struct FusedCfg {
    static constexpr int WARPS = 8;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int NB8 = 2 * NB16;
    static constexpr int NT = NB8 * (NB8 + 1) / 2;  // 8x8 tiles on/below the diagonal
    static constexpr int SLOT_ROWS = 16;
    static constexpr int SLOT_BYTES = NB16 * 2048;  // NB16 boxes of 16 rows x 128 B
    static constexpr int NS = (24576 / SLOT_BYTES) > 8 ? 8 : (24576 / SLOT_BYTES);
    static constexpr int RING_BYTES = NS * SLOT_BYTES;  // per warp
    static constexpr int EL_HESS = (2 * NT + NB8 + 1) * 32;  // doubles per CTA partial record
    static constexpr int EL_GRAD = (NB8 + 1) * 32;
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + WARPS * RING_BYTES + WARPS * NS * 8 + 2 * 16 * NB16 * 8;
};

template <int NB16, bool HESS, int MODE>
__global__ void __launch_bounds__(256, 1) fused_eval_kernel(const __grid_constant__ CUtensorMap tmap, const FusedArgs a) {
    using C = FusedCfg<NB16>;
    constexpr int NB8 = C::NB8, NT = C::NT, NS = C::NS, WARPS = C::WARPS;
    constexpr int EL = HESS ? C::EL_HESS : C::EL_GRAD;

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* ring_all = smem;
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(smem + WARPS * C::RING_BYTES);
    double* beta_s = reinterpret_cast<double*>(bars_all + WARPS * NS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < WARPS * NS; ++i) mbar_init(smem_u32(bars_all + i), 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 16 * NB16) {
        beta_s[threadIdx.x] = (MODE == 0 || MODE == 2) ? a.beta[threadIdx.x] : 0.0;
        beta_s[16 * NB16 + threadIdx.x] = (MODE == 2) ? a.beta1[threadIdx.x] : 0.0;
    }
    __syncthreads();

    unsigned char* ring = ring_all + warp * C::RING_BYTES;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t bars_u32 = smem_u32(bars_all + warp * NS);

    // rows handed out in 32-row blocks, round-robin over all warps of the grid
    const long long nblk = (a.n + 31) >> 5;
    const long long gw = (long long)blockIdx.x * WARPS + warp;
    const long long GW = (long long)gridDim.x * WARPS;
    const long long nmine = (nblk > gw) ? (nblk - gw + GW - 1) / GW : 0;
    const long long nslots = 2 * nmine;

    // producer state (meaningful in lane 0): next slot sequence number to request
    long long q_issue = 0;
    int ring_issue = 0;
    auto issue_next = [&]() {
        // slot q covers rows [(gw + (q>>1)*GW)*32 + (q&1)*16, +16)
        const long long row0 = (gw + (q_issue >> 1) * GW) * 32 + (q_issue & 1) * 16;
        const uint32_t bar = bars_u32 + ring_issue * 8;
        const uint32_t dst = ring_u32 + ring_issue * C::SLOT_BYTES;
        mbar_arrive_expect_tx(bar, C::SLOT_BYTES);
#pragma unroll
        for (int j = 0; j < NB16; ++j) tma_load_2d(dst + j * 2048, &tmap, 16 * j, (int)row0, bar);
        ++q_issue;
        ring_issue = (ring_issue + 1 == NS) ? 0 : ring_issue + 1;
    };
    if (lane == 0) {
        const int pre = nslots < NS ? (int)nslots : NS;
        for (int i = 0; i < pre; ++i) issue_next();
    }

    double acc[HESS ? 2 * NT : 2];
    double gacc[NB8];
    double objacc = 0.0;
#pragma unroll
    for (int i = 0; i < (HESS ? 2 * NT : 2); ++i) acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NB8; ++i) gacc[i] = 0.0;

    uint32_t phase_bits = 0;
    int ring_use = 0;  // ring index of the block's first slot
    bool bad_domain = false;

    // lane (g,t) owns row R of each 32-row block after the transposing reduction
    const int ss_own = g & 3;
    const int R = 16 * (g >> 2) + 2 * t + (ss_own & 1) + 8 * (ss_own >> 1);

    for (long long i = 0; i < nmine; ++i) {
        const int s0 = ring_use;
        const int s1 = (s0 + 1 == NS) ? 0 : s0 + 1;
        mbar_wait(bars_u32 + s0 * 8, (phase_bits >> s0) & 1u);
        mbar_wait(bars_u32 + s1 * 8, (phase_bits >> s1) & 1u);
        phase_bits ^= (1u << s0) | (1u << s1);
        const unsigned char* sp0 = ring + s0 * C::SLOT_BYTES;
        const unsigned char* sp1 = ring + s1 * C::SLOT_BYTES;

        const long long row = (gw + i * GW) * 32 + R;
        const bool valid = row < a.n;
        double rr = 0.0, ww = 0.0;

        if (MODE == 0) {
            // ---- phase 1: linear predictor for the 32 rows of this block ----
            double part[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const unsigned char* sp = (s < 4) ? sp0 : sp1;
                const int ss = s & 3;
                const int xr = 2 * t + (ss & 1);
                const unsigned char* rp = sp + (xr + 8 * (ss >> 1)) * 128 + ((g ^ xr) << 4);
                double d = 0.0;
#pragma unroll
                for (int j = 0; j < NB16; ++j) {
                    const double2 v = *reinterpret_cast<const double2*>(rp + j * 2048);
                    const double2 b = *reinterpret_cast<const double2*>(beta_s + 16 * j + 2 * g);
                    d = fma(v.x, b.x, d);
                    d = fma(v.y, b.y, d);
                }
                part[s] = d;
            }
            // transposing reduction over the 8 lanes that share t: lane (g,t) keeps k-step s == g
            double y = transpose_reduce8(part, g);
            if (valid) {
                if (a.offset) y += a.offset[row];
                const double ob = a.obs[row];
                const double se = a.se ? a.se[row] : 1.0;
                const RowTerms1 rt = row_terms1(a.family, a.link, a.fast != 0, y, ob, se);
                const double wt = a.weights ? a.weights[row] : 1.0;
                objacc += wt * rt.l;
                rr = wt * rt.r;
                ww = wt * rt.w;
                bad_domain |= !rt.ok;
            }
        } else if (MODE == 2) {
            // ---- two linear predictors over the same rows (mean / variance model) ----
            double part0[8], part1[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const unsigned char* sp = (s < 4) ? sp0 : sp1;
                const int ss = s & 3;
                const int xr = 2 * t + (ss & 1);
                const unsigned char* rp = sp + (xr + 8 * (ss >> 1)) * 128 + ((g ^ xr) << 4);
                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int j = 0; j < NB16; ++j) {
                    const double2 v = *reinterpret_cast<const double2*>(rp + j * 2048);
                    const double2 b0 = *reinterpret_cast<const double2*>(beta_s + 16 * j + 2 * g);
                    const double2 b1 = *reinterpret_cast<const double2*>(beta_s + 16 * NB16 + 16 * j + 2 * g);
                    d0 = fma(v.x, b0.x, d0);
                    d0 = fma(v.y, b0.y, d0);
                    d1 = fma(v.x, b1.x, d1);
                    d1 = fma(v.y, b1.y, d1);
                }
                part0[s] = d0;
                part1[s] = d1;
            }
            double y0 = transpose_reduce8(part0, g);
            double y1 = transpose_reduce8(part1, g);
            if (valid) {
                if (a.offset) y0 += a.offset[row];
                if (a.offset1) y1 += a.offset1[row];
                const RowTerms2 rt = row_terms2(a.link, a.link1, y0, y1, a.obs[row], a.fast != 0);
                const double wt = a.weights ? a.weights[row] : 1.0;
                objacc += wt * rt.l;
                a.out_r0[row] = wt * rt.r0;
                a.out_r1[row] = wt * rt.r1;
                a.out_w00[row] = wt * rt.w00;
                a.out_w01[row] = wt * rt.w01;
                a.out_w11[row] = wt * rt.w11;
                bad_domain |= !rt.ok;
            }
            // no gradient / Hessian here: both slots are consumed, re-arm them
            __syncwarp();
            if (lane == 0 && q_issue < nslots) issue_next();
            if (lane == 0 && q_issue < nslots) issue_next();
            ring_use = (s1 + 1 == NS) ? 0 : s1 + 1;
            continue;
        } else {
            if (valid) {
                if (a.rvec) rr = a.rvec[row];  // (null: cross block of two parameters, Hessian only)
                ww = HESS ? a.wvec[row] : 0.0;
            }
        }

        // ---- phase 2: gradient FMAs + DMMA tiles, one 16-row slot at a time ----
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const unsigned char* sp = h ? sp1 : sp0;
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) {
                const int src = 16 * h + 4 * ss + t;  // lane owning k-step (4h+ss), row t
                const double rv = shfl_idx_d(rr, src);
                const double wv = HESS ? shfl_idx_d(ww, src) : 0.0;
                const int xr = 2 * t + (ss & 1);
                const unsigned char* rp = sp + (xr + 8 * (ss >> 1)) * 128 + ((g ^ xr) << 4);
                double av[NB8];
#pragma unroll
                for (int j = 0; j < NB16; ++j) {
                    const double2 v = *reinterpret_cast<const double2*>(rp + j * 2048);
                    av[2 * j] = v.x;
                    av[2 * j + 1] = v.y;
                }
#pragma unroll
                for (int jb = 0; jb < NB8; ++jb) gacc[jb] = fma(rv, av[jb], gacc[jb]);
                if (HESS) {
                    double bv[NB8];
#pragma unroll
                    for (int jb = 0; jb < NB8; ++jb) bv[jb] = wv * av[jb];
#pragma unroll
                    for (int I = 0; I < NB8; ++I) {
#pragma unroll
                        for (int J = 0; J <= I; ++J) {
                            const int tile = I * (I + 1) / 2 + J;
                            dmma884(acc[2 * tile], acc[2 * tile + 1], av[I], bv[J]);
                        }
                    }
                }
            }
            // slot consumed by every lane -> lane 0 asks TMA to refill it
            __syncwarp();
            if (lane == 0 && q_issue < nslots) issue_next();
        }
        ring_use = (s1 + 1 == NS) ? 0 : s1 + 1;
    }

    if (bad_domain) atomicOr(a.status, 1);

    // ---- epilogue: cross-warp reduction through the idle ring memory ----
    __syncwarp();
    double* scratch = reinterpret_cast<double*>(ring);  // EL doubles per warp (EL*8 <= RING_BYTES)
    static_assert(EL * 8 <= C::RING_BYTES, "reduction scratch must fit in the warp's ring");
    if (HESS) {
#pragma unroll
        for (int e = 0; e < 2 * NT; ++e) scratch[e * 32 + lane] = acc[e];
    }
    constexpr int GOFF = HESS ? 2 * NT : 0;
#pragma unroll
    for (int jb = 0; jb < NB8; ++jb) scratch[(GOFF + jb) * 32 + lane] = gacc[jb];
    scratch[(GOFF + NB8) * 32 + lane] = objacc;
    __syncthreads();
    double* out = a.partials + (size_t)blockIdx.x * EL;
    for (int e = threadIdx.x; e < EL; e += C::THREADS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += reinterpret_cast<const double*>(ring_all + w * C::RING_BYTES)[e];
        out[e] = s;
    }
}

}  // namespace anml
