// Banded evaluator for a model whose only variable is one B-spline (BASELINE config #3).
//
// A row of a spline design has degree+1 non-zeros, so the dense path (build the (n, nb) matrix,
// stream it through the DMMA kernel) moves 8*nb bytes per row to do (degree+1)^2 useful
// multiply-adds.  Here the basis is evaluated on the fly from x ("for C3 with
// on-the-fly basis P_X = 1, i.e. just x"); the kernel reads 8*(1 + #row vectors) bytes per row and
// is meant to be HBM-bound.
//
// Round-2 formulation (the round-1 kernel did a per-row interval search and kept a band block per
// lane; it was bound by instruction issue at 0.28 of the HBM roofline):
//   * the rows are sorted by x once at attach (radix sort of (x, row)), and the first sorted row of
//     every knot interval is found once (banded_segments_kernel, a binary search per knot), so a
//     knot interval is a CONTIGUOUS run of rows;
//   * every CTA owns an equal, contiguous slice of the sorted rows and walks the (one or two) knot
//     intervals that cross it: inside a run nothing depends on the row -- no search, no flush;
//   * on one interval every basis function is a polynomial of degree d in u = x - knot_left
//     (Taylor coefficients c[q][j] tabulated once per spline by banded_coeff_kernel from the same
//     local_basis_ders as the design kernel).  So the linear predictor is ONE polynomial
//     y = sum_j (sum_q c[q][j] beta[k+q]) u^j (d FMAs, coefficients folded per run), and the
//     gradient / Hessian band of the interval are linear in the power moments
//         G_j = sum_i r_i u_i^j   (j = 0..d)        M_s = sum_i w_i u_i^s   (s = 0..2d)
//     -- 3d+2 FMAs per row instead of (d+1) + (d+1)(d+2)/2 + the d(d+1) of the basis values.
//     banded_finish_kernel maps the reduced moments to the band:
//         grad[k+q] += sum_j c[q][j] G_j,   H[k+p][k+q] += sum_{a,b} c[p][a] c[q][b] M_{a+b}.
//     The change of basis costs a few bits (|c| u^j terms of alternating sign, ~35x the result),
//     i.e. ~1e-14 relative: two orders inside the 1e-10 / 1e-9 parity tolerances;
//   * two rows per thread per load (16-byte loads), next iteration's loads issued before the
//     current rows are processed;
//   * one partial record per CTA, one entry per (interval, moment): CTA-level reduction once per
//     run, records summed over CTAs in a fixed order (banded_reduce_kernel).  No floating-point
//     atomics, no memset of the records (every CTA writes its whole record): bitwise reproducible.
#pragma once
#include "common.cuh"
#include "finish.cuh"
#include "rowmath.cuh"
#include "spline.cuh"

namespace anml {

struct BandedArgs {
    const double* xs;       // abscissae, sorted ascending; n rounded up to even (+2) doubles allocated
    const double* obs;      // per-row vectors gathered into the same order (se / offset / weights may be null)
    const double* se;
    const double* offset;
    const double* weights;
    long long n;
    const double* knots;    // nk knots (device)
    const double* coef;     // [nk-1][degree+1 bases][degree+1 powers] Taylor coefficients about the left knot
    const long long* seg;   // [nk] first sorted row of interval k (seg[0] = 0, seg[nk-1] = n)
    int nk;
    const double* beta;     // nb = nk - 1 + degree coefficients
    int family, link, fast;
    long long rows_per_cta; // even
    double* partials;       // [gridDim.x][rec], rec = (nk-1) * NM + 1
    int* status;
};

template <int DEG>
struct BandedCfg {
    static constexpr int NG = DEG + 1;       // gradient moments
    static constexpr int NH = 2 * DEG + 1;   // Hessian moments
    static constexpr int NM = NG + NH;       // per interval
    static constexpr int THREADS = 256;
};

// EXTRAS: any of obs_se / offset / weights is present (compile-time, so that the plain case -- x and obs
// only, BASELINE config #3 -- keeps two loads per row pair and a small register footprint)
// Row pairs fetched ahead of the one being processed, per thread (registers), and resident CTAs per SM.
// Measured at n = 5e7 (profiles/r2_ab_banded.txt): 3 CTAs x 1 pair ahead 0.180 ms; 2 CTAs x 2 / 3 / 4 pairs ahead
// 0.223 / 0.223 / 0.225 ms; 3 CTAs x 2 pairs ahead spills (0.306 ms).  Occupancy beats prefetch depth here.
#ifndef ANML_BANDED_DEPTH
#define ANML_BANDED_DEPTH 1
#endif
#ifndef ANML_BANDED_MINB
#define ANML_BANDED_MINB 3
#endif
template <int DEG, bool HESS, bool EXTRAS>
__global__ void __launch_bounds__(256, EXTRAS ? 2 : ANML_BANDED_MINB) spline_banded_eval_kernel(const BandedArgs a) {
    using C = BandedCfg<DEG>;
    constexpr int NG = C::NG, NH = C::NH, NM = C::NM;
    __shared__ double red[8][NM + 1];
    const int nint = a.nk - 1;
    const int rec = nint * NM + 1;
    double* out = a.partials + (size_t)blockIdx.x * rec;
    for (int i = threadIdx.x; i < rec; i += blockDim.x) out[i] = 0.0;  // intervals this CTA never meets

    const long long lo = (long long)blockIdx.x * a.rows_per_cta;
    long long hi = lo + a.rows_per_cta;
    if (hi > a.n) hi = a.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double2* xs2 = reinterpret_cast<const double2*>(a.xs);
    const double2* ob2 = reinterpret_cast<const double2*>(a.obs);
    const double2* se2 = reinterpret_cast<const double2*>(a.se);
    const double2* of2 = reinterpret_cast<const double2*>(a.offset);
    const double2* wt2 = reinterpret_cast<const double2*>(a.weights);
    const bool has_se = EXTRAS && a.se != nullptr, has_off = EXTRAS && a.offset != nullptr, has_wt = EXTRAS && a.weights != nullptr;

    double objacc = 0.0;
    bool bad_domain = false;
    __syncthreads();  // the zero fill above is ordered before the run results below

    int k = 0;
    if (lo < hi) {
        while (k < nint - 1 && a.seg[k + 1] <= lo) ++k;
    }
    for (; lo < hi && k < nint; ++k) {
        const long long sa = a.seg[k], sb = a.seg[k + 1];
        if (sa >= hi) break;
        const long long ra = sa > lo ? sa : lo, rb = sb < hi ? sb : hi;
        if (ra >= rb) continue;  // empty interval (repeated knots)
        // ---- per-run constants: left knot, polynomial of the linear predictor
        const double knot = a.knots[k];
        double d[DEG + 1];
#pragma unroll
        for (int j = 0; j <= DEG; ++j) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q <= DEG; ++q) s = fma(a.coef[(k * (DEG + 1) + q) * (DEG + 1) + j], a.beta[k + q], s);
            d[j] = s;
        }
        double G[NG], M[NH];
#pragma unroll
        for (int j = 0; j < NG; ++j) G[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NH; ++j) M[j] = 0.0;

        auto row = [&](double x, double ob, double se, double off, double wt) {
            const double u = x - knot;
            double y = d[DEG];
#pragma unroll
            for (int j = DEG - 1; j >= 0; --j) y = fma(y, u, d[j]);
            const RowTerms1 rt = row_terms1<false>(a.family, a.link, a.fast != 0, y + off, ob, se);
            bad_domain |= !rt.ok;
            double l = rt.l, r = rt.r, w = rt.w;
            if (has_wt) {
                l *= wt; r *= wt; w *= wt;
            }
            objacc += l;
            double pw[NH];  // u^s
            pw[0] = 1.0;
            if (NH > 1) pw[1] = u;
#pragma unroll
            for (int s = 2; s < NH; ++s) pw[s] = pw[s >> 1] * pw[s - (s >> 1)];
            G[0] += r;
#pragma unroll
            for (int j = 1; j < NG; ++j) G[j] = fma(r, pw[j], G[j]);
            if (HESS) {
                M[0] += w;
#pragma unroll
                for (int s = 1; s < NH; ++s) M[s] = fma(w, pw[s], M[s]);
            }
        };
        struct Pair {
            double2 x, ob, se, off, wt;  // (se, off, wt fold to constants without EXTRAS)
        };
        auto load = [&](long long p, bool live) {
            Pair q;
            q.x = make_double2(0.0, 0.0); q.ob = q.x; q.off = q.x;
            q.se = make_double2(1.0, 1.0); q.wt = q.se;
            if (live) {
                q.x = xs2[p];
                q.ob = ob2[p];
                if (EXTRAS) {
                    if (has_se) q.se = se2[p];
                    if (has_off) q.off = of2[p];
                    if (has_wt) q.wt = wt2[p];
                }
            }
            return q;
        };
        // pairs (2p, 2p+1) that overlap [ra, rb); only the first and last pair can be partial
        const long long p0 = ra >> 1, p1 = (rb + 1) >> 1;
        constexpr int DEPTH = EXTRAS ? 2 : ANML_BANDED_DEPTH;
        long long p = p0 + threadIdx.x;
        Pair ahead[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) ahead[d] = load(p + d * C::THREADS, p + d * C::THREADS < p1);
        for (; p < p1; p += C::THREADS) {
            const Pair cur = ahead[0];
#pragma unroll
            for (int d = 0; d + 1 < DEPTH; ++d) ahead[d] = ahead[d + 1];
            ahead[DEPTH - 1] = load(p + DEPTH * C::THREADS, p + DEPTH * C::THREADS < p1);
            const long long i0 = 2 * p;
            if (i0 >= ra) row(cur.x.x, cur.ob.x, cur.se.x, cur.off.x, cur.wt.x);
            if (i0 + 1 < rb) row(cur.x.y, cur.ob.y, cur.se.y, cur.off.y, cur.wt.y);
        }
        // ---- CTA reduction of the run's moments (fixed order) -> record
#pragma unroll
        for (int j = 0; j < NG; ++j) G[j] = warp_sum(G[j]);
        if (HESS) {
#pragma unroll
            for (int j = 0; j < NH; ++j) M[j] = warp_sum(M[j]);
        }
        __syncthreads();  // red[] of the previous run has been read
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < NG; ++j) red[warp][j] = G[j];
#pragma unroll
            for (int j = 0; j < NH; ++j) red[warp][NG + j] = HESS ? M[j] : 0.0;
        }
        __syncthreads();
        if (threadIdx.x < NM) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
            out[k * NM + threadIdx.x] = s;
        }
    }
    objacc = warp_sum(objacc);
    __syncthreads();
    if (lane == 0) red[warp][NM] = objacc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][NM];
        out[rec - 1] = s;
    }
    if (bad_domain) atomicOr(a.status, 1);
}

// ---- the same evaluation with the row vectors staged through shared memory by bulk async copies ----
// The register-prefetch kernel above keeps 24 warps x 1 row pair in flight per SM (~25 KB): not enough
// outstanding bytes to saturate HBM (ncu: DRAM 56 % busy, long-scoreboard stalls 5.2 per issue), and a
// deeper register prefetch costs the occupancy it needs (profiles/r2_ab_banded.txt).  Here thread 0 of
// the CTA streams the CTA's slice in chunks of BULK_ROWS rows per vector with cp.async.bulk (the TMA
// engine, completion on an mbarrier) into a ring of BULK_STAGES stages: (stages - 1) x vectors x 8 KB in
// flight per CTA whatever the register budget; every thread pulls its four rows of a chunk into
// registers with two 16-byte shared-memory loads per vector, and its warp releases the stage at once
// (mbarrier `empty`, one arrival per warp): no CTA-wide barrier in the chunk loop.  Runs of constant knot
// interval are walked inside the chunk loop: a chunk that holds a run boundary is visited by both runs.
#ifndef ANML_BULK_ROWS_PLAIN
#define ANML_BULK_ROWS_PLAIN 1024
#endif
constexpr int BULK_ROWS_PLAIN = ANML_BULK_ROWS_PLAIN;  // rows per chunk and vector without / with extra row vectors
constexpr int BULK_ROWS_EXTRAS = 1024;                 // (8 KB per vector: four rows per thread)
constexpr int BULK_STAGES = 3;

#ifndef ANML_BULK_MINB
#define ANML_BULK_MINB 2  // the copies, not the warps, hide the HBM latency here: registers matter more than occupancy
#endif
template <int DEG, bool HESS, bool EXTRAS>
__global__ void __launch_bounds__(256, EXTRAS ? 2 : ANML_BULK_MINB) spline_banded_bulk_kernel(const BandedArgs a) {
    using C = BandedCfg<DEG>;
    constexpr int NG = C::NG, NH = C::NH, NM = C::NM;
    constexpr int BULK_ROWS = EXTRAS ? BULK_ROWS_EXTRAS : BULK_ROWS_PLAIN;
    constexpr int NPAIR = BULK_ROWS / 512;  // 16-byte row pairs per thread and chunk
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    __shared__ double red[8][NM + 1];
    __shared__ __align__(8) unsigned long long bar_mem[2 * BULK_STAGES];  // full[stage] (tx bytes), empty[stage] (8 warps)
    const int nint = a.nk - 1;
    const int rec = nint * NM + 1;
    double* out = a.partials + (size_t)blockIdx.x * rec;
    for (int i = threadIdx.x; i < rec; i += blockDim.x) out[i] = 0.0;  // intervals this CTA never meets

    const long long lo = (long long)blockIdx.x * a.rows_per_cta;  // even
    long long hi = lo + a.rows_per_cta;
    if (hi > a.n) hi = a.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool has_se = EXTRAS && a.se != nullptr, has_off = EXTRAS && a.offset != nullptr, has_wt = EXTRAS && a.weights != nullptr;
    // vectors of a stage: x, obs, then the extras that are present
    const double* vec[5] = {a.xs, a.obs, nullptr, nullptr, nullptr};
    int nv = 2, i_se = -1, i_off = -1, i_wt = -1;
    if (has_se) { i_se = nv; vec[nv++] = a.se; }
    if (has_off) { i_off = nv; vec[nv++] = a.offset; }
    if (has_wt) { i_wt = nv; vec[nv++] = a.weights; }
    const uint32_t stage_bytes = (uint32_t)nv * BULK_ROWS * 8;
    const uint32_t ring = smem_u32(bulk_smem);
    const uint32_t full = smem_u32(bar_mem), empty = full + BULK_STAGES * 8;
    const int nchunks = lo < hi ? (int)((hi - lo + BULK_ROWS - 1) / BULK_ROWS) : 0;

    auto issue = [&](int c) {  // thread 0: chunk c -> stage c % BULK_STAGES
        const int s = c % BULK_STAGES;
        const long long cs = lo + (long long)c * BULK_ROWS;
        long long rows = hi - cs < BULK_ROWS ? hi - cs : BULK_ROWS;
        rows += rows & 1;  // 16-byte granules; the arrays carry two spare doubles
        const uint32_t bytes = (uint32_t)rows * 8;
        mbar_arrive_expect_tx(full + s * 8, bytes * nv);
#pragma unroll
        for (int v = 0; v < 5; ++v)
            if (v < nv) bulk_load_1d(ring + s * stage_bytes + v * BULK_ROWS * 8, vec[v] + cs, bytes, full + s * 8);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < BULK_STAGES; ++s) {
            mbar_init(full + s * 8, 1);
            mbar_init(empty + s * 8, 8);
        }
        mbar_fence_init();
        for (int c = 0; c < nchunks && c < BULK_STAGES - 1; ++c) issue(c);
    }
    double objacc = 0.0;
    bool bad_domain = false;
    __syncthreads();  // barriers initialised; the zero fill above is ordered before the run results below

    int k = 0;
    if (lo < hi) {
        while (k < nint - 1 && a.seg[k + 1] <= lo) ++k;
    }
    double knot = 0.0, d[DEG + 1], G[NG], M[NH];
    long long run_end = hi;
    auto begin_run = [&]() {
        knot = a.knots[k];
        const long long e = a.seg[k + 1];
        run_end = e < hi ? e : hi;
#pragma unroll
        for (int j = 0; j <= DEG; ++j) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q <= DEG; ++q) s = fma(a.coef[(k * (DEG + 1) + q) * (DEG + 1) + j], a.beta[k + q], s);
            d[j] = s;
        }
#pragma unroll
        for (int j = 0; j < NG; ++j) G[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NH; ++j) M[j] = 0.0;
    };
    auto end_run = [&]() {  // CTA reduction of the run's moments (fixed order) -> record
#pragma unroll
        for (int j = 0; j < NG; ++j) G[j] = warp_sum(G[j]);
        if (HESS) {
#pragma unroll
            for (int j = 0; j < NH; ++j) M[j] = warp_sum(M[j]);
        }
        __syncthreads();  // red[] of the previous run has been read
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < NG; ++j) red[warp][j] = G[j];
#pragma unroll
            for (int j = 0; j < NH; ++j) red[warp][NG + j] = HESS ? M[j] : 0.0;
        }
        __syncthreads();
        if (threadIdx.x < NM) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
            out[k * NM + threadIdx.x] = s;
        }
    };
    auto row = [&](double x, double ob, double se, double off, double wt) {
        const double u = x - knot;
        double y = d[DEG];
#pragma unroll
        for (int j = DEG - 1; j >= 0; --j) y = fma(y, u, d[j]);
        const RowTerms1 rt = row_terms1<false>(a.family, a.link, a.fast != 0, y + off, ob, se);
        bad_domain |= !rt.ok;
        double l = rt.l, r = rt.r, w = rt.w;
        if (has_wt) {
            l *= wt; r *= wt; w *= wt;
        }
        objacc += l;
        double pw[NH];  // u^s
        pw[0] = 1.0;
        if (NH > 1) pw[1] = u;
#pragma unroll
        for (int s = 2; s < NH; ++s) pw[s] = pw[s >> 1] * pw[s - (s >> 1)];
        G[0] += r;
#pragma unroll
        for (int j = 1; j < NG; ++j) G[j] = fma(r, pw[j], G[j]);
        if (HESS) {
            M[0] += w;
#pragma unroll
            for (int s = 1; s < NH; ++s) M[s] = fma(w, pw[s], M[s]);
        }
    };

    if (lo < hi) begin_run();
    for (int c = 0; c < nchunks; ++c) {
        const int s = c % BULK_STAGES;
        // keep BULK_STAGES - 1 chunks in flight: the stage of chunk c - 1 is refilled once all 8 warps left it
        if (threadIdx.x == 0 && c + BULK_STAGES - 1 < nchunks) {
            if (c > 0) mbar_wait(empty + ((c - 1) % BULK_STAGES) * 8, (uint32_t)(((c - 1) / BULK_STAGES) & 1));
            issue(c + BULK_STAGES - 1);
        }
        mbar_wait(full + s * 8, (uint32_t)((c / BULK_STAGES) & 1));
        const long long cs = lo + (long long)c * BULK_ROWS;
        const long long ce = cs + BULK_ROWS < hi ? cs + BULK_ROWS : hi;
        // this thread's rows of the chunk: NPAIR pairs, 512 rows apart (coalesced 16-byte shared-memory loads)
        const uint32_t at = ring + s * stage_bytes + threadIdx.x * 16;
        double2 x2[NPAIR], ob2[NPAIR], se2[NPAIR], of2[NPAIR], wt2[NPAIR];
#pragma unroll
        for (int h = 0; h < NPAIR; ++h) {
            x2[h] = lds128(at + h * 4096);
            ob2[h] = lds128(at + h * 4096 + BULK_ROWS * 8);
            se2[h] = make_double2(1.0, 1.0); of2[h] = make_double2(0.0, 0.0); wt2[h] = make_double2(1.0, 1.0);
            if (EXTRAS) {
                if (has_se) se2[h] = lds128(at + h * 4096 + i_se * BULK_ROWS * 8);
                if (has_off) of2[h] = lds128(at + h * 4096 + i_off * BULK_ROWS * 8);
                if (has_wt) wt2[h] = lds128(at + h * 4096 + i_wt * BULK_ROWS * 8);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s * 8);  // this warp holds its rows in registers: the stage may be refilled
        const long long r0 = cs + 2 * threadIdx.x;
        if (ce <= run_end) {
            // common case: the whole chunk lies in the current run
#pragma unroll
            for (int h = 0; h < NPAIR; ++h) {
                const long long r = r0 + 512 * h;
                if (r < ce) row(x2[h].x, ob2[h].x, se2[h].x, of2[h].x, wt2[h].x);
                if (r + 1 < ce) row(x2[h].y, ob2[h].y, se2[h].y, of2[h].y, wt2[h].y);
            }
        } else {
            long long pos = cs;
            while (pos < ce) {
                const long long stop = ce < run_end ? ce : run_end;
#pragma unroll
                for (int h = 0; h < NPAIR; ++h) {
                    const long long r = r0 + 512 * h;
                    if (r >= pos && r < stop) row(x2[h].x, ob2[h].x, se2[h].x, of2[h].x, wt2[h].x);
                    if (r + 1 >= pos && r + 1 < stop) row(x2[h].y, ob2[h].y, se2[h].y, of2[h].y, wt2[h].y);
                }
                pos = stop;
                if (pos == run_end && pos < hi) {  // the run ends inside this chunk
                    end_run();
                    ++k;
                    while (k < nint - 1 && a.seg[k + 1] <= pos) ++k;  // empty intervals (repeated knots)
                    begin_run();
                }
            }
        }
        if (ce == run_end && ce < hi) {  // the run ends exactly at the chunk's end
            end_run();
            ++k;
            while (k < nint - 1 && a.seg[k + 1] <= ce) ++k;
            begin_run();
        }
    }
    if (lo < hi) end_run();
    objacc = warp_sum(objacc);
    __syncthreads();
    if (lane == 0) red[warp][NM] = objacc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][NM];
        out[rec - 1] = s;
    }
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

// seg[k] = first sorted row with x >= knots[k] (k = 1..nk-2), seg[0] = 0, seg[nk-1] = n: rows of interval
// k are [seg[k], seg[k+1]) -- the same integer as clip(searchsorted(knots, x, "right") - 1, 0, nk-2)
__global__ void banded_segments_kernel(const double* __restrict__ xs, long long n, const double* __restrict__ knots, int nk,
                                       long long* __restrict__ seg) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nk) return;
    if (k == 0) {
        seg[0] = 0;
        return;
    }
    if (k == nk - 1) {
        seg[k] = n;
        return;
    }
    const double kn = knots[k];
    long long lo = 0, hi = n;  // first index with xs[i] >= kn
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (xs[mid] < kn) lo = mid + 1; else hi = mid;
    }
    seg[k] = lo;
}

// partial records summed over CTAs in a fixed order: a 32 x 32 block owns 32 entries, row ty of the
// block walks the records ty, ty + 32, ... (loads coalesce across the entries), then the 32 row sums are
// added in order.
__global__ void __launch_bounds__(1024) banded_reduce_kernel(const double* __restrict__ partials, long long nrec, int rec,
                                                            double* __restrict__ reduced) {
    __shared__ double part[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int e = blockIdx.x * 32 + tx;
    double s0 = 0.0, s1 = 0.0;
    if (e < rec) {
        long long w = ty;
        for (; w + 32 < nrec; w += 64) {
            s0 += partials[w * rec + e];
            s1 += partials[(w + 32) * rec + e];
        }
        if (w < nrec) s0 += partials[w * rec + e];
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

// reduced moments -> packed [obj | grad(P) | hess(P x P) | flag], Gaussian prior terms and the domain
// flag included (one launch; round 1 used assemble + add_priors + set_flag).  One thread per packed
// entry; entry (i, j) gathers the intervals that hold both bases.
struct BandedFinishArgs {
    const double* reduced;  // [(nk-1) * NM + 1]
    const double* coef;     // [nk-1][deg+1][deg+1]
    int nint, deg, hess, P;
    double* packed;
    const double* beta;
    PriorDev prior;
    int* status;
};

__global__ void __launch_bounds__(256) banded_finish_kernel(const BandedFinishArgs a) {
    extern __shared__ double res_s[];  // unscaled residuals of the linear prior rows (gradient / objective CTA)
    const int deg = a.deg, ng = deg + 1, nm = ng + 2 * deg + 1, P = a.P;
    const int total = 1 + P + (a.hess ? P * P : 0);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool lin = a.prior.on && a.prior.m > 0;
    if (lin && (long long)blockIdx.x * blockDim.x < 1 + P) {  // CTA-uniform
        for (int i = threadIdx.x; i < a.prior.m; i += blockDim.x) {
            double s = 0.0;
            for (int j = 0; j < P; ++j) s = fma(a.prior.lin_mat[(size_t)i * P + j], a.beta[j], s);
            res_s[i] = s - a.prior.lin_mean[i];
        }
        __syncthreads();
    }
    if (e >= total) return;
    if (e == 0) {
        double o = a.reduced[a.nint * nm];
        if (a.prior.on) {
            for (int j = 0; j < P; ++j) {
                const double dd = a.beta[j] - a.prior.mean[j];
                o += 0.5 * dd * dd * a.prior.inv_var[j];
            }
            for (int i = 0; i < a.prior.m; ++i) o += 0.5 * res_s[i] * res_s[i] * a.prior.lin_inv_var[i];
        }
        a.packed[0] = o;
        a.packed[1 + P + (size_t)P * P] = (double)(*a.status);
        *a.status = 0;
    } else if (e <= P) {
        const int i = e - 1;
        double s = 0.0;
        for (int k = (i - deg > 0 ? i - deg : 0); k <= i && k < a.nint; ++k) {
            const double* c = a.coef + (k * ng + (i - k)) * ng;
            const double* Gk = a.reduced + k * nm;
            double v = 0.0;
            for (int j = 0; j < ng; ++j) v = fma(c[j], Gk[j], v);
            s += v;
        }
        if (a.prior.on) {
            double pg = (a.beta[i] - a.prior.mean[i]) * a.prior.inv_var[i];
            for (int r = 0; r < a.prior.m; ++r) pg = fma(a.prior.lin_mat[(size_t)r * P + i], res_s[r] * a.prior.lin_inv_var[r], pg);
            s += pg;
        }
        a.packed[e] = s;
    } else {
        const int i = (e - 1 - P) / P, j = (e - 1 - P) % P;
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        double s = 0.0;
        for (int k = (hi - deg > 0 ? hi - deg : 0); k <= lo && k < a.nint; ++k) {
            const double* cp = a.coef + (k * ng + (hi - k)) * ng;
            const double* cq = a.coef + (k * ng + (lo - k)) * ng;
            const double* Mk = a.reduced + k * nm + ng;
            double v = 0.0;
            for (int x = 0; x < ng; ++x) {
                double t = 0.0;
                for (int y = 0; y < ng; ++y) t = fma(cq[y], Mk[x + y], t);
                v = fma(cp[x], t, v);
            }
            s += v;
        }
        if (a.prior.on) {
            if (a.prior.m > 0) s += a.prior.lin_hess[(size_t)hi * P + lo];
            if (i == j) s += a.prior.inv_var[i];
        }
        a.packed[e] = s;  // (i, j) and (j, i) evaluate the same expression: exactly symmetric
    }
}

__global__ void iota_kernel(int32_t* p, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}

__global__ void gather_rows_kernel(double* __restrict__ dst, const double* __restrict__ src, const int32_t* __restrict__ perm, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[perm[i]];
}

}  // namespace anml
