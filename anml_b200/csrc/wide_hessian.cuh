// Hessian blocks wider than one 64-column tile:  H[ro.., co..] = Xa^T diag(w) Xb
// for designs of any width (BASELINE config #5: P = 1024), or for two different
// designs (cross block of a two-parameter model).
//
// GEMM-shaped: output tiled in 128 (columns of Xa) x 64 (columns of Xb) blocks, the
// reduction dimension (rows of X, 1e6..1e8) split over CTAs.  CTA = 4 DMMA warps, ONE
// PER SCHEDULER (two DMMA-issuing warps on a scheduler interfere: 21 vs 36 TFLOP/s in
// tools/microbench/mstream2_bench.cu), each owning 64 x 32 of the block (32
// DMMA.8x8x4 accumulator tiles = 128 registers), + 1 producer warp.
//   * stage = 16 rows: A panel (16 x 128 of Xa, 8 TMA boxes, 128B swizzle), B panel
//     (16 x 64 of Xb, 4 boxes; on blocks that touch the diagonal of one design it is a
//     half of the A panel and is not loaded) and the 16 row weights (cp.async.bulk);
//     mbarrier `full` (tx bytes) / `empty` (4 warps);
//   * per 4-row k-step a warp loads 4 + 2 LDS.128 fragments, scales the B fragments
//     by w (4 DMUL), issues 32 DMMAs;
//   * blockIdx = chunk * npairs + pair: CTAs that run together read the same rows of
//     X for different column blocks, so the panels are served by L2, not HBM;
//   * symmetric case (same design, diagonal parameter block): only block pairs with
//     bj <= 2*bi+1 are launched and warps entirely above the diagonal skip their DMMAs;
//   * each CTA writes its 128 x 64 partial; wide_reduce_kernel sums the row chunks in a
//     fixed order and mirrors the lower triangle (exactly symmetric, no atomics).
#pragma once
#include "common.cuh"

namespace anml {

constexpr int WIDE_STAGES = 6;
constexpr int WIDE_PANEL_BYTES = 16384 + 8192;  // A panel + B panel
constexpr int WIDE_THREADS = 160;               // 4 DMMA warps + 1 producer warp
constexpr int WIDE_TILE_ELEMS = 128 * 64;
constexpr int WIDE_SMEM_BYTES = 1024 + WIDE_STAGES * WIDE_PANEL_BYTES + WIDE_STAGES * 128 + WIDE_STAGES * 2 * 8;

struct WideArgs {
    long long n_pad;      // rows, multiple of 32 (w is zero beyond n, X is zero-filled by TMA)
    long long chunk_rows; // rows per CTA along the reduction, multiple of 16
    int npairs;           // block pairs per chunk
    int nbj;              // 64-column blocks of Xb (non-symmetric enumeration)
    int symmetric;        // 1: pairs enumerate bj <= 2*bi+1 of one design
    int shared_panels;    // 1: Xa and Xb are the same matrix
    const double* w;      // device (n_pad)
    double* partials;     // [nchunks][npairs][128*64]
};

// pair -> (bi: 128-column block of Xa, bj: 64-column block of Xb)
__device__ __forceinline__ void pair_to_blocks(int pair, int symmetric, int nbj, int& bi, int& bj) {
    if (symmetric) {
        bi = 0;  // bi*(bi+1) pairs precede row bi (row bi holds 2*bi+2 blocks)
        while ((bi + 1) * (bi + 2) <= pair) ++bi;
        bj = pair - bi * (bi + 1);
    } else {
        bi = pair / nbj;
        bj = pair - bi * nbj;
    }
}

__global__ void __launch_bounds__(WIDE_THREADS, 1)
wide_hessian_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const WideArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t panels = base;                                          // [stage][A 16 KB | B 8 KB]
    const uint32_t wbase = panels + WIDE_STAGES * WIDE_PANEL_BYTES;        // [stage][16 doubles]
    const uint32_t full_bar = wbase + WIDE_STAGES * 128;                   // [stage]
    const uint32_t empty_bar = full_bar + WIDE_STAGES * 8;                 // [stage]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int chunk = blockIdx.x / a.npairs;
    const int pair = blockIdx.x - chunk * a.npairs;
    int bi, bj;
    pair_to_blocks(pair, a.symmetric, a.nbj, bi, bj);
    // B columns [64 bj, +64) lie inside A columns [128 bi, +128) of the same matrix: reuse the A panel
    const bool b_in_a = a.shared_panels && (bj >> 1) == bi;
    const bool diag = a.symmetric && (bj >> 1) == bi;

    const long long row_begin = (long long)chunk * a.chunk_rows;
    long long row_end = row_begin + a.chunk_rows;
    if (row_end > a.n_pad) row_end = a.n_pad;
    const int nstages = row_end > row_begin ? (int)((row_end - row_begin) >> 4) : 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < WIDE_STAGES; ++s) {
            mbar_init(full_bar + s * 8, 1);
            mbar_init(empty_bar + s * 8, 4);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 4) {
        // ------------------------------ producer ------------------------------
        if (lane == 0) {
            const uint32_t bytes = (b_in_a ? 16384u : 24576u) + 128u;
            for (int it = 0; it < nstages; ++it) {
                const int s = it % WIDE_STAGES;
                const int use = it / WIDE_STAGES;
                if (use > 0) mbar_wait(empty_bar + s * 8, (uint32_t)((use - 1) & 1));
                const long long row0 = row_begin + (long long)it * 16;
                const uint32_t bar = full_bar + s * 8;
                const uint32_t dst = panels + s * WIDE_PANEL_BYTES;
                mbar_arrive_expect_tx(bar, bytes);
#pragma unroll
                for (int j = 0; j < 8; ++j) tma_load_2d(dst + j * 2048, &tmap_a, bi * 128 + 16 * j, (int)row0, bar);
                if (!b_in_a) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) tma_load_2d(dst + 16384 + j * 2048, &tmap_b, bj * 64 + 16 * j, (int)row0, bar);
                }
                bulk_load_1d(wbase + s * 128, a.w + row0, 128u, bar);
            }
        }
    } else {
        // ------------------------------ DMMA warps ----------------------------
        const int wi = warp >> 1, wj = warp & 1;  // 64-column slice of A, 32-column slice of B
        // skip sub-tiles that lie entirely above the diagonal: rows 128 bi + 64 wi + [0,64), cols 64 bj + 32 wj + [0,32)
        const bool active = !(diag && (128 * bi + 64 * wi + 63 < 64 * bj + 32 * wj));
        double acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.0;
        const uint32_t a_off = (uint32_t)(4 * wi) * 2048;
        const uint32_t b_off = (b_in_a ? (uint32_t)(4 * (bj & 1)) * 2048 : 16384u) + (uint32_t)(2 * wj) * 2048;
        for (int it = 0; it < nstages; ++it) {
            const int s = it % WIDE_STAGES;
            mbar_wait(full_bar + s * 8, (uint32_t)((it / WIDE_STAGES) & 1));
            if (active) {
                const uint32_t sp = panels + s * WIDE_PANEL_BYTES;
                // two fragment sets: k-step ss+1 is loaded (and its B fragments scaled) behind the DMMAs of k-step ss
                double av[2][8], bv[2][4], wv[2];
                auto load_frags = [&](int set, int ss) {
                    const int xr = 2 * t + (ss & 1);
                    const uint32_t rowoff = (uint32_t)(xr + 8 * (ss >> 1)) * 128 + (uint32_t)((g ^ xr) << 4);
                    wv[set] = lds64(wbase + s * 128 + (uint32_t)(xr + 8 * (ss >> 1)) * 8);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double2 v = lds128(sp + a_off + j * 2048 + rowoff);
                        av[set][2 * j] = v.x;
                        av[set][2 * j + 1] = v.y;
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const double2 v = lds128(sp + b_off + j * 2048 + rowoff);
                        bv[set][2 * j] = v.x;
                        bv[set][2 * j + 1] = v.y;
                    }
                };
                load_frags(0, 0);
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[0][j] *= wv[0];
#pragma unroll
                for (int ss = 0; ss < 4; ++ss) {
                    const int cur = ss & 1, nxt = cur ^ 1;
                    if (ss + 1 < 4) load_frags(nxt, ss + 1);
#pragma unroll
                    for (int I = 0; I < 8; ++I)
#pragma unroll
                        for (int J = 0; J < 4; ++J)
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(acc[2 * (I * 4 + J)]), "+d"(acc[2 * (I * 4 + J) + 1]) : "d"(av[cur][I]), "d"(bv[cur][J]));
                    if (ss + 1 < 4) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) asm volatile("mul.f64 %0, %0, %1;" : "+d"(bv[nxt][j]) : "d"(wv[nxt]));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + s * 8);
        }
        // scatter the warp's 64 x 32 slice into the CTA's 128 x 64 partial (row-major)
        double* out = a.partials + ((size_t)chunk * a.npairs + pair) * WIDE_TILE_ELEMS;
#pragma unroll
        for (int I = 0; I < 8; ++I)
#pragma unroll
            for (int J = 0; J < 4; ++J)
#pragma unroll
                for (int reg = 0; reg < 2; ++reg) {
                    const int r = 64 * wi + 16 * (I >> 1) + 2 * g + (I & 1);
                    const int c = 32 * wj + 16 * (J >> 1) + 2 * (2 * t + reg) + (J & 1);
                    out[r * 64 + c] = active ? acc[2 * (I * 4 + J) + reg] : 0.0;
                }
    }
}

// H[ro + i][co + j] = sum over row chunks of the partial tiles (fixed order); the
// symmetric case mirrors the lower triangle, an off-diagonal parameter block is also
// written transposed.
__global__ void __launch_bounds__(256) wide_reduce_kernel(const double* __restrict__ partials, int nchunks, int npairs, int nbj,
                                                          int symmetric, int pa, int pb, double* __restrict__ H, int P_total,
                                                          int ro, int co) {
    const long long total = (long long)npairs * WIDE_TILE_ELEMS;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int pair = (int)(e >> 13);
        const int idx = (int)(e & (WIDE_TILE_ELEMS - 1));
        int bi, bj;
        pair_to_blocks(pair, symmetric, nbj, bi, bj);
        const int i = bi * 128 + (idx >> 6), j = bj * 64 + (idx & 63);
        if (i >= pa || j >= pb) continue;
        if (symmetric && j > i) continue;
        double s = 0.0;
        for (int c = 0; c < nchunks; ++c) s += partials[((size_t)c * npairs + pair) * WIDE_TILE_ELEMS + idx];
        H[(size_t)(ro + i) * P_total + co + j] = s;
        if (symmetric) H[(size_t)(ro + j) * P_total + co + i] = s;
        if (ro != co) H[(size_t)(co + j) * P_total + ro + i] = s;
    }
}

}  // namespace anml
